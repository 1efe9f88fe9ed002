"""Builds librmcl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "librmcl_b200.so")
SOURCES = ["capi.cu", "ema.cu", "enqueue.cu", "pgd.cu", "infonce.cu", "infonce_simt.cu", "infonce_tc.cu", "infonce_tc2.cu",
           "queue_stats.cu", "barlow.cu", "barlow_gram.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-shared", "-cudart", "static"]


def needs_rebuild():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "rmcl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=OUT, defines=()):
    """``out``/``defines`` build an experiment variant next to the product library (tools/ab_*.py).
    Every source is compiled to its own object (in parallel, cached by modification time under build/obj) and the
    objects are linked into the shared library: a change to one kernel costs one nvcc invocation."""
    if out == OUT and not force and not needs_rebuild():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tag = "product" if (out == OUT and not defines) else os.path.splitext(os.path.basename(out))[0]
    objdir = os.path.join(os.path.dirname(os.path.dirname(HERE)), "build", "obj", tag)
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in FLAGS if f not in ("-shared", "-cudart", "static")] + list(defines) + (["-Xptxas", "-v"] if verbose else [])
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "rmcl_b200.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        stale = force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(newest_header, os.path.getmtime(os.path.join(HERE, src)))
        if not stale:
            return obj, None
        r = subprocess.run([nvcc] + cflags + ["-c", src, "-o", obj], cwd=HERE, capture_output=True, text=True)
        return obj, r

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = ""
    for obj, r in results:
        if r is None:
            continue
        log += r.stdout + r.stderr
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {os.path.basename(obj)}")
    r = subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] +
                       [obj for obj, _ in results], cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking librmcl_b200.so")
    if verbose:
        print(log)
    return out


TORCH_OUT = os.path.join(os.path.dirname(HERE), "rmcl_b200_torch.so")


def build_torch_ext(force=False):
    """rmcl_b200_torch.so: the TORCH_LIBRARY(rmcl, ...) operators of torch_ext.cpp, host C++ only (g++), linked against
    librmcl_b200.so next to it (rpath $ORIGIN) and torch's own libraries.  In-tree like the CUDA library, so it travels
    with the snapshot."""
    src = os.path.join(HERE, "torch_ext.cpp")
    deps = [src, os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "rmcl_b200.h")]
    if not force and os.path.isfile(TORCH_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_OUT) for d in deps):
        return TORCH_OUT
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", src, "-o", TORCH_OUT,
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home}/include"]
    cmd += [f"-L{p}" for p in ce.library_paths()] + [f"-L{cuda_home}/lib64", f"-L{os.path.dirname(HERE)}"]
    cmd += ["-l:librmcl_b200.so", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-lcudart",
            "-Wl,-rpath,$ORIGIN"] + [f"-Wl,-rpath,{p}" for p in ce.library_paths()]
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building rmcl_b200_torch.so")
    return TORCH_OUT


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = OUT
    if "--out" in sys.argv:
        out = os.path.join(os.path.dirname(HERE), sys.argv[sys.argv.index("--out") + 1])
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out, defines=defs))
    if out == OUT:
        print(build_torch_ext(force="--force" in sys.argv))
