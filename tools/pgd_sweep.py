"""Same-box sweep of PGD-kernel build variants (csrc/pgd.cu macros).  `--build` (run here, nvcc cross-compiles) links one
library per variant next to the product library — only pgd.cu is recompiled per variant — so that the .so files travel with
the snapshot; without arguments (on the GPU box) tools/pgd_time.py is timed against each through the ctypes binding."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robust-multimodal-contrastive-learning_b200")
VARIANTS = {
    "base": [],
    "pf2": ["-DRMCL_PGD_L2_PREFETCH=2"],
    "u8": ["-DRMCL_PGD_P1_UNROLL=8"],
    "u8_cta4": ["-DRMCL_PGD_P1_UNROLL=8", "-DRMCL_PGD_MIN_CTAS=4"],
    "pf2_u8": ["-DRMCL_PGD_L2_PREFETCH=2", "-DRMCL_PGD_P1_UNROLL=8"],
    "pf2_norm_only": ["-DRMCL_PGD_L2_PREFETCH=2", "-DRMCL_PGD_EXPERIMENT=18"],
    "u8_norm_only": ["-DRMCL_PGD_P1_UNROLL=8", "-DRMCL_PGD_EXPERIMENT=18"],
}
if "--build" in sys.argv:
    sys.path.insert(0, os.path.join(PKG, "csrc"))
    import build
    objdir = os.path.join(ROOT, "build", "sweep")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in build.FLAGS if f not in ("-shared", "-cudart", "static")]
    objs = []
    for src in build.SOURCES:
        if src == "pgd.cu":
            continue
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if not os.path.isfile(o) or os.path.getmtime(o) < max(os.path.getmtime(os.path.join(build.HERE, f)) for f in os.listdir(build.HERE)):
            subprocess.run(["nvcc"] + flags + ["-c", src, "-o", o], cwd=build.HERE, check=True)
        objs.append(o)
    for name, defs in VARIANTS.items():
        o = os.path.join(objdir, f"pgd_{name}.o")
        subprocess.run(["nvcc"] + flags + defs + ["-c", "pgd.cu", "-o", o], cwd=build.HERE, check=True)
        out = os.path.join(PKG, f"librmcl_b200_pgd_{name}.so")
        subprocess.run(["nvcc", "-shared", "-cudart", "static", "-o", out, o] + objs, check=True)
        print(out, flush=True)
else:
    for name in VARIANTS:
        env = dict(os.environ, RMCL_B200_LIB=f"librmcl_b200_pgd_{name}.so", RMCL_B200_FFI="ctypes")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pgd_time.py")], env=env, check=False)
