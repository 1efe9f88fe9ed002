"""Same-box sweep of PGD-kernel build variants (csrc/pgd.cu macros): builds each variant next to the product library
(cross-compiled beforehand with `python tools/pgd_sweep.py --build`, so that the .so files travel with the snapshot) and
times tools/pgd_time.py against each through the ctypes binding."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "base": [],
    "ilv": ["-DRMCL_PGD_INTERLEAVE=1"],
    "ilv_c64": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_CHUNK_KB=64"],
    "ilv_c16": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_CHUNK_KB=16"],
    "ilv_b80": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_BATCH_MB=80", "-DRMCL_PGD_BATCH_MB_L2=80"],
    "ilv_b12": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_BATCH_MB=12", "-DRMCL_PGD_BATCH_MB_L2=24"],
    "ilv_cta6": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_MIN_CTAS=6"],
    "ilv_cta4": ["-DRMCL_PGD_INTERLEAVE=1", "-DRMCL_PGD_MIN_CTAS=4"],
}
if "--build" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "robust-multimodal-contrastive-learning_b200", "csrc"))
    import build
    for name, defs in VARIANTS.items():
        print(build.build(force=True, out=os.path.join(os.path.dirname(build.OUT), f"librmcl_b200_pgd_{name}.so"), defines=defs))
else:
    for name in VARIANTS:
        env = dict(os.environ, RMCL_B200_LIB=f"librmcl_b200_pgd_{name}.so", RMCL_B200_FFI="ctypes")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pgd_time.py")], env=env, check=False)
