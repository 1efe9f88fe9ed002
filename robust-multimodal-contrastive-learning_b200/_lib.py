"""ctypes binding of librmcl_b200.so (include/rmcl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this
raises.  Build it with ``python __graft_entry__.py build`` (or ``csrc/build.py``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RMCL_B200_LIB selects an experiment build of the same library (csrc/build.py --out ... -D...); it must
# exist — there is still no fallback.
LIB_PATH = os.path.join(_HERE, os.environ.get("RMCL_B200_LIB", "librmcl_b200.so"))

RMCL_F32, RMCL_BF16, RMCL_BF16_HILO = 0, 1, 2
PGD_MODES = {"ref_linf": 0, "sign_linf": 1, "l2": 2}
INFONCE_PATHS = {"auto": 0, "simt": 1, "tcgen05": 2}
BARLOW_PATHS = {"auto": 0, "direct": 1, "gram": 2}
FLAG_NORMALIZE_K, FLAG_NO_GRAD, FLAG_DEBUG_PARTIAL_ONLY = 1, 2, 4

EXPORTS = (
    "rmcl_last_error", "rmcl_version", "rmcl_sm_count", "rmcl_ema_plan", "rmcl_ema_multi",
    "rmcl_infonce_workspace_bytes", "rmcl_infonce_fwd_bwd", "rmcl_enqueue", "rmcl_pgd_workspace_bytes", "rmcl_pgd_step", "rmcl_step_host",
    "rmcl_profile_enable", "rmcl_profile_infonce_ms", "rmcl_enqueue_shadow", "rmcl_debug_tc_timeline",
    "rmcl_debug_tc_timeline_words", "rmcl_queue_stats", "rmcl_infonce_fwd_bwd_diag",
    "rmcl_barlow_workspace_bytes", "rmcl_barlow_fwd_bwd", "rmcl_gather_enqueue_p2p", "rmcl_infonce_describe", "rmcl_queue_split",
)


class EmaChunk(C.Structure):
    _fields_ = [("k", C.c_void_p), ("q", C.c_void_p), ("n", C.c_uint64)]


class RmclError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"rmcl_b200: {LIB_PATH} is missing — the CUDA library has not been built and there is no "
            "CPU fallback. Run `python __graft_entry__.py build`.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, u32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint, C.c_size_t
    L.rmcl_last_error.restype = C.c_char_p
    L.rmcl_last_error.argtypes = []
    L.rmcl_version.restype = i32
    L.rmcl_sm_count.restype = i32
    L.rmcl_ema_plan.restype = i64
    L.rmcl_ema_plan.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64), i32, i32, C.c_uint64,
                                C.POINTER(EmaChunk)]
    L.rmcl_ema_multi.restype = i32
    L.rmcl_ema_multi.argtypes = [vp, i64, f64, i32, vp]
    L.rmcl_infonce_workspace_bytes.restype = sz
    L.rmcl_infonce_workspace_bytes.argtypes = [i32, i32, i64, i32, i32]
    L.rmcl_infonce_fwd_bwd.restype = i32
    L.rmcl_infonce_fwd_bwd.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i64, i64, f32, f32, u32, i32,
                                       vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.rmcl_infonce_describe.restype = i32
    L.rmcl_infonce_describe.argtypes = [i32, i32, i64, i32, i32, i32, C.c_char_p, sz]
    L.rmcl_enqueue.restype = i32
    L.rmcl_enqueue.argtypes = [vp, i32, vp, i32, vp, i32, i32, i64, i64, vp]
    if "RMCL_B200_LIB" in os.environ and not hasattr(L, "rmcl_enqueue_shadow"):   # A/B against an older build
        L.rmcl_enqueue_shadow = L.rmcl_debug_tc_timeline = L.rmcl_debug_tc_timeline_words = L.rmcl_version
    L.rmcl_enqueue_shadow.restype = i32
    L.rmcl_enqueue_shadow.argtypes = [vp, i32, vp, i64, i32, vp, i32, vp, i32, i32, i64, i64, vp]
    L.rmcl_queue_split.restype = i32
    L.rmcl_queue_split.argtypes = [vp, i32, i64, i64, vp, i64, vp]
    L.rmcl_debug_tc_timeline.restype = i32
    L.rmcl_debug_tc_timeline.argtypes = [vp]
    L.rmcl_debug_tc_timeline_words.restype = i32
    L.rmcl_debug_tc_timeline_words.argtypes = []
    if hasattr(L, "rmcl_queue_stats"):
        L.rmcl_queue_stats.restype = i32
        L.rmcl_queue_stats.argtypes = [vp, i32, i32, i64, i64, f32, vp, vp, vp, vp]
        L.rmcl_infonce_fwd_bwd_diag.restype = i32
        L.rmcl_infonce_fwd_bwd_diag.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i64, i64, f32, f32, u32, i32,
                                                vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, sz, vp]
    L.rmcl_barlow_workspace_bytes.restype = sz
    L.rmcl_barlow_workspace_bytes.argtypes = [i32, i32]
    L.rmcl_barlow_fwd_bwd.restype = i32
    L.rmcl_barlow_fwd_bwd.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, f32, f32, f32, f32, f32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.rmcl_gather_enqueue_p2p.restype = i32
    L.rmcl_gather_enqueue_p2p.argtypes = [vp, vp, vp, vp, i32, vp, i64, i32, vp, i32, i32, i32, i32, i64, i64, vp]
    L.rmcl_pgd_step.restype = i32
    L.rmcl_pgd_step.argtypes = [vp, i32, vp, i32, i32, i64, f32, f32, i32, vp, sz, vp]
    L.rmcl_pgd_workspace_bytes.restype = sz
    L.rmcl_pgd_workspace_bytes.argtypes = [i32, i64, i32]
    L.rmcl_step_host.restype = i32
    L.rmcl_step_host.argtypes = [vp, i64, f64, i32, vp, vp, i32, vp, vp, vp, i32, vp, i32, i32, i64, f32, i32,
                                 vp, vp, vp, vp, vp, vp, sz, vp]
    L.rmcl_profile_enable.restype = i32
    L.rmcl_profile_enable.argtypes = [i32]
    L.rmcl_profile_infonce_ms.restype = i32
    L.rmcl_profile_infonce_ms.argtypes = [C.POINTER(C.c_float)]
    _lib = L
    return L


TORCH_LIB_PATH = os.path.join(_HERE, "rmcl_b200_torch.so")
_torch_ops = None


def ffi():
    """Which binding the device-tensor wrappers use: "torch" (default) = the TORCH_LIBRARY(rmcl, ...) operators of
    rmcl_b200_torch.so (csrc/torch_ext.cpp: C++ argument checks, ATen allocation, then the C-ABI call); "ctypes" =
    the raw C-ABI symbols from Python (RMCL_B200_FFI=ctypes; what INTEGRATION.md's stub and tests/test_capi_cpu.py use)."""
    v = os.environ.get("RMCL_B200_FFI", "torch")
    if v not in ("torch", "ctypes"):
        raise ValueError(f"RMCL_B200_FFI must be 'torch' or 'ctypes', got {v!r}")
    return v


def torch_ops():
    """torch.ops.rmcl, loading rmcl_b200_torch.so on first use.  No fallback: a missing extension is an error."""
    global _torch_ops
    if _torch_ops is None:
        import torch
        lib()                                   # librmcl_b200.so first (the extension links against it)
        if not os.path.isfile(TORCH_LIB_PATH):
            raise ImportError(
                f"rmcl_b200: {TORCH_LIB_PATH} is missing — the torch extension has not been built. "
                "Run `python __graft_entry__.py build` (or set RMCL_B200_FFI=ctypes to use the raw C-ABI binding).")
        torch.ops.load_library(TORCH_LIB_PATH)
        _torch_ops = _TorchOps(torch.ops.rmcl)
    return _torch_ops


class _TorchOps:
    """torch.ops.rmcl with the library's failures re-raised as :class:`RmclError` (a RuntimeError), so that both bindings
    report a bad shape / unsupported path the same way."""

    def __init__(self, ns):
        self._ns = ns

    def __getattr__(self, name):
        op = getattr(self._ns, name)

        def call(*args):
            try:
                return op(*args)
            except RuntimeError as e:
                msg = str(e)
                if "rmcl_" in msg and not isinstance(e, NotImplementedError):
                    raise RmclError(msg.split("\n")[0]) from None
                raise
        call.__name__ = name
        self.__dict__[name] = call
        return call


def check(rc, what):
    if rc != 0:
        msg = lib().rmcl_last_error().decode("utf-8", "replace")
        raise RmclError(f"{what} failed (status {rc}): {msg}")
