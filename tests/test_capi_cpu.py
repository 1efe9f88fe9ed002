"""CPU: the C-ABI library loads, exports every symbol include/rmcl_b200.h declares, validates its
arguments, and the product refuses to run without a GPU (no silent fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import rmcl_b200
    from rmcl_b200 import _lib
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rmcl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rmcl_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(L):
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rmcl_b200.h but not exported"
    from rmcl_b200 import _lib
    assert set(_lib.EXPORTS) == set(names)


def test_library_is_in_tree_and_has_sm100a_code(L):
    from rmcl_b200 import _lib
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)
    blob = open(_lib.LIB_PATH, "rb").read()
    assert b"sm_100a" in blob


def test_ema_plan_host_logic(L):
    from rmcl_b200 import _lib
    n = 3
    kp = (C.c_void_p * n)(0x1000, 0x200000, 0x400000)
    qp = (C.c_void_p * n)(0x100000, 0x300000, 0x500000)
    ne = (C.c_uint64 * n)(10, 16384 * 2 + 5, 0)
    cnt = L.rmcl_ema_plan(kp, qp, ne, n, _lib.RMCL_F32, 16384, None)
    assert cnt == 1 + 3 + 0
    out = (_lib.EmaChunk * cnt)()
    assert L.rmcl_ema_plan(kp, qp, ne, n, _lib.RMCL_F32, 16384, out) == cnt
    assert (out[0].k, out[0].q, out[0].n) == (0x1000, 0x100000, 10)
    assert (out[1].k, out[1].n) == (0x200000, 16384)
    assert (out[2].k, out[2].q, out[2].n) == (0x200000 + 16384 * 4, 0x300000 + 16384 * 4, 16384)
    assert (out[3].k, out[3].n) == (0x200000 + 2 * 16384 * 4, 5)
    # bf16 halves the byte stride
    assert L.rmcl_ema_plan(kp, qp, ne, n, _lib.RMCL_BF16, 16384, out) == cnt
    assert out[2].k == 0x200000 + 16384 * 2
    assert L.rmcl_ema_plan(None, qp, ne, n, _lib.RMCL_F32, 16384, None) == -1
    assert b"rmcl_ema_plan" in L.rmcl_last_error()


def test_argument_validation_without_gpu(L):
    from rmcl_b200 import _lib
    one = C.c_void_p(0x1000)
    # K % B != 0  (the reference's commented-out assert, objectives.py:245)
    assert L.rmcl_enqueue(one, 0, one, 0, one, 8, 16, 30, 30, None) == -1
    assert b"multiple" in L.rmcl_last_error()
    assert L.rmcl_enqueue(None, 0, one, 0, one, 8, 16, 32, 32, None) == -1
    assert L.rmcl_pgd_step(one, 0, one, 0, 0, 10, 0.1, 0.1, 0, None, 0, None) == -1
    assert L.rmcl_pgd_step(one, 0, one, 0, 2, 10, 0.1, 0.1, 7, None, 0, None) == -1
    assert L.rmcl_pgd_workspace_bytes(4, 1000, 0) >= (2 + 2 * 4) * 4 + 2 * (4 + 4) * 4
    assert L.rmcl_pgd_workspace_bytes(0, 1000, 0) == 0
    assert L.rmcl_ema_multi(None, 5, 0.999, 0, None) == -1
    assert L.rmcl_ema_multi(None, 0, 0.999, 0, None) == 0          # empty list is a no-op
    assert L.rmcl_infonce_fwd_bwd(one, 0, one, 0, one, 0, 4, 8, 16, 16, 0.0, 1.0, 0, 0, None, None, None, None, None,
                                  None, None, None, one, 0, None) == -1  # tau must be > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import rmcl_b200
    from rmcl_b200._lib import RmclError
    with pytest.raises(RuntimeError):
        rmcl_b200.ops.enqueue_(torch.zeros(4, 8), torch.zeros(2, 4), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError):
        rmcl_b200.ops.infonce_fwd_bwd(torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(4, 8), 0.07)
    L = rmcl_b200._lib.lib()
    assert L.rmcl_sm_count() < 0 and len(L.rmcl_last_error()) > 0
    # a launch attempt without a device reports RMCL_E_CUDA, it does not compute anything
    buf = (C.c_float * 64)()
    ws = (C.c_char * 4096)()
    aligned = (C.addressof(ws) + 255) // 256 * 256
    rc = L.rmcl_pgd_step(buf, 0, buf, 0, 2, 32, 0.1, 0.1, 0, C.c_void_p(aligned), 3072, None)
    assert rc == -4
    assert RmclError is not None


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "robust-multimodal-contrastive-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "rmcl_oracle" not in src and "ref_harness" not in src, f


def _declared_prototypes():
    """name -> number of parameters, parsed from the (comment-stripped) header."""
    text = open(os.path.join(ROOT, "include", "rmcl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(rmcl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_ctypes_signatures_match_the_header(L):
    """ABI drift guard: every binding in rmcl_b200/_lib.py passes exactly as many arguments as include/rmcl_b200.h declares."""
    protos = _declared_prototypes()
    checked = 0
    for name, n_params in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        assert len(fn.argtypes) == n_params, f"{name}: header declares {n_params} parameters, binding passes {len(fn.argtypes)}"
        checked += 1
    assert checked >= 12


def test_new_entry_points_validate_arguments_without_gpu(L):
    one = C.c_void_p(0x1000)
    # Barlow Twins: local rows must lie inside the gathered batch; path must be known
    assert L.rmcl_barlow_fwd_bwd(one, 0, one, 0, 8, 64, 4, 8, 0.125, 0.0051, 1.0, 0.0051, 1.0, 0, None, None, None, None, None,
                                 one, 0, None) == -1
    assert b"bad sizes" in L.rmcl_last_error()
    assert L.rmcl_barlow_fwd_bwd(one, 0, one, 0, 8, 64, 0, 8, 0.125, 0.0051, 1.0, 0.0051, 1.0, 9, None, None, None, None, None,
                                 one, 0, None) == -1
    assert b"bad path" in L.rmcl_last_error()
    # peer-memory exchange: rank inside the world, queue length a multiple of the gathered batch
    assert L.rmcl_gather_enqueue_p2p(one, one, one, one, 0, None, 0, 1, one, 2, 2, 8, 16, 64, 64, None) == -1
    assert L.rmcl_gather_enqueue_p2p(one, one, one, one, 0, None, 0, 1, one, 0, 2, 8, 16, 40, 40, None) == -1
    assert b"multiple" in L.rmcl_last_error()
    assert L.rmcl_gather_enqueue_p2p(one, one, one, one, 0, one, 64, 3, one, 0, 2, 8, 16, 64, 64, None) == -1
    assert b"shadow_planes" in L.rmcl_last_error()
    # hi/lo queue planes: sizes and alignment are checked before anything is launched
    assert L.rmcl_queue_split(one, 64, 100, 100, one, 100, None) == -1            # K % 8 != 0
    assert b"rmcl_queue_split" in L.rmcl_last_error()
    assert L.rmcl_enqueue_shadow(one, 0, one, 64, 3, one, 0, one, 8, 16, 64, 64, None) == -1
    assert b"shadow_planes" in L.rmcl_last_error()


def test_torch_extension_loads_and_registers_the_operators():
    """rmcl_b200_torch.so (csrc/torch_ext.cpp): TORCH_LIBRARY(rmcl, ...) over the C-ABI.  On a CPU-only host the
    operators exist, reject CPU tensors (there is no CPU kernel behind any of them) and nothing computes."""
    import torch
    from rmcl_b200 import _lib
    tx = _lib.torch_ops()
    for name in ("ema_multi_", "infonce_fwd_bwd", "infonce_loss", "queue_stats", "enqueue_", "pgd_step_", "barlow_fwd_bwd"):
        assert hasattr(tx, name), name
    with pytest.raises((RuntimeError, NotImplementedError)):
        tx.pgd_step_(torch.zeros(2, 4), torch.zeros(2, 4), 0.1, 0.1, 0)
    with pytest.raises((RuntimeError, NotImplementedError)):
        tx.infonce_fwd_bwd(torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(4, 8), 0.07, 1.0, False, True, 0, None, None, None,
                           1e-6, 255, False)
    # the schema is part of the boundary: in-place ops are declared as such
    assert "Tensor(a!) queue" in str(torch.ops.rmcl.enqueue_.default._schema)
    assert "Tensor(a!) delta" in str(torch.ops.rmcl.pgd_step_.default._schema)
