"""GPU: the TORCH_LIBRARY(rmcl, ...) operators (csrc/torch_ext.cpp) against the raw C-ABI binding (ctypes) — the two
bindings call the same `extern "C"` entry points and must produce bit-identical results — plus what only the torch
layer provides: C++ autograd for the loss, CUDA-graph capture of a whole kernels-only step, allocator-backed workspaces.
"""
import pytest
import torch

import rmcl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def ops():
    import rmcl_b200
    return rmcl_b200.ops


def _both(monkeypatch, fn):
    out = {}
    for ffi in ("ctypes", "torch"):
        monkeypatch.setenv("RMCL_B200_FFI", ffi)
        out[ffi] = fn()
    torch.cuda.synchronize()
    return out["ctypes"], out["torch"]


@pytest.mark.parametrize("B,C,K,qdt,path", [(64, 128, 4096, torch.float32, "simt"), (128, 128, 8192, torch.bfloat16, "auto"),
                                           (96, 768, 2048, torch.bfloat16, "auto")])
def test_infonce_bindings_agree_bit_for_bit(ops, monkeypatch, B, C, K, qdt, path):
    g = torch.Generator().manual_seed(B + C)
    q, k = torch.randn(B, C, generator=g).to(DEV), torch.randn(B, C, generator=g).to(DEV)
    queue = torch.randn(C, K, generator=g).to(qdt).to(DEV)
    stats = ops.QueueStats(queue)
    a, b = _both(monkeypatch, lambda: ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, path=path))
    assert set(a) == set(b)
    for name in a:
        assert torch.equal(a[name], b[name]), name
    a, b = _both(monkeypatch, lambda: ops.infonce_fwd_bwd(q, torch.nn.functional.normalize(k, dim=1), queue, 0.07, path=path, diag=stats,
                                                          want=("loss", "dq", "argmax")))
    assert set(a) == set(b) == {"loss", "dq", "argmax", "diag"}
    for name in a:
        assert torch.equal(a[name], b[name]), name
    a, b = _both(monkeypatch, lambda: ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, need_grad=False, path=path,
                                                          want=("argmax", "k_hat", "dq")))
    assert set(a) == set(b) == {"argmax", "k_hat"}            # no dq without a gradient request, in either binding


def test_other_ops_bindings_agree(ops, monkeypatch):
    g = torch.Generator().manual_seed(3)
    pk = [torch.randn(s, generator=g) for s in [(300, 7), (1025,), (64, 64)]]
    pq = [torch.randn(s, generator=g) for s in [(300, 7), (1025,), (64, 64)]]

    def ema():
        kd, qd = [t.to(DEV) for t in pk], [t.to(DEV) for t in pq]
        ops.ema_multi_(ops.EmaPlan(kd, qd), 0.999)
        return kd
    a, b = _both(monkeypatch, ema)
    want = O.momentum_update(pk, pq, 0.999)
    for x, y, w in zip(a, b, want):
        assert torch.equal(x, y) and torch.equal(x.cpu(), w)

    grad = torch.randn(20, 3, 384, 384, generator=g).to(DEV)          # two sample batches

    def pgd(mode):
        d = torch.zeros_like(grad)
        for _ in range(3):
            ops.pgd_step_(d, grad, 0.05, 8 / 255, mode)
        return d
    for mode in ("ref_linf", "sign_linf", "l2"):
        a, b = _both(monkeypatch, lambda: pgd(mode))
        assert torch.equal(a, b), mode

    queue, keys = torch.randn(64, 1024, generator=g), torch.randn(32, 64, generator=g)

    def enq():
        qd, pd, sh = queue.to(DEV), torch.tensor([992], dtype=torch.int64, device=DEV), ops.QueueShadow()
        ops.enqueue_(qd, keys.to(DEV), pd, shadow=sh)
        ops.enqueue_(qd, keys.to(DEV) * 2, pd)
        return qd, pd, sh.tensor
    a, b = _both(monkeypatch, enq)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    wq, wp = O.dequeue_and_enqueue(queue, 992, keys, 1024)
    wq, wp = O.dequeue_and_enqueue(wq, wp, keys * 2, 1024)
    assert torch.equal(a[0].cpu(), wq) and a[1].item() == wp

    kb = torch.randn(64, 512, generator=g).to(DEV)
    qb = 0.7 * kb + 0.7 * torch.randn(64, 512, generator=g).to(DEV)
    for bpath in ("gram", "direct", "auto"):
        a, b = _both(monkeypatch, lambda: ops.barlow_fwd_bwd(qb, kb, 1 / 64, 0.0051, path=bpath))
        for name in a:
            assert torch.equal(a[name], b[name]), (bpath, name)
    from rmcl_b200._lib import RmclError
    for ffi in ("ctypes", "torch"):                     # both bindings report a library error as RmclError
        monkeypatch.setenv("RMCL_B200_FFI", ffi)
        with pytest.raises(RmclError):
            ops.enqueue_(torch.zeros(16, 30, device=DEV), torch.zeros(8, 16, device=DEV), torch.zeros(1, dtype=torch.int64, device=DEV))


def test_cxx_autograd_loss_matches_the_oracle(ops, monkeypatch):
    """torch.ops.rmcl.infonce_loss: the C++ autograd function (forward = the fused kernels, backward = dq * upstream)."""
    monkeypatch.setenv("RMCL_B200_FFI", "torch")
    g = torch.Generator().manual_seed(5)
    q, k = torch.randn(16, 128, generator=g), torch.nn.functional.normalize(torch.randn(16, 128, generator=g), dim=1)
    queue = torch.randn(128, 2048, generator=g)
    qd = q.to(DEV).requires_grad_(True)
    loss, argmax = ops.infonce_loss(qd, k.to(DEV), queue.to(DEV), 0.07, "simt")
    (loss * 3.0).backward()
    ref = O.info_nce(q, k, queue, 0.07, grad_out=3.0)
    rel = lambda a, b: ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
    assert rel(loss, ref["loss"]) < 1e-4 and rel(qd.grad, ref["dq"]) < 1e-4
    assert torch.equal(argmax.cpu(), ref["argmax"])
    # raw key in, normalised key out of the same launch
    k_raw = torch.randn(16, 128, generator=g)
    loss2, _, k_hat = ops.infonce_loss(q.to(DEV), k_raw.to(DEV), queue.to(DEV), 0.07, "simt", normalize_k=True, return_k_hat=True)
    assert rel(k_hat, O.l2_normalize(k_raw)) < 1e-6 and not k_hat.requires_grad
    with torch.inference_mode():                                 # no Autograd key: the CUDA-key kernel serves the call
        l3, _ = ops.infonce_loss(q.to(DEV), k.to(DEV), queue.to(DEV), 0.07, "simt")
    assert torch.equal(l3, loss.detach())


def test_whole_step_is_cuda_graph_capturable(ops, monkeypatch):
    """EMA -> InfoNCE -> enqueue -> PGD update through the torch operators inside one CUDA graph: outputs and
    workspaces come from the caching allocator, the queue pointer lives on the device, nothing syncs."""
    monkeypatch.setenv("RMCL_B200_FFI", "torch")
    g = torch.Generator().manual_seed(7)
    B, C, K = 128, 128, 4096
    pk = [torch.randn(s, generator=g).to(DEV) for s in [(768, 128), (3, 7)]]
    pq = [torch.randn(s, generator=g).to(DEV) for s in [(768, 128), (3, 7)]]
    plan = ops.EmaPlan(pk, pq)
    q, k = torch.randn(B, C, generator=g).to(DEV), torch.randn(B, C, generator=g).to(DEV)
    queue = torch.randn(C, K, generator=g).bfloat16().to(DEV)
    ptr = torch.zeros(1, dtype=torch.int64, device=DEV)
    grad, delta = torch.randn(4, 3, 64, 64, generator=g).to(DEV), torch.zeros(4, 3, 64, 64, device=DEV)

    def step():
        ops.ema_multi_(plan, 0.999, check_storage=False)
        r = ops.infonce_fwd_bwd(q, k, queue, 0.07, normalize_k=True, want=("loss", "dq", "k_hat"))
        ops.enqueue_(queue, r["k_hat"], ptr)
        ops.pgd_step_(delta, grad, 0.05, 8 / 255)
        return r

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()                                              # warm-up: workspaces allocated outside the capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    state = [t.clone() for t in pk] + [queue.clone(), ptr.clone(), delta.clone()]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        r = step()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    got = [t.clone() for t in pk] + [queue.clone(), ptr.clone(), delta.clone(), r["loss"].clone(), r["dq"].clone()]
    # eager replay of the same two steps from the saved state
    for t, s_ in zip(pk + [queue, ptr, delta], state):
        t.copy_(s_)
    for _ in range(2):
        r2 = step()
    torch.cuda.synchronize()
    want = pk + [queue, ptr, delta, r2["loss"], r2["dq"]]
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert ptr.item() == (2 + 2) * B % K
