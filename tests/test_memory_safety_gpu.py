"""Memory-safety and race evidence for the kernels, through the raw C-ABI (race detection).

compute-sanitizer is closed on this GPU pool (the tool says so itself), so the two failure classes it would catch are
tested directly:

* out-of-bounds WRITES — every buffer an entry point is handed (inputs, outputs, the caller-owned workspace) is carved out
  of an arena pre-filled with a poison byte, with poisoned guard zones on both sides; after the call every byte outside
  the carved regions must still be poison, inputs must be unchanged, and workspaces are exactly
  ``rmcl_*_workspace_bytes`` long (a kernel that needs more than it declares writes into a guard zone).
* data races / missing synchronisation — every kernel is deterministic by construction (fixed-order merges, no float
  atomics on results), so a race shows as run-to-run differences: each op is run 12 times on the same inputs, alternating
  between two streams, and all results must be bit-identical.
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
POISON = 0xA5
GUARD = 8192


@pytest.fixture(scope="module")
def L():
    from rmcl_b200 import _lib
    return _lib.lib()


class Arena:
    """Poisoned device arena; ``take`` carves 256-byte aligned regions separated by GUARD poisoned bytes."""

    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), POISON, dtype=torch.uint8, device=DEV)
        self.off = GUARD
        self.regions = []

    def take(self, shape, dtype, init=None):
        n = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        base = self.buf.data_ptr()
        start = self.off + GUARD
        start += (-(base + start)) % 256
        assert start + n + GUARD <= self.buf.numel(), "arena too small"
        view = self.buf[start:start + n].view(dtype).view(shape)
        self.regions.append((start, n))
        self.off = start + n
        if init is not None:
            view.copy_(init.to(dtype).view(shape))
        return view

    def take_bytes(self, n, zero=True):
        v = self.take((n,), torch.uint8)
        if zero:
            v.zero_()
        return v

    def check(self):
        torch.cuda.synchronize()
        mask = torch.ones(self.buf.numel(), dtype=torch.bool, device=DEV)
        for start, n in self.regions:
            mask[start:start + n] = False
        bad = (self.buf != POISON) & mask
        assert not bool(bad.any()), f"write outside the buffers at arena offset {int(bad.nonzero()[0])}"


def P(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


F32, BF16, HILO = 0, 1, 2


def _dt(t):
    return BF16 if t.dtype == torch.bfloat16 else F32


# ------------------------------------------------------------------------------------------------ InfoNCE
INFONCE_CASES = [
    # B, C, K, q/k dtype, queue dtype, path (0 auto / 1 simt / 2 tcgen05), flags
    (130, 256, 4104, torch.bfloat16, torch.bfloat16, 2, 1),   # fused tcgen05, ragged rows and last tile
    (256, 128, 8192, torch.bfloat16, torch.bfloat16, 2, 0),
    (70, 64, 1032, torch.float32, torch.bfloat16, 2, 1),
    (130, 768, 2056, torch.bfloat16, torch.bfloat16, 2, 1),   # two-pass tcgen05
    (33, 100, 1000, torch.float32, torch.float32, 1, 1),      # CUDA cores, odd everything
    (130, 128, 4104, torch.float32, "hilo", 2, 1),            # fp32-accurate split-operand path
]


def _infonce_call(L, B, Cd, K, qdt, kind, path, flags, gen):
    a = Arena(64 << 20)
    q = a.take((B, Cd), qdt, torch.randn(B, Cd, generator=gen, device=DEV))
    k = a.take((B, Cd), qdt, torch.randn(B, Cd, generator=gen, device=DEV))
    if kind == "hilo":
        qf = a.take((Cd, K), torch.float32, torch.randn(Cd, K, generator=gen, device=DEV))
        queue = a.take((2 * Cd, K), torch.bfloat16)
        assert L.rmcl_queue_split(P(qf), Cd, K, K, P(queue), K, _stream()) == 0, L.rmcl_last_error()
        qd = HILO
    else:
        queue = a.take((Cd, K), kind, torch.randn(Cd, K, generator=gen, device=DEV))
        qd = _dt(queue)
    outs = {n: a.take(s, d) for n, s, d in (("loss", (1,), torch.float32), ("row", (B,), torch.float32), ("lse", (B,), torch.float32),
                                            ("pos", (B,), torch.float32), ("argmax", (B,), torch.int64), ("dq", (B, Cd), torch.float32),
                                            ("dk", (B, Cd), torch.float32), ("khat", (B, Cd), torch.float32))}
    nws = L.rmcl_infonce_workspace_bytes(B, Cd, K, qd, path)
    assert nws > 0, L.rmcl_last_error()
    ws = a.take_bytes(nws)
    snap = (q.clone(), k.clone(), queue.clone())
    rc = L.rmcl_infonce_fwd_bwd(P(q), _dt(q), P(k), _dt(k), P(queue), qd, B, Cd, K, K, 0.07, 1.0, flags, path, P(outs["loss"]),
                                P(outs["row"]), P(outs["lse"]), P(outs["pos"]), P(outs["argmax"]), P(outs["dq"]), P(outs["dk"]),
                                P(outs["khat"]), P(ws), nws, _stream())
    assert rc == 0, L.rmcl_last_error()
    a.check()
    for before, after in zip(snap, (q, k, queue)):
        assert torch.equal(before, after), "an input was modified"
    assert all(bool(torch.isfinite(outs[n].float()).all()) for n in ("loss", "row", "lse", "pos", "dq", "dk", "khat"))
    assert int(outs["argmax"].min()) >= 0 and int(outs["argmax"].max()) <= K
    return {n: v.clone() for n, v in outs.items()}


@pytest.mark.parametrize("B,Cd,K,qdt,kind,path,flags", INFONCE_CASES, ids=[f"B{c[0]}C{c[1]}K{c[2]}-{c[4] if isinstance(c[4], str) else str(c[4])[6:]}" for c in INFONCE_CASES])
def test_infonce_stays_inside_its_buffers_and_is_deterministic(L, B, Cd, K, qdt, kind, path, flags):
    ref = _infonce_call(L, B, Cd, K, qdt, kind, path, flags, torch.Generator(device=DEV).manual_seed(1))
    side = torch.cuda.Stream()
    for rep in range(11):
        if rep % 2:
            with torch.cuda.stream(side):
                out = _infonce_call(L, B, Cd, K, qdt, kind, path, flags, torch.Generator(device=DEV).manual_seed(1))
        else:
            out = _infonce_call(L, B, Cd, K, qdt, kind, path, flags, torch.Generator(device=DEV).manual_seed(1))
        for n in ref:
            assert torch.equal(ref[n], out[n]), f"run {rep}: {n} differs between identical runs"


# ------------------------------------------------------------------------------------------------ PGD
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["ref_linf", "sign_linf", "l2"])
@pytest.mark.parametrize("ddt,gdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.float32, torch.bfloat16)],
                         ids=["f32", "d16", "g16"])
@pytest.mark.parametrize("B,N", [(37, 3 * 64 * 64), (40, 185 * 768), (5, 1001), (150, 12352)])
def test_pgd_stays_inside_its_buffers_and_is_deterministic(L, mode, ddt, gdt, B, N):
    def run():
        gen = torch.Generator(device=DEV).manual_seed(B + mode)
        a = Arena(96 << 20)
        delta = a.take((B, N), ddt, torch.randn(B, N, generator=gen, device=DEV) * 0.01)
        grad = a.take((B, N), gdt, torch.randn(B, N, generator=gen, device=DEV))
        nws = L.rmcl_pgd_workspace_bytes(B, N, _dt(grad))
        assert nws > 0
        ws = a.take_bytes(nws)
        g0 = grad.clone()
        for _ in range(3):      # consecutive calls on one workspace: the control words must come back to zero in between
            rc = L.rmcl_pgd_step(P(delta), _dt(delta), P(grad), _dt(grad), B, N, 0.05, 8.0 / 255.0 if mode != 2 else 0.1, mode,
                                 P(ws), nws, _stream())
            assert rc == 0, L.rmcl_last_error()
        a.check()
        assert torch.equal(g0, grad), "the gradient was modified"
        assert bool(torch.isfinite(delta.float()).all())
        return delta.clone()

    ref = run()
    side = torch.cuda.Stream()
    for rep in range(5):
        if rep % 2:
            with torch.cuda.stream(side):
                out = run()
        else:
            out = run()
        assert torch.equal(ref, out), f"run {rep} differs between identical runs"


# ------------------------------------------------------------------------------------------------ enqueue, statistics, EMA
@pytest.mark.parametrize("qdt,kdt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)],
                         ids=["f32", "k16", "q16"])
def test_enqueue_and_queue_statistics_stay_inside_their_buffers(L, qdt, kdt):
    B, Cd, K = 48, 128, 48 * 7
    gen = torch.Generator(device=DEV).manual_seed(3)
    a = Arena(16 << 20)
    queue = a.take((Cd, K), qdt, torch.randn(Cd, K, generator=gen, device=DEV))
    shadow = a.take((2 * Cd, K), torch.bfloat16)
    shadow.zero_()
    keys = a.take((B, Cd), kdt, torch.randn(B, Cd, generator=gen, device=DEV))
    ptr = a.take((1,), torch.int64)
    ptr.fill_(K - 2 * B)
    want = queue.clone()
    for step in range(3):       # the third step wraps to column 0
        p0 = int(ptr.item())
        want[:, p0:p0 + B] = keys.t().to(qdt)
        if qdt == torch.float32:
            rc = L.rmcl_enqueue_shadow(P(queue), _dt(queue), P(shadow), K, 2, P(keys), _dt(keys), P(ptr), B, Cd, K, K, _stream())
        else:
            rc = L.rmcl_enqueue(P(queue), _dt(queue), P(keys), _dt(keys), P(ptr), B, Cd, K, K, _stream())
        assert rc == 0, L.rmcl_last_error()
        assert int(ptr.item()) == (p0 + B) % K
    a.check()
    assert torch.equal(queue, want)
    col = a.take((K,), torch.float32)
    sv = a.take((Cd,), torch.float32)
    su = a.take((Cd,), torch.float32)
    assert L.rmcl_queue_stats(P(queue), _dt(queue), Cd, K, K, 1e-6, P(col), P(sv), P(su), _stream()) == 0, L.rmcl_last_error()
    a.check()
    assert torch.equal(queue, want)
    torch.testing.assert_close(col, (queue.float() ** 2).sum(0), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_ema_stays_inside_its_buffers(L, dtype):
    from rmcl_b200 import ops
    shapes = [(1,), (7, 5), (768,), (40, 768), (3072, 33), (16385,), (3,)]
    gen = torch.Generator(device=DEV).manual_seed(5)
    a = Arena(8 << 20)
    ks = [a.take(s, dtype, torch.randn(s, generator=gen, device=DEV)) for s in shapes]
    qs = [a.take(s, dtype, torch.randn(s, generator=gen, device=DEV)) for s in shapes]
    q0 = [q.clone() for q in qs]
    k0 = [k.clone() for k in ks]
    plan = ops.EmaPlan(ks, qs)
    for _ in range(2):
        ops.ema_multi_(plan, 0.99)
    a.check()
    for q, before in zip(qs, q0):
        assert torch.equal(q, before), "a query parameter was modified"
    for k, before in zip(ks, k0):
        assert bool(torch.isfinite(k.float()).all()) and not torch.equal(k, before)


# ------------------------------------------------------------------------------------------------ Barlow-Twins
@pytest.mark.parametrize("Bg,D,path,b0,Bl", [(96, 1024, 2, 0, 96), (128, 8192, 2, 32, 64), (200, 512, 1, 0, 200), (128, 1024, 1, 64, 64)],
                         ids=["gram96", "gram128x8192-rows32..96", "direct200", "direct128-rows64..128"])
def test_barlow_stays_inside_its_buffers_and_is_deterministic(L, Bg, D, path, b0, Bl):
    def run():
        gen = torch.Generator(device=DEV).manual_seed(Bg + D)
        a = Arena(96 << 20)
        k = a.take((Bg, D), torch.float32, torch.randn(Bg, D, generator=gen, device=DEV))
        q = a.take((Bg, D), torch.float32, 0.7 * k + 0.7 * torch.randn(Bg, D, generator=gen, device=DEV))
        on, off, loss = (a.take((1,), torch.float32) for _ in range(3))
        dq = a.take((Bl, D), torch.float32)
        cdiag = a.take((D,), torch.float32)
        nws = L.rmcl_barlow_workspace_bytes(Bg, D)
        assert nws > 0, L.rmcl_last_error()
        ws = a.take_bytes(nws, zero=False)      # this workspace needs no initialisation: poison is as good as anything
        snap = (q.clone(), k.clone())
        rc = L.rmcl_barlow_fwd_bwd(P(q), F32, P(k), F32, Bg, D, b0, Bl, 1.0 / Bg, 0.0051, 1.0, 0.0051, 1.0, path, P(on), P(off),
                                   P(loss), P(dq), P(cdiag), P(ws), nws, _stream())
        assert rc == 0, L.rmcl_last_error()
        a.check()
        assert torch.equal(snap[0], q) and torch.equal(snap[1], k), "an input was modified"
        out = torch.cat([on, off, loss, dq.flatten(), cdiag]).clone()
        assert bool(torch.isfinite(out).all())
        return out

    ref = run()
    side = torch.cuda.Stream()
    for rep in range(5):
        if rep % 2:
            with torch.cuda.stream(side):
                out = run()
        else:
            out = run()
        assert torch.equal(ref, out), f"run {rep} differs between identical runs"
