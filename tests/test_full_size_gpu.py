"""GPU parity at the sizes bench.py times (BASELINE cfg3 / cfg5) and of the code paths only those sizes reach.

* PGD update at B=128 on pixels (128x3x384x384, 226 MB) and on the patch+token embeddings (128x185x768): the
  multi-batch ticket pipeline of csrc/pgd.cu (norm of batch r || update of batch r-1, per-sample arrival counters,
  L2 evict hints) only runs when B*N*4 exceeds one 24 MB batch, i.e. from 14 pixel samples up.  All three modes,
  fp32 and bf16 delta, five consecutive steps, adversarial samples placed in later batches, against
  oracle.pgd_update (attack/pgd_attack_vilt.py:162-173): bit-exact for ref_linf / sign_linf, <= 1e-4 for l2.
* the same pipeline on an odd N (scalar path) and on a vector-path N with a ragged last chunk.
* the two-pass tcgen05 InfoNCE at cfg5's full size (512 x 768 x 262144) against the float64 oracle on a row subset.
* EmaPlan re-validation when a parameter's storage is re-bound (the reference's own ``param_k.data = ...``).
"""
import pytest
import torch

import rmcl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    import rmcl_b200
    return rmcl_b200.ops


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _adversarial_grad(shape, step, seed):
    """randn gradient (generated on the GPU: 56 M elements) with per-step scale and, on step 1 and 3, the
    adversarial samples of SURVEY 8(d) cfg3 placed far from sample 0, i.e. in later sample batches."""
    g = torch.Generator(device=DEV).manual_seed(seed * 100 + step)
    grad = torch.randn(shape, generator=g, device=DEV) * 10.0 ** (step - 3)
    B = shape[0]
    flat = grad.view(B, -1)
    if step in (1, 3) and B >= 8:
        flat[B // 2].zero_()                                  # all-zero sample: denominator clamps to 1e-8
        flat[B - 3, flat.shape[1] // 3] = 1e4                 # one huge element dominates the norm
        flat[B // 3] = 1e-41                                  # denormals
        flat[B - 1, -1] = -7e3                                # ... and one in the very last element of the tensor
    return grad


PGD_FULL = [((128, 3, 384, 384), "pixel"), ((128, 185, 768), "embed")]


@pytest.mark.parametrize("shape,space", PGD_FULL, ids=[s for _, s in PGD_FULL])
@pytest.mark.parametrize("mode", ["ref_linf", "sign_linf", "l2"])
@pytest.mark.parametrize("ddt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_pgd_cfg3_full_size_five_steps(ops, shape, space, mode, ddt):
    lr, eps = {"ref_linf": (0.05, 8.0 / 255.0), "sign_linf": (2.0 / 255.0, 8.0 / 255.0), "l2": (0.5, 1.0)}[mode]
    dd = torch.zeros(shape, dtype=ddt, device=DEV)
    ref = torch.zeros(shape, dtype=ddt if mode != "l2" else torch.float64)
    for step in range(5):
        grad = _adversarial_grad(shape, step, seed=len(shape))
        ops.pgd_step_(dd, grad, lr, eps, mode)
        gc = grad.cpu()
        if mode == "l2":
            ref = O.pgd_update(ref, gc.double(), lr, eps, mode="l2")
            if ddt == torch.float32:
                assert rel_err(dd, ref) < 1e-4, step
                assert dd.view(shape[0], -1).norm(dim=1).max().item() <= eps * (1 + 1e-5)
            else:
                assert rel_err(dd, ref) < 2e-2, step
                ref = dd.double().cpu()                       # bf16 storage: follow the rounded trajectory
        else:
            ref = O.pgd_update(ref, gc, lr, eps, mode=mode)
            got = dd.cpu()
            assert torch.equal(got, ref), f"{mode} step {step}: {(got.float() - ref.float()).abs().max().item():.3e}"
    if mode != "l2":
        bound = torch.tensor(eps, dtype=ddt).item()            # ATen casts the clamp bound to the tensor dtype
        assert 0 < dd.abs().max().item() <= bound + 1e-7


@pytest.mark.parametrize("B,N", [(64, 1_000_003), (40, 1_000_004), (30, 2_500_001)])
@pytest.mark.parametrize("mode", ["ref_linf", "l2"])
def test_pgd_multi_batch_ragged(ops, B, N, mode):
    """Several sample batches with N odd (scalar path, unaligned sample starts) and with N a multiple of 4 whose
    last chunk is partial; the last batch is short."""
    lr, eps = (0.05, 8.0 / 255.0) if mode == "ref_linf" else (0.5, 0.1)
    dd = torch.zeros(B, N, device=DEV)
    ref = torch.zeros(B, N, dtype=torch.float32 if mode == "ref_linf" else torch.float64)
    for step in range(3):
        grad = _adversarial_grad((B, N), step, seed=N % 97)
        ops.pgd_step_(dd, grad, lr, eps, mode)
        if mode == "ref_linf":
            ref = O.pgd_update(ref, grad.cpu(), lr, eps)
            assert torch.equal(dd.cpu(), ref), step
        else:
            ref = O.pgd_update(ref, grad.cpu().double(), lr, eps, mode="l2")
            assert rel_err(dd, ref) < 1e-4, step


def test_pgd_back_to_back_calls_share_the_workspace(ops):
    """Twenty launches on one stream without a sync in between: the self-resetting control words of call n must be
    back to zero before call n+1 starts (stream order), whatever the interleaving of the CTAs."""
    shape = (32, 3, 384, 384)
    grads = [_adversarial_grad(shape, s % 5, seed=9) for s in range(4)]
    dd = torch.zeros(shape, device=DEV)
    for i in range(20):
        ops.pgd_step_(dd, grads[i % 4], 0.05, 8.0 / 255.0)
    ref = torch.zeros(shape)
    gc = [g.cpu() for g in grads]
    for i in range(20):
        ref = O.pgd_update(ref, gc[i % 4], 0.05, 8.0 / 255.0)
    assert torch.equal(dd.cpu(), ref)


# ------------------------------------------------------------------------------------------- PGD, dtype pairs x shapes
MIXED_SHAPES = [(1, 4), (2, 64), (5, 6144), (300, 20000), (3, 900_000), (148, 12288), (149, 12352), (37, 3, 64, 64)]


@pytest.mark.parametrize("shape", MIXED_SHAPES, ids=[str(s) for s in MIXED_SHAPES])
@pytest.mark.parametrize("ddt", [torch.float32, torch.bfloat16], ids=["d32", "d16"])
@pytest.mark.parametrize("gdt", [torch.float32, torch.bfloat16], ids=["g32", "g16"])
def test_pgd_dtype_pairs_vs_oracle(ops, shape, ddt, gdt):
    """All four (delta, grad) dtype pairs of csrc/pgd.cu at shapes with one chunk per sample, many chunks per sample,
    ragged last chunks, fewer / one more work item than resident CTAs, N not a multiple of the 8-element bf16 vector
    (scalar path): ref_linf bit-exact against oracle.pgd_update over three steps with adversarial samples, l2 (with and
    without the projection) within 1e-4 (fp32 delta) / 2e-2 (bf16 delta) of the float64 oracle."""
    B = shape[0]
    g = torch.Generator(device=DEV).manual_seed(B)
    dd = (torch.randn(shape, generator=g, device=DEV) * 0.01).to(ddt)
    ref = dd.cpu()
    for step in range(3):
        grad = (_adversarial_grad(shape, step, seed=B)).to(gdt)
        ops.pgd_step_(dd, grad, 0.05, 8.0 / 255.0, "ref_linf")
        ref = O.pgd_update(ref, grad.cpu(), 0.05, 8.0 / 255.0)
        assert torch.equal(dd.cpu(), ref), step
    for eps in (0.1, 0.0):            # with and without the projection (three sums / one sum)
        a = (torch.randn(shape, generator=g, device=DEV) * 0.01).to(ddt)
        a.view(B, -1)[-1] *= 30.0     # starts outside the ball
        ref = a.double().cpu()
        for step in range(2):
            grad = (_adversarial_grad(shape, step + 2, seed=B + 1)).to(gdt)
            ops.pgd_step_(a, grad, 0.5, eps, "l2")
            ref = O.pgd_update(ref, grad.cpu().double(), 0.5, eps, mode="l2")
            assert rel_err(a, ref) < (1e-4 if ddt == torch.float32 else 2e-2), (eps, step)
            if ddt != torch.float32:
                ref = a.double().cpu()


@pytest.mark.parametrize("mode", ["ref_linf", "sign_linf", "l2"])
def test_pgd_nan_gradient_poisons_its_sample_only(ops, mode):
    """torch.norm(p=inf) / clamp propagate a NaN (attack/pgd_attack_vilt.py:164-165, 170): one NaN gradient element
    turns that sample's whole perturbation into NaN (ref_linf, l2) and leaves the other samples exactly as without it;
    sign_linf follows torch.sign, which maps NaN to 0."""
    shape = (6, 3, 64, 64)
    g = torch.Generator().manual_seed(5)
    grad = torch.randn(shape, generator=g)
    grad[4, 1, 7, 9] = float("nan")
    d0 = torch.randn(shape, generator=g) * 0.01
    lr, eps = (0.5, 0.1) if mode == "l2" else (0.05, 8.0 / 255.0)
    want = O.pgd_update(d0.double() if mode == "l2" else d0, grad.double() if mode == "l2" else grad, lr, eps, mode=mode)
    got = ops.pgd_step_(d0.to(DEV), grad.to(DEV), lr, eps, mode).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.isnan(got).sum().item() == (0 if mode == "sign_linf" else 3 * 64 * 64)   # torch.sign(nan) == 0
    ok = ~torch.isnan(want)
    if mode == "l2":
        assert rel_err(got[ok], want[ok]) < 1e-4
    else:
        assert torch.equal(got[ok], want[ok])


def test_pgd_modes_alternate_on_one_workspace(ops):
    """16 launches alternating between the three modes on one workspace, no sync in between: the self-resetting control
    words serve the one-phase and the two-phase plans alike."""
    shape = (40, 185, 768)
    grads = [_adversarial_grad(shape, s, seed=21) for s in range(4)]
    dd = torch.zeros(shape, device=DEV)
    plan = ["ref_linf", "ref_linf", "sign_linf", "l2", "ref_linf", "sign_linf", "l2", "ref_linf"] * 2
    for i, mode in enumerate(plan):
        ops.pgd_step_(dd, grads[i % 4], 0.02, 8.0 / 255.0 if mode != "l2" else 1.0, mode)
    ref = torch.zeros(shape, dtype=torch.float64)
    gc = [x.cpu().double() for x in grads]
    for i, mode in enumerate(plan):
        if mode == "l2":
            ref = O.pgd_update(ref, gc[i % 4], 0.02, 1.0, mode="l2")
        else:   # fp32 arithmetic of the reference rule, carried in the float64 trajectory
            ref = O.pgd_update(ref.float(), gc[i % 4].float(), 0.02, 8.0 / 255.0, mode=mode).double()
    assert rel_err(dd, ref) < 1e-4


# ------------------------------------------------------------------------------------------- InfoNCE, cfg5 full size
def _rows_oracle_bf16(q_rows, k_rows, queue_bf16, T, B_full, chunk=32768):
    """float64 InfoNCE of a few rows against the whole queue, streamed over column chunks (the 768 x 262144 queue is
    1.6 GB in float64): same operand rounding as the kernels (q^, k^ in bf16, bf16 queue), exact accumulation.
    Restates objectives.py:326-334+351 for rows of a batch of ``B_full`` (the mean's 1/B enters the gradient)."""
    qh = O.l2_normalize(q_rows.double())
    qh16 = qh.float().bfloat16().double()
    k16 = k_rows.float().bfloat16().double()
    pos = (qh16 * k16).sum(1) / T
    K = queue_bf16.shape[1]
    negs = [qh16 @ queue_bf16[:, c:c + chunk].double() / T for c in range(0, K, chunk)]
    logits = torch.cat([pos[:, None]] + negs, dim=1)
    lse = torch.logsumexp(logits, 1)
    p = torch.exp(logits - lse[:, None])
    dqh = (p[:, :1] - 1) * k16
    for i, c in enumerate(range(0, K, chunk)):
        dqh = dqh + p[:, 1 + c:1 + c + chunk] @ queue_bf16[:, c:c + chunk].double().T
    dqh = dqh / (T * B_full)
    n = q_rows.double().norm(dim=1, keepdim=True).clamp_min(1e-12)
    dq = (dqh - qh * (qh * dqh).sum(1, keepdim=True)) / n
    return {"lse": lse, "pos": pos, "loss_per_row": lse - pos, "dq": dq, "argmax": logits.argmax(1), "logits": logits}


@pytest.mark.parametrize("normalized_queue", [True, False], ids=["unit_keys", "randn_init"])
def test_infonce_cfg5_full_size_vs_oracle_rows(ops, normalized_queue):
    """BASELINE cfg5 per GPU: q [512,768], queue [768,262144] bf16 (403 MB) through the two-pass tcgen05 kernels,
    against the float64 oracle on 24 rows spread over the four row blocks."""
    B, C, K = 512, 768, 262144
    g = torch.Generator(device=DEV).manual_seed(5)
    q = torch.randn(B, C, generator=g, device=DEV)
    k = torch.nn.functional.normalize(torch.randn(B, C, generator=g, device=DEV), dim=1)
    queue = torch.randn(C, K, generator=g, device=DEV)
    if normalized_queue:
        queue = torch.nn.functional.normalize(queue, dim=0)
    queue = queue.bfloat16()
    res = ops.infonce_fwd_bwd(q, k, queue, 0.07, path="tcgen05")
    torch.cuda.synchronize()
    rows = torch.tensor([0, 1, 63, 64, 127, 128, 129, 200, 255, 256, 257, 300, 383, 384, 385, 400, 450, 500, 509, 510, 511, 31, 95, 159])
    ref = _rows_oracle_bf16(q[rows].cpu(), k[rows].cpu(), queue.cpu(), 0.07, B)
    assert rel_err(res["lse"][rows], ref["lse"]) < 1e-4
    assert (res["loss_per_row"][rows].double().cpu() - ref["loss_per_row"]).abs().max() < 2e-2 * ref["logits"].abs().max()
    assert (res["pos"][rows].double().cpu() - ref["pos"]).abs().max() < 1e-4 * ref["logits"].abs().max()
    assert rel_err(res["dq"][rows], ref["dq"]) < 1e-2
    top2 = ref["logits"].topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(res["argmax"][rows].cpu()[clear], ref["argmax"][clear])
    # whole-batch, size-independent properties: loss = mean of the row losses; dq orthogonal to q; finite everywhere
    assert torch.isfinite(res["dq"]).all() and torch.isfinite(res["lse"]).all()
    assert abs(res["loss"].item() - res["loss_per_row"].double().mean().item()) < 1e-5 * abs(res["loss"].item())
    assert ((res["dq"].double() * q.double()).sum(1).abs().max() / res["dq"].abs().max()).item() < 1e-3


# ------------------------------------------------------------------------------------------------ EMA plan upkeep
def test_ema_plan_follows_rebound_storage(ops):
    """``param.data = ...`` (what the reference's EMA line does, objectives.py:223), ``.to()``/``.half()`` or a
    flattening wrapper re-bind a parameter's storage; the cached chunk table must follow instead of updating
    the orphaned storage."""
    torch.manual_seed(0)
    pk = [torch.nn.Parameter(torch.randn(s, device=DEV), requires_grad=False) for s in [(300, 7), (1025,), (64, 64)]]
    pq = [torch.nn.Parameter(torch.randn_like(p)) for p in pk]
    plan = ops.EmaPlan(pk, pq)
    ops.ema_multi_(plan, 0.9)
    assert plan.rebuilds == 0
    k_host = [p.detach().cpu().clone() for p in pk]
    # re-bind: key tensor 1 through the reference's own update expression, query tensor 2 through a fresh copy
    pk[1].data = pk[1].data * 1.0
    pq[2].data = pq[2].data.clone()
    with torch.no_grad():
        pq[2].add_(1.0)                                       # visible only through the NEW storage
    ops.ema_multi_(plan, 0.9)
    assert plan.rebuilds == 1
    want = O.momentum_update(k_host, [p.detach().cpu() for p in pq], 0.9)
    for a, b in zip(pk, want):
        assert torch.equal(a.detach().cpu(), b)
    ops.ema_multi_(plan, 0.9)
    assert plan.rebuilds == 1                                 # unchanged storage: no rebuild


def test_shadow_layer_copies_and_freezes():
    """vilt_module.py:270-273 (_shadow_layer): key layer initialised from the query layer, frozen."""
    import rmcl_b200
    torch.manual_seed(1)
    q_layer = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4, bias=False)).to(DEV)
    k_layer = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4, bias=False)).to(DEV)
    rmcl_b200.shadow_layer(q_layer, k_layer)
    for pq, pk in zip(q_layer.parameters(), k_layer.parameters()):
        assert torch.equal(pq, pk) and pk.data_ptr() != pq.data_ptr()
        assert pq.requires_grad and not pk.requires_grad
    # ... and the EMA over the freshly shadowed pair is the identity, bit for bit (k == q  =>  k*m + q*(1-m) ~ k)
    plan = rmcl_b200.ops.EmaPlan(list(k_layer.parameters()), list(q_layer.parameters()))
    rmcl_b200.ops.ema_multi_(plan, 1.0)
    for pq, pk in zip(q_layer.parameters(), k_layer.parameters()):
        assert torch.equal(pq, pk)
