"""CPU: pins oracle/rmcl_oracle.py against golden vectors produced by running the unmodified
reference (oracle/make_golden.py).  fp32 on CPU with the same ATen expressions => exact match
is expected; a few ulps are tolerated where thread count can change a reduction order."""
import numpy as np
import pytest
import torch

import rmcl_oracle as O

TINY = ["ref_tiny_c16", "ref_tiny_c128"]
ALL = TINY + ["ref_cfg1_vilt_b32"]


def _steps(g):
    return range(g.i("meta/steps"))


@pytest.mark.parametrize("name", TINY)
def test_ema_matches_reference_bit_exact(golden, name):
    g = golden(name)
    for s in _steps(g):
        n = g.i(f"step{s}/ema/n")
        kb = [g.t(f"step{s}/ema/k_before/{i}") for i in range(n)]
        q = [g.t(f"step{s}/ema/q/{i}") for i in range(n)]
        out = O.momentum_update(kb, q, g.f(f"step{s}/momentum"))
        for i in range(n):
            assert torch.equal(out[i], g.t(f"step{s}/ema/k_after/{i}")), (s, i)


def test_ema_cfg1_real_vilt(golden):
    g = golden("ref_cfg1_vilt_b32")
    assert g.i("step0/ema/n") == 161 and int(g.np("step0/ema/numels").sum()) == 111_694_848
    for i in g.np("step0/ema/kept"):
        out = O.momentum_update([g.t(f"step0/ema/k_before/{i}")], [g.t(f"step0/ema/q/{i}")], g.f("step0/momentum"))[0]
        assert torch.equal(out, g.t(f"step0/ema/k_after/{i}"))


@pytest.mark.parametrize("name", ALL)
def test_infonce_matches_reference(golden, name):
    g = golden(name)
    for s in _steps(g):
        p = f"step{s}"
        k, queue, T = g.t(f"{p}/k_hat"), g.t(f"{p}/queue_before"), g.f(f"{p}/temperature")
        r = O.info_nce(g.t(f"{p}/q_raw"), k, queue, T)
        torch.testing.assert_close(r["logits"], g.t(f"{p}/logits"), rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(r["loss"], g.t(f"{p}/loss"), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["dq"], g.t(f"{p}/dq_raw"), rtol=1e-5, atol=1e-7)
        assert torch.equal(r["argmax"], g.t(f"{p}/logits").argmax(-1))
        # PGD-inner call sites divide the loss by adv_steps (pgd_attack_vilt.py:158)
        n_pgd = g.i(f"{p}/n_pgd")
        for a in range(n_pgd):
            r = O.info_nce(g.t(f"{p}/pgd{a}/q_raw"), k, queue, T, loss_div=n_pgd)
            # the recorder sits on CrossEntropyLoss.forward, i.e. before the division
            torch.testing.assert_close(r["loss"] * n_pgd, g.t(f"{p}/pgd{a}/loss"), rtol=1e-6, atol=1e-6)
            torch.testing.assert_close(r["dq"], g.t(f"{p}/pgd{a}/dq_raw"), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(g.t(f"{p}/moco_loss"), g.t(f"{p}/loss"))  # image view only => mean of one


@pytest.mark.parametrize("name", ALL)
def test_enqueue_matches_reference_bit_exact(golden, name):
    g = golden(name)
    K, B = g.i("meta/K"), g.i("meta/B")
    for s in _steps(g):
        p = f"step{s}"
        ptr0 = g.i(f"{p}/ptr_before")
        newq, newp = O.dequeue_and_enqueue(g.t(f"{p}/queue_before"), ptr0, g.t(f"{p}/k_hat"), K, per_step_bs=B)
        assert newp == g.i(f"{p}/ptr_after") == (ptr0 + B) % K
        assert torch.equal(newq[:, ptr0:ptr0 + B], g.t(f"{p}/queue_after_cols"))
        changed = (newq != g.t(f"{p}/queue_before")).any(0).nonzero().flatten()
        assert torch.equal(changed, g.t(f"{p}/queue_changed_cols"))
        assert newq.double().sum().item() == pytest.approx(g.f(f"{p}/queue_after_sum64"), rel=0, abs=1e-9)


def test_enqueue_skip_and_overrun():
    q = torch.zeros(4, 16)
    keys = torch.ones(8, 4)
    q2, p2 = O.dequeue_and_enqueue(q, 0, keys, 16, per_step_bs=4)   # gathered batch != per_step_bs -> no-op
    assert q2 is q and p2 == 0
    q3, p3 = O.dequeue_and_enqueue(q, 8, keys, 16)
    assert p3 == 0 and q3[:, 8:].eq(1).all() and q3[:, :8].eq(0).all()
    with pytest.raises(ValueError):
        O.dequeue_and_enqueue(q, 12, keys, 16)


@pytest.mark.parametrize("name", ["ref_pgd_5step", "ref_pgd_noclamp"])
def test_pgd_update_matches_reference_bit_exact(golden, name):
    g = golden(name)
    n, lr, eps = g.i("n_pgd"), g.f("lr"), g.f("eps")
    delta = torch.zeros_like(g.t("delta_final"))
    for s in range(n):
        grad = g.t(f"pgd{s}/grad").view_as(delta)
        delta = O.pgd_update(delta, grad, lr, eps)
        if eps > 0:
            assert torch.equal(delta, g.t(f"pgd{s}/delta_after"))
    assert torch.equal(delta, g.t("delta_final"))
    if eps > 0:
        assert delta.abs().max().item() <= float(np.float32(eps))
    else:
        assert delta.abs().max().item() > 0.05  # no clamp branch (pgd_attack_vilt.py:172)


def test_pgd_update_cfg1(golden):
    g = golden("ref_cfg1_vilt_b32")
    grad = g.t("step0/pgd0/grad")
    out = O.pgd_update(torch.zeros_like(grad), grad, g.f("step0/adv_lr"), g.f("step0/adv_eps"))
    assert torch.equal(out, g.t("step0/pgd0/delta_after"))


def test_pgd_zero_gradient_and_modes():
    d = torch.full((2, 3, 4), 0.01)
    g = torch.zeros(2, 3, 4)
    g[1, 0, 0] = 2.0
    out = O.pgd_update(d, g, 0.05, 0.1)
    assert torch.equal(out[0], d[0])                      # g == 0 -> unchanged (denominator clamps to 1e-8)
    assert out[1, 0, 0].item() == pytest.approx(0.06)
    s = O.pgd_update(d, g, 0.02, 0.025, mode="sign_linf")
    assert s[1, 0, 0].item() == pytest.approx(0.025) and torch.equal(s[0], d[0])
    l2 = O.pgd_update(torch.zeros(2, 3, 4), torch.randn(2, 3, 4), 10.0, 1.0, mode="l2")
    assert torch.allclose(l2.view(2, -1).norm(dim=1), torch.ones(2), atol=1e-6)


def test_infonce_known_answers():
    # K=1 closed form
    q = torch.tensor([[3.0, 4.0]])
    k = torch.tensor([[0.6, 0.8]])
    queue = torch.tensor([[1.0], [0.0]])
    r = O.info_nce(q, k, queue, 0.5)
    assert r["loss"].item() == pytest.approx(O.closed_form_single_negative(1.0, 0.6, 0.5), rel=1e-6)
    # K identical columns: loss = log(1 + K*exp((s_neg - s_pos)/T))
    K = 37
    r = O.info_nce(q, k, queue.repeat(1, K), 0.5)
    assert r["loss"].item() == pytest.approx(np.log1p(K * np.exp((0.6 - 1.0) / 0.5)), rel=1e-6)
    assert r["argmax"].item() == 0


def test_diagnostics_restatement_matches_reference_values(golden):
    g = golden("ref_tiny_c16")
    q_hat = O.l2_normalize(g.t("step0/q_raw"))
    d = O.queue_diagnostics(q_hat, g.t("step0/k_hat"), g.t("step0/queue_before"))
    for ours, theirs in (("pos_dist", "pos_dist_attacked_img"), ("neg_dist", "neg_dist_attacked_img"),
                         ("pos_cosine", "pos_cosine_attacked_img"), ("neg_cosine", "neg_cosine_attacked_img"),
                         ("pos_dot", "pos_dot_attacked_img"), ("neg_dot", "neg_dot_attacked_img")):
        torch.testing.assert_close(d[ours], g.t(f"step0/diag/{theirs}"), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ Barlow Twins
@pytest.mark.parametrize("name", ["a", "b"])
def test_barlow_twins_matches_reference_expression_chain(golden, name):
    """oracle.barlow_twins against the reference's own lines (objectives.py:480-486 run verbatim, with the
    in-place div_/all_reduce/add_/pow_ and autograd backward, by oracle/make_golden.py:make_barlow)."""
    g = golden("ref_barlow_vectors")
    q, k = g.t(f"{name}/q"), g.t(f"{name}/k")
    r = O.barlow_twins([q], [k], q.shape[0], 0.0051)
    for key in ("on_diag", "off_diag", "loss"):
        assert abs(float(r[key]) - g.f(f"{name}/{key}")) <= 2e-6 * abs(g.f(f"{name}/{key}")), key
    want = g.t(f"{name}/dq")
    assert (r["dq"][0] - want).abs().max().item() <= 2e-6 * want.abs().max().item()


def test_barlow_twins_facade_golden_is_consistent(golden):
    """The whole-step golden's returned sums obey the reference's own bookkeeping (objectives.py:486,555):
    barlowtwins_loss = (invariance + redundancy) / 1 view, total = sum of the three "*loss*" keys."""
    g = golden("ref_barlow_facade")
    for s in range(g.i("meta/steps")):
        inv, red = g.f(f"step{s}/ret/barlowtwins_loss_invariance_img"), g.f(f"step{s}/ret/barlowtwins_loss_redundancy_img")
        assert g.f(f"step{s}/ret/barlowtwins_loss") == pytest.approx(inv + red, rel=1e-6)
        assert g.f(f"step{s}/total_loss") == pytest.approx(2 * (inv + red), rel=1e-6)


def test_barlow_twins_rank_split_equals_gathered_batch():
    """sum_r q_r.T k_r (the reference's all-reduce of c) == (gathered q).T (gathered k): the identity the CUDA
    path relies on to replace the 268 MB all-reduce by an all-gather of the projections."""
    g = torch.Generator().manual_seed(0)
    qs = [torch.randn(6, 40, generator=g, dtype=torch.float64) for _ in range(3)]
    ks = [torch.randn(6, 40, generator=g, dtype=torch.float64) for _ in range(3)]
    a = O.barlow_twins(qs, ks, 18, 0.0051)
    b = O.barlow_twins([torch.cat(qs)], [torch.cat(ks)], 18, 0.0051)
    assert torch.allclose(a["c"], b["c"], rtol=1e-12, atol=1e-14)
    assert torch.allclose(torch.cat(a["dq"]), b["dq"][0], rtol=1e-12, atol=1e-14)
