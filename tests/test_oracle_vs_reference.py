"""CPU: re-runs the UNMODIFIED reference live (when /root/reference is present — the authoring container; skipped on the
GPU box) and pins oracle/rmcl_oracle.py to it, function by function, on a fresh recording:

  vilt/modules/objectives.py:217-447   compute_moco_contrastive   (EMA 219-224, InfoNCE 326-351, enqueue 238-248)
  attack/pgd_attack_vilt.py:130-175    PGDAttack_moco.pgd_attack  (inner InfoNCE 147-158, update 162-173)

and checks that the committed golden file of the same case (tests/golden/ref_tiny_c16.npz, written by
oracle/make_golden.py) is what the reference still produces — i.e. the fixtures the GPU parity tests use are the
reference's outputs, not the oracle's.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402
import rmcl_oracle as O  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def live():
    """Two consecutive steps of the reference on the tiny stand-in module (same case as golden ref_tiny_c16)."""
    import make_golden as MG
    ref = ref_harness.load_reference()
    ref_harness.ensure_process_group()
    B, C, K, n_pgd, lr, eps = 4, 16, 64, 1, 0.05, 8.0 / 255.0     # == make_golden.py's ref_tiny_c16 case (seed 0)
    mod = MG.build_tiny_module(ref, B, C, K, hidden=32, n_pgd=n_pgd, lr=lr, eps=eps, T=0.07, m=0.999, seed=0)
    k_layers = [mod.k_text_embeddings, mod.k_token_type_embeddings, mod.k_transformer, mod.k_moco_head]
    q_layers = [mod.text_embeddings, mod.token_type_embeddings, mod.transformer, mod.moco_head]
    out, ema = {}, []
    for s in range(2):
        mod.zero_grad()
        batch = MG.tiny_batch(B, 16, 6, 0 * 100 + s)
        pkb, pq, pka, _ = MG.run_reference_step(ref, mod, batch, k_layers, q_layers, f"step{s}", out)
        ema.append((pkb, pq, pka))
        with torch.no_grad():
            for l in q_layers:
                for p in l.parameters():
                    if p.grad is not None:
                        p.add_(-0.1 * p.grad)
    return {"out": out, "ema": ema, "B": B, "C": C, "K": K, "n_pgd": n_pgd, "lr": lr, "eps": eps}


def _t(x):
    return torch.from_numpy(np.array(x))


def test_reference_still_produces_the_committed_golden(live, golden):
    g = golden("ref_tiny_c16")
    if (g.i("meta/B"), g.i("meta/C"), g.i("meta/K")) != (live["B"], live["C"], live["K"]):
        pytest.skip("golden ref_tiny_c16 was generated for another shape")
    for key, val in live["out"].items():
        if val is None or key not in g:
            continue
        a, b = np.asarray(val, dtype=np.float64), np.asarray(g.np(key), dtype=np.float64)
        assert a.shape == b.shape, key
        assert np.allclose(a, b, rtol=1e-5, atol=1e-7), key


def test_oracle_ema_is_bit_exact_with_the_reference(live):
    for pkb, pq, pka in live["ema"]:
        got = O.momentum_update(pkb, pq, 0.999)
        for a, b in zip(got, pka):
            assert torch.equal(a, b)


def test_oracle_infonce_matches_the_reference_call_sites(live):
    out = live["out"]
    for s in range(2):
        p = f"step{s}"
        k_hat, queue = _t(out[f"{p}/k_hat"]), _t(out[f"{p}/queue_before"])
        assert torch.equal(O.l2_normalize(_t(out[f"{p}/k_raw"])), k_hat)
        res = O.info_nce(_t(out[f"{p}/q_raw"]), k_hat, queue, 0.07)                       # objectives.py:326-351
        assert torch.allclose(res["logits"], _t(out[f"{p}/logits"]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(res["loss"], _t(out[f"{p}/loss"]), rtol=1e-6)
        assert torch.allclose(res["dq"], _t(out[f"{p}/dq_raw"]), rtol=1e-5, atol=1e-8)
        for a in range(live["n_pgd"]):                                                   # pgd_attack_vilt.py:147-158
            res = O.info_nce(_t(out[f"{p}/pgd{a}/q_raw"]), k_hat, queue, 0.07, loss_div=live["n_pgd"])
            assert torch.allclose(res["loss"] * live["n_pgd"], _t(out[f"{p}/pgd{a}/loss"]), rtol=1e-6)
            assert torch.allclose(res["dq"], _t(out[f"{p}/pgd{a}/dq_raw"]), rtol=1e-5, atol=1e-8)


def test_oracle_enqueue_and_pgd_update_are_bit_exact_with_the_reference(live):
    out = live["out"]
    B, K = live["B"], live["K"]
    for s in range(2):
        p = f"step{s}"
        q_after, ptr_after = O.dequeue_and_enqueue(_t(out[f"{p}/queue_before"]), int(out[f"{p}/ptr_before"]), _t(out[f"{p}/k_hat"]), K)
        ptr0 = int(out[f"{p}/ptr_before"])
        assert ptr_after == int(out[f"{p}/ptr_after"])
        assert torch.equal(q_after[:, ptr0:ptr0 + B], _t(out[f"{p}/queue_after_cols"]))
        assert q_after.double().sum().item() == pytest.approx(float(out[f"{p}/queue_after_sum64"]), rel=0, abs=1e-9)
        delta = None
        for a in range(live["n_pgd"]):                                                   # pgd_attack_vilt.py:162-173
            grad = _t(out[f"{p}/pgd{a}/grad"])
            want = _t(out[f"{p}/pgd{a}/delta_after"])
            delta = torch.zeros_like(want) if delta is None else delta
            delta = O.pgd_update(delta, grad.view_as(want), live["lr"], live["eps"])
            assert torch.equal(delta, want)
