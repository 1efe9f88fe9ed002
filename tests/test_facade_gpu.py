"""GPU: the drop-in facade (rmcl_b200.compute_moco_contrastive / compute_pgd / PGDAttack_moco) run
end to end on a small stand-in LightningModule, against what the UNMODIFIED reference produced for
the same module state and batches (tests/golden/ref_facade_*.npz, written by
oracle/make_golden.py:make_facade from vilt/modules/objectives.py:217-447 +
attack/pgd_attack_vilt.py:130-175 on CPU).

The module below re-creates the stand-in of oracle/make_golden.py (toy encoder, the reference's
MOCOHead layout heads.py:129-143) so that the golden state_dict loads 1:1.  Backbone forwards run
in torch on the GPU, everything between them in the CUDA kernels (fp32 parity path)."""
from copy import deepcopy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


class ToyBlock(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.qkv, self.proj = nn.Linear(h, 3 * h), nn.Linear(h, h)
        self.fc1, self.fc2 = nn.Linear(h, 2 * h), nn.Linear(2 * h, h)

    def forward(self, x, mask=None):
        q, k, v = self.qkv(self.n1(x)).chunk(3, dim=-1)
        a = torch.softmax(q @ k.transpose(1, 2) / q.shape[-1] ** 0.5, dim=-1)
        x = x + self.proj(a @ v)
        x = x + self.fc2(torch.nn.functional.gelu(self.fc1(self.n2(x))))
        return x, a


class ToyTransformer(nn.Module):
    def __init__(self, h, patch):
        super().__init__()
        self.patch_embed = nn.Conv2d(3, h, patch, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, h))
        self.blocks = nn.ModuleList([ToyBlock(h), ToyBlock(h)])
        self.norm = nn.LayerNorm(h)

    def visual_embed(self, img, max_image_len=200, mask_it=False):
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        return x, torch.ones(x.shape[0], x.shape[1], dtype=torch.long, device=x.device), None, None


class ToyPooler(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.dense = nn.Linear(h, h)

    def forward(self, x):
        return torch.tanh(self.dense(x[:, 0]))


class MOCOHead(nn.Module):  # parameter layout of vilt/modules/heads.py:129-143
    def __init__(self, i, h, o):
        super().__init__()
        self.projector = nn.Sequential(nn.Linear(i, h), nn.LayerNorm(h), nn.ReLU(), nn.Linear(h, o, bias=False))

    def forward(self, x):
        return self.projector(x)


class TinyModule(nn.Module):
    """Duck-typed stand-in for ViLTransformerSS: exactly the attributes the objective reads
    (vilt_module.py:69-107)."""

    def __init__(self, g, infonce_path):
        super().__init__()
        import rmcl_b200
        h, C, K = g.i("meta/hidden"), g.i("meta/C"), g.i("meta/K")
        self.text_embeddings = nn.Embedding(50, h)
        self.token_type_embeddings = nn.Embedding(2, h)
        self.transformer = ToyTransformer(h, 8)
        self.pooler = ToyPooler(h)
        self.moco_head = MOCOHead(h, h, C)
        self.k_text_embeddings = deepcopy(self.text_embeddings)
        self.k_token_type_embeddings = deepcopy(self.token_type_embeddings)
        self.k_transformer = deepcopy(self.transformer)
        self.k_moco_head = deepcopy(self.moco_head)
        for l in (self.k_text_embeddings, self.k_token_type_embeddings, self.k_transformer, self.k_moco_head):
            for p in l.parameters():
                p.requires_grad = False
        self.momentum, self.temperature = g.f("meta/m"), g.f("meta/T")
        self.text_view, self.image_view, self.augmentation = False, True, False
        self.num_negative, self.per_step_bs = K, g.i("meta/B")
        self.cosine = nn.CosineSimilarity(dim=1, eps=1e-6)
        self.register_buffer("proj_queue", torch.zeros(C, K))
        self.register_buffer("proj_queue_ptr", torch.zeros(1, dtype=torch.long))
        cfg = dict(adv_steps_img=g.i("meta/n_pgd"), adv_lr_img=g.f("meta/lr"), adv_max_norm_img=g.f("meta/eps"),
                   max_image_len=200)
        self.pgd_attacker = rmcl_b200.PGDAttack_moco(cfg, infonce_path=infonce_path)
        self.infonce_path = infonce_path
        self.max_image_len = 200
        self.train_moco_loss = lambda x: x
        self.val_moco_loss = lambda x: x
        self.logged = {}
        self.load_state_dict({k[len("state/"):]: g.t(k) for k in g.z.files if k.startswith("state/")})

    def log(self, name, value, **kw):
        self.logged[name] = value

    def infer(self, batch, mask_text=False, mask_image=False, **kw):
        import rmcl_b200
        return rmcl_b200.PGDAttack.infer(self, batch, mask_text, mask_image, **kw)

    def infer_k(self, batch, mask_text=False, mask_image=False):
        import types
        import rmcl_b200
        view = types.SimpleNamespace(text_embeddings=self.k_text_embeddings,
                                     token_type_embeddings=self.k_token_type_embeddings,
                                     transformer=self.k_transformer, pooler=self.pooler, max_image_len=200)
        return rmcl_b200.PGDAttack.infer(view, batch, mask_text, mask_image)


def _batch(g, s):
    img = g.t(f"step{s}/batch/image").to(DEV)
    ids = g.t(f"step{s}/batch/text_ids").to(DEV)
    return {"image": [img], "text": ["x"] * img.shape[0], "text_ids": ids,
            "text_labels": torch.full_like(ids, -100), "text_masks": torch.ones_like(ids)}


def _close(a, b, rtol, what):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= rtol * max(ref, 1e-12), f"{what}: max err {err:.3e} vs scale {ref:.3e}"


@pytest.mark.parametrize("name,path", [("ref_facade_c128", "simt"), ("ref_facade_c16", "simt"), ("ref_facade_c128", "auto")])
def test_compute_moco_contrastive_matches_reference(golden, name, path):
    """path="simt": exact fp32 products on the CUDA cores; path="auto" on the fp32 queue with C = 128: the fp32-accurate
    split-operand tcgen05 kernels for the main-step AND the PGD-inner InfoNCE (the default of the drop-in) — same bars."""
    import rmcl_b200
    g = golden(name)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    mod = TinyModule(g, path).to(DEV).train()
    for s in range(g.i("meta/steps")):
        mod.zero_grad()
        mod.logged.clear()
        ret = rmcl_b200.compute_moco_contrastive(mod, _batch(g, s))
        assert [k for k in ret if "loss" in k] == ["moco_loss"]            # vilt_module.py:475 sums "*loss*" keys
        ret["moco_loss"].backward()
        torch.cuda.synchronize()
        _close(ret["moco_loss"], g.np(f"step{s}/ret/moco_loss"), 1e-4, f"step {s} moco_loss")
        for k in g.z.files:
            if k.startswith(f"step{s}/ret/") and not k.endswith("moco_loss"):
                _close(ret[k.split("/")[-1]], g.np(k), 2e-4, k)
        # queue: the enqueued keys come out of a GPU forward of the key encoder -> fp32 noise only
        assert mod.proj_queue_ptr.item() == g.i(f"step{s}/ptr_after")
        _close(mod.proj_queue, g.np(f"step{s}/queue_after"), 1e-5, f"step {s} queue")
        # EMA'd key parameters: elementwise fp32 arithmetic on identical inputs -> bit-exact at step 0
        for k, v in mod.named_parameters():
            if k.startswith("k_"):
                want = g.t(f"step{s}/k_after/{k}")
                if s == 0:
                    assert torch.equal(v.detach().cpu(), want), k
                else:   # step 1 runs on parameters nudged by step 0's gradients (fp32-level differences between the paths)
                    _close(v, want, 1e-5 if path == "simt" else 1e-4, k)
            elif f"step{s}/grad/{k}" in g:
                _close(v.grad, g.np(f"step{s}/grad/{k}"), 5e-3, f"step {s} grad {k}")
        rate = mod.logged["moco_attack/PGD_success_rate"]
        assert float(rate) == pytest.approx(g.f(f"step{s}/log/moco_attack/PGD_success_rate"), abs=1e-6)
        _close(mod.logged["moco_attack/train/delta"], g.np(f"step{s}/log/moco_attack/train/delta"), 1e-3, "delta norm")
        with torch.no_grad():   # the generator's optimiser-like nudge between the steps
            for k, v in mod.named_parameters():
                if v.grad is not None:
                    v.add_(-0.1 * v.grad)


def test_facade_bf16_queue_takes_tcgen05_path(golden):
    """Same step with a bf16 queue (tcgen05 kernel): loss within the bf16 tolerance of the reference."""
    import rmcl_b200
    g = golden("ref_facade_c128")
    mod = TinyModule(g, "auto").to(DEV).train()
    mod.proj_queue = mod.proj_queue.bfloat16()
    ret = rmcl_b200.compute_moco_contrastive(mod, _batch(g, 0))
    ret["moco_loss"].backward()
    torch.cuda.synchronize()
    _close(ret["moco_loss"], g.np("step0/ret/moco_loss"), 2e-2, "bf16 moco_loss")
    assert mod.proj_queue_ptr.item() == g.i("step0/ptr_after")
    assert mod.moco_head.projector[0].weight.grad is not None


def test_facade_autocast_selects_the_bf16_shadow(golden):
    """Under autocast (Lightning precision=16, the reference's training mode) the fp32 queue buffer stays the checkpointed
    state but the main-step InfoNCE reads its bf16 shadow (tcgen05 path); every enqueue keeps the shadow current, also
    when a later call runs in full precision."""
    import rmcl_b200
    g = golden("ref_facade_c128")
    mod = TinyModule(g, "auto").to(DEV).train()
    assert mod.proj_queue.dtype == torch.float32 and "_rmcl_queue_shadow" not in mod.__dict__
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ret = rmcl_b200.compute_moco_contrastive(mod, _batch(g, 0))
    ret["moco_loss"].backward()
    sh = mod.__dict__["_rmcl_queue_shadow"]
    assert torch.equal(sh.get(mod.proj_queue), mod.proj_queue.bfloat16())            # shadow == bf16(queue) after the enqueue
    _close(ret["moco_loss"], g.np("step0/ret/moco_loss"), 5e-2, "autocast moco_loss")  # bf16 backbone + bf16 InfoNCE operands
    assert mod.proj_queue_ptr.item() == g.i("step0/ptr_after")
    mod.zero_grad()
    rmcl_b200.compute_moco_contrastive(mod, _batch(g, 1))["moco_loss"].backward()      # full precision: fp32 buffer, fp32 kernel
    assert torch.equal(sh.get(mod.proj_queue), mod.proj_queue.bfloat16())            # ... and the shadow followed


def test_pgd_inner_loss_on_the_bf16_shadow(golden):
    """PGDAttack_moco(inner_queue="shadow"): the inner InfoNCE on the tcgen05 path; perturbation within the bf16 bar of
    the fp32 attack, sign agreement >= 99.9 % (north-star criterion for the PGD-perturbed tensors)."""
    import rmcl_b200
    from rmcl_b200 import ops
    g = golden("ref_facade_c128")
    mod = TinyModule(g, "auto").to(DEV).train()
    sh = mod.__dict__["_rmcl_queue_shadow"] = ops.QueueShadow()
    sh.get(mod.proj_queue)
    cfg = dict(adv_steps_img=g.i("meta/n_pgd"), adv_lr_img=g.f("meta/lr"), adv_max_norm_img=g.f("meta/eps"), max_image_len=200)
    k = torch.nn.functional.normalize(torch.randn(g.i("meta/B"), g.i("meta/C"), device=DEV), dim=1)
    d32 = rmcl_b200.PGDAttack_moco(cfg).pgd_attack(mod, deepcopy(_batch(g, 0)), k_modality=k)
    d16 = rmcl_b200.PGDAttack_moco(cfg, inner_queue="shadow").pgd_attack(mod, deepcopy(_batch(g, 0)), k_modality=k)
    eps = g.f("meta/eps")
    diff = (d16 - d32).abs()
    nz = d32.abs() > 5e-2 * eps          # elements at the noise level of the bf16 path (mean |diff| ~ 4e-3 eps) flip freely
    agree = (torch.sign(d16)[nz] == torch.sign(d32)[nz]).float().mean().item()
    print(f"bf16-inner PGD: max diff {diff.max().item() / eps:.3e} eps, mean {diff.mean().item() / eps:.3e} eps, sign agreement {agree:.5f}")
    assert diff.mean().item() <= 1e-2 * eps and diff.max().item() <= 0.25 * eps
    assert agree >= 0.999


def test_moco_module_api(golden):
    """MoCo sketch API (MoCo/MoCo_RMCL.py): method names, two enqueues per step, pointer advance."""
    import rmcl_b200

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.t, self.i = nn.Linear(12, 64), nn.Linear(20, 64)

        def forward(self, batch):
            return self.t(batch["t"]), self.i(batch["i"])

    torch.manual_seed(0)
    moco = rmcl_b200.MoCo({}, dim=64, K=256, m=0.9, T=0.07, encoder_q=Enc(), encoder_k=Enc(), infonce_path="simt").to(DEV)
    for pq, pk in zip(moco.encoder_q.parameters(), moco.encoder_k.parameters()):
        assert torch.equal(pq, pk) and not pk.requires_grad
    with torch.no_grad():
        for p in moco.encoder_q.parameters():
            p.add_(0.1)
    k0 = [p.clone() for p in moco.encoder_k.parameters()]
    batch = {"t": torch.randn(8, 12, device=DEV), "i": torch.randn(8, 20, device=DEV)}
    out, labels, logs, _ = moco(batch, materialize_logits=True)
    (out["loss_txt"] + out["loss_img"]).backward()
    assert moco.txt_img_queue_ptr.item() == 16                                    # text keys then image keys
    for a, b, q in zip(k0, moco.encoder_k.parameters(), moco.encoder_q.parameters()):
        assert torch.equal(b, a * 0.9 + q.detach() * (1.0 - 0.9))                # _momentum_update_key_encoder
    ce = torch.nn.functional.cross_entropy(out["txt"], labels["txt"])
    assert abs(ce.item() - out["loss_txt"].item()) < 1e-4 * abs(ce.item())
    assert moco.encoder_q.t.weight.grad is not None


# ------------------------------------------------- the other attackers of attack/pgd_attack_vilt.py
class OthersModule(nn.Module):
    """Stand-in of oracle/make_golden.py:make_pgd_others — toy encoder plus stand-in task heads."""

    def __init__(self, g):
        super().__init__()
        h, D, n_ans = g.i("meta/hidden"), g.i("meta/D"), g.i("meta/n_ans")
        self.text_embeddings = nn.Embedding(50, h)
        self.token_type_embeddings = nn.Embedding(3, h)
        self.transformer = ToyTransformer(h, 8)
        self.pooler = ToyPooler(h)
        self.barlowtwins_head = nn.Sequential(nn.Linear(h, D), nn.ReLU(), nn.Linear(D, D))
        self.nlvr2_classifier = nn.Sequential(nn.Linear(2 * h, h), nn.GELU(), nn.Linear(h, 2))
        self.vqa_classifier = nn.Sequential(nn.Linear(h, h), nn.GELU(), nn.Linear(h, n_ans))
        self.adv_lr = g.f("meta/adv_lr")
        self.max_image_len = 200

        class _H(dict):
            __getattr__ = dict.__getitem__
        self.hparams = _H(config={"vqav2_label_size": n_ans})
        own = self.state_dict()
        self.load_state_dict({k[len("state/"):]: g.t(k) for k in g.z.files
                              if k.startswith("state/") and k[len("state/"):] in own})

    @property
    def device(self):
        return next(self.parameters()).device

    def infer(self, batch, mask_text=False, mask_image=False, **kw):
        import rmcl_b200
        return rmcl_b200.PGDAttack.infer(self, batch, mask_text, mask_image, **kw)


def test_other_pgd_attackers_match_reference(golden):
    """PGDAttack_bartlowtwins / _nlvr2 / _vqa (pgd_attack_vilt.py:178-483) on the shared update kernel:
    perturbations after 3 steps against the unmodified reference classes run on CPU.  The update is
    bit-exact given the gradient; the gradient itself comes from torch on another device, hence 1e-4 of
    eps (elements sitting on the +-eps clamp are identical)."""
    import rmcl_b200
    from rmcl_b200 import pgd_attack as P
    g = golden("ref_pgd_others")
    mod = OthersModule(g).to(DEV).train()
    cfg = dict(adv_steps_img=g.i("meta/n_pgd"), adv_lr_img=g.f("meta/lr"), adv_max_norm_img=g.f("meta/eps"),
               max_image_len=200, attack_idx=[True, True])
    ids = g.t("batch/text_ids").to(DEV)
    base = {"text": ["x"] * ids.shape[0], "text_ids": ids, "text_labels": torch.full_like(ids, -100),
            "text_masks": torch.ones_like(ids), "vqa_labels": [[1, 3], [0], [5, 6, 7], [10]],
            "vqa_scores": [[1.0, 0.3], [0.6], [0.3, 0.3, 1.0], [0.9]]}
    eps = g.f("meta/eps")

    def close(got, name):
        want = g.t(f"delta/{name}")
        err = (got.detach().cpu() - want).abs().max().item()
        assert err <= 1e-4 * eps + 1e-7, f"{name}: max |delta - ref| = {err:.3e}"
        assert got.abs().max().item() <= float(np.float32(eps))

    one = dict(base, image=[g.t("batch/image").to(DEV)])
    close(P.PGDAttack_bartlowtwins(cfg, fused_loss=False).pgd_attack(mod, deepcopy(one), k_modality=g.t("k_barlowtwins").to(DEV)),
          "barlowtwins")                       # the reference's fp32 inner loss in torch
    close(P.PGDAttack_vqa(cfg).pgd_attack(mod, deepcopy(one)), "vqa")
    two = dict(base, image=[g.t("batch/image").to(DEV)], image_0=[g.t("batch/image_0").to(DEV)],
               image_1=[g.t("batch/image_1").to(DEV)], answers=[0, 1, 1, 0])
    d0, d1 = P.PGDAttack_nlvr2(cfg).pgd_attack(mod, deepcopy(two))
    close(d0, "nlvr2_0")
    close(d1, "nlvr2_1")
    d0, d1 = P.PGDAttack_nlvr2(dict(cfg, attack_idx=[False, True])).pgd_attack(mod, deepcopy(two))
    assert d0.abs().max().item() == 0.0
    close(d1, "nlvr2_only1_1")
    # fused inner loss (bf16 tensor-core Gram path): same perturbation within the bf16 bar, signs agree
    df = P.PGDAttack_bartlowtwins(cfg).pgd_attack(mod, deepcopy(one), k_modality=g.t("k_barlowtwins").to(DEV))   # the default
    want = g.t("delta/barlowtwins")
    assert (df.cpu() - want).abs().max().item() <= 2e-2 * eps
    nz = want.abs() > 1e-3 * eps
    assert (torch.sign(df.cpu())[nz] == torch.sign(want)[nz]).float().mean().item() >= 0.999
    # irtr: the reference body cannot run (undefined name); check the documented intent on its own terms
    k_txt = torch.nn.functional.normalize(torch.randn(ids.shape[0], 16, device=DEV), dim=1)
    mod.moco_head = MOCOHead(g.i("meta/hidden"), g.i("meta/hidden"), 16).to(DEV)
    d = P.PGDAttack_irtr(cfg).pgd_attack(mod, deepcopy(one), k_txt)
    assert d.shape == one["image"][0].shape and 0 < d.abs().max().item() <= float(np.float32(eps))
    # the copy_modules=True variant (the reference's deepcopy) gives the same perturbation
    a = P.PGDAttack_vqa(cfg, copy_modules=True).pgd_attack(mod, deepcopy(one))
    b = P.PGDAttack_vqa(cfg).pgd_attack(mod, deepcopy(one))
    assert torch.equal(a, b)


# ------------------------------------------------- Barlow-Twins objective (objectives.py:449-602)
class BarlowModule(nn.Module):
    """Stand-in of oracle/make_golden.py:make_barlow: toy encoder + stand-in projection head."""

    def __init__(self, g):
        super().__init__()
        import rmcl_b200
        h, D = g.i("meta/hidden"), g.i("meta/D")
        self.text_embeddings = nn.Embedding(50, h)
        self.token_type_embeddings = nn.Embedding(2, h)
        self.transformer = ToyTransformer(h, 8)
        self.pooler = ToyPooler(h)
        self.barlowtwins_head = nn.Sequential(nn.Linear(h, D), nn.ReLU(), nn.Linear(D, D))
        self.adv_lr, self.per_step_bs = g.f("meta/adv_lr"), g.i("meta/B")
        self.text_view, self.image_view, self.augmentation = False, True, False
        self.cosine = nn.CosineSimilarity(dim=1, eps=1e-6)
        self.max_image_len = 200
        cfg = dict(adv_steps_img=g.i("meta/n_pgd"), adv_lr_img=g.f("meta/lr"), adv_max_norm_img=g.f("meta/eps"),
                   max_image_len=200)
        self.pgd_attacker = rmcl_b200.PGDAttack_bartlowtwins(cfg, fused_loss=False)   # fp32 inner loss: the golden run is held to 1e-3 on the diagnostics
        for phase in ("train", "val"):
            for name in ("barlowtwins_loss", "barlowtwins_loss_invariance_img", "barlowtwins_loss_redundancy_img"):
                setattr(self, f"{phase}_{name}", lambda x: x)
        self.logged = {}
        own = self.state_dict()
        self.load_state_dict({k[len("state/"):]: g.t(k) for k in g.z.files
                              if k.startswith("state/") and k[len("state/"):] in own})

    @property
    def device(self):
        return next(self.parameters()).device

    def log(self, name, value, **kw):
        self.logged[name] = value

    def infer(self, batch, mask_text=False, mask_image=False, **kw):
        import rmcl_b200
        return rmcl_b200.PGDAttack.infer(self, batch, mask_text, mask_image, **kw)


def test_compute_barlowtwins_contrastive_matches_reference(golden):
    """The drop-in Barlow-Twins objective on the fused tcgen05 loss against the unmodified reference (CPU, fp32,
    materialised 96 x 96 matrix): two steps with backward through sum of the "*loss*" keys (vilt_module.py:475).
    The kernel rounds q, k and w(c - I) to bf16 (fp32 accumulation): north-star bf16 bar 2e-2."""
    import rmcl_b200
    g = golden("ref_barlow_facade")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    mod = BarlowModule(g).to(DEV).train()
    for s in range(g.i("meta/steps")):
        mod.zero_grad()
        mod.logged.clear()
        ret = rmcl_b200.compute_barlowtwins_contrastive(mod, _batch(g, s))
        want_keys = sorted(k.split("/")[-1] for k in g.z.files if k.startswith(f"step{s}/ret/"))
        assert sorted(ret) == want_keys
        total = sum(v for k, v in ret.items() if "loss" in k)
        total.backward()
        torch.cuda.synchronize()
        _close(total, g.np(f"step{s}/total_loss"), 2e-2, f"step {s} total loss")
        for k in want_keys:
            # step 1 runs on parameters nudged by step 0's (bf16-path) gradients: the diagnostics inherit that drift
            _close(ret[k], g.np(f"step{s}/ret/{k}"), 2e-2 if "loss" in k else (1e-3 if s == 0 else 5e-3), f"step {s} {k}")
        for k in g.z.files:
            if k.startswith(f"step{s}/log/"):
                name = k[len(f"step{s}/log/"):]
                _close(torch.as_tensor(mod.logged[name]), g.np(k), 2e-2, f"step {s} log {name}")
        n_grads = 0
        for k, v in mod.named_parameters():
            if f"step{s}/grad/{k}" in g:
                _close(v.grad, g.np(f"step{s}/grad/{k}"), 3e-2, f"step {s} grad {k}")
                n_grads += 1
        assert n_grads >= 10
        with torch.no_grad():
            for k, v in mod.named_parameters():
                if v.grad is not None:
                    v.add_(-0.02 * v.grad)
