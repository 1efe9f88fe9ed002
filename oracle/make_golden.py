"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference) on CPU.  Run in the authoring container only:

    python oracle/make_golden.py            # tiny cases (seconds)
    python oracle/make_golden.py --cfg1     # + real ViLT-B/32 at BASELINE cfg1 (~1 min)

What runs is the reference's own code:
  * vilt/modules/objectives.py:217-447  compute_moco_contrastive  (EMA, InfoNCE,
    compute_pgd, enqueue)
  * attack/pgd_attack_vilt.py:130-175   PGDAttack_moco.pgd_attack
A recorder intercepts the ATen entry points those lines call (F.normalize,
CrossEntropyLoss.forward, torch.norm, torch.clamp) so the *inputs and outputs of
each reference expression* are captured without touching the reference source.

"tiny" cases drive compute_moco_contrastive with a small stand-in LightningModule
(real reference objective + attacker + MOCOHead; toy encoder) so that shapes other
than C=128 and several consecutive steps can be pinned cheaply.  "cfg1" drives the
real ViLTransformerSS (random init, load_pretrained no-op'ed, drop_rate 0).
"""
import argparse
import contextlib
import os
import sys
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


# --------------------------------------------------------------------- recorder
class Recorder:
    """Intercepts the torch entry points the reference's hot path calls."""

    def __init__(self):
        self.normalize_in = []   # every F.normalize input (k-proj, q clean, PGD q's, q attacked)
        self.ce = []             # (logits, loss) of every CrossEntropyLoss call
        self.norm_in = []        # torch.norm inputs  (PGD: grad viewed [B,-1])
        self.clamp_out = []      # torch.clamp outputs with dim>=3 (PGD: clamped delta)

    @contextlib.contextmanager
    def active(self):
        F = torch.nn.functional
        o_norm, o_ce, o_tnorm, o_clamp = F.normalize, nn.CrossEntropyLoss.forward, torch.norm, torch.clamp
        rec = self

        def normalize(x, *a, **k):
            if x.requires_grad:
                x.retain_grad()
            rec.normalize_in.append(x)
            return o_norm(x, *a, **k)

        def ce_forward(self_, logits, labels):
            out = o_ce(self_, logits, labels)
            rec.ce.append((logits.detach().clone(), out.detach().clone()))
            return out

        def tnorm(x, *a, **k):
            if k.get("p", None) == float("inf"):     # pgd_attack_vilt.py:164 only
                rec.norm_in.append(x.detach().clone())
            return o_tnorm(x, *a, **k)

        def clamp(x, *a, **k):
            out = o_clamp(x, *a, **k)
            if x.dim() >= 3 and (x.shape[-1] > 1):
                rec.clamp_out.append(out.detach().clone())
            return out

        F.normalize, nn.CrossEntropyLoss.forward, torch.norm, torch.clamp = normalize, ce_forward, tnorm, clamp
        try:
            yield self
        finally:
            F.normalize, nn.CrossEntropyLoss.forward, torch.norm, torch.clamp = o_norm, o_ce, o_tnorm, o_clamp


def _np(t):
    return t.detach().cpu().numpy().copy()


def run_reference_step(ref, pl_module, batch, k_layers, q_layers, tag, out, backward=True):
    """One call of the reference's compute_moco_contrastive, fully recorded."""
    n_pgd = pl_module.pgd_attacker.adv_steps_img
    pk_before = [p.detach().clone() for l in k_layers for p in l.parameters()]
    pq = [p.detach().clone() for l in q_layers for p in l.parameters()]
    queue_before = pl_module.proj_queue.detach().clone()
    ptr_before = int(pl_module.proj_queue_ptr)

    rec = Recorder()
    with rec.active():
        ret = ref.objectives.compute_moco_contrastive(pl_module, batch)
        if backward:
            ret["moco_loss"].backward()

    pk_after = [p.detach().clone() for l in k_layers for p in l.parameters()]
    # order of F.normalize calls: k-proj, q clean, n_pgd x PGD q, q attacked
    assert len(rec.normalize_in) == 3 + n_pgd, len(rec.normalize_in)
    assert len(rec.ce) == n_pgd + 1
    assert len(rec.norm_in) == n_pgd
    k_raw = rec.normalize_in[0]
    k_hat = torch.nn.functional.normalize(k_raw.detach(), dim=1)

    out[f"{tag}/momentum"] = np.float64(pl_module.momentum)
    out[f"{tag}/temperature"] = np.float64(pl_module.temperature)
    out[f"{tag}/n_pgd"] = np.int64(n_pgd)
    out[f"{tag}/adv_lr"] = np.float64(pl_module.pgd_attacker.adv_lr_img)
    out[f"{tag}/adv_eps"] = np.float64(pl_module.pgd_attacker.adv_max_norm_img)
    out[f"{tag}/k_raw"] = _np(k_raw)
    out[f"{tag}/k_hat"] = _np(k_hat)
    out[f"{tag}/queue_before"] = _np(queue_before)
    out[f"{tag}/ptr_before"] = np.int64(ptr_before)
    out[f"{tag}/ptr_after"] = np.int64(int(pl_module.proj_queue_ptr))
    qa = pl_module.proj_queue.detach()
    changed = (qa != queue_before).any(dim=0).nonzero().flatten()
    out[f"{tag}/queue_changed_cols"] = _np(changed)
    out[f"{tag}/queue_after_cols"] = _np(qa[:, ptr_before:ptr_before + k_hat.shape[0]])
    out[f"{tag}/queue_after_sum64"] = np.float64(qa.double().sum().item())
    # clean query: only logits.argmax is used (objectives.py:275)
    out[f"{tag}/q_clean_raw"] = _np(rec.normalize_in[1])
    # PGD inner InfoNCE calls (loss divided by adv_steps, pgd_attack_vilt.py:158)
    for s in range(n_pgd):
        qx = rec.normalize_in[2 + s]
        out[f"{tag}/pgd{s}/q_raw"] = _np(qx)
        out[f"{tag}/pgd{s}/dq_raw"] = _np(qx.grad)
        out[f"{tag}/pgd{s}/logits"] = _np(rec.ce[s][0])
        out[f"{tag}/pgd{s}/loss"] = _np(rec.ce[s][1])
        out[f"{tag}/pgd{s}/grad"] = _np(rec.norm_in[s])           # [B, numel]
        out[f"{tag}/pgd{s}/delta_after"] = _np(rec.clamp_out[s]) if rec.clamp_out else None
    # final (attacked) query InfoNCE — the training loss (objectives.py:324-351)
    qf = rec.normalize_in[2 + n_pgd]
    out[f"{tag}/q_raw"] = _np(qf)
    out[f"{tag}/logits"] = _np(rec.ce[n_pgd][0])
    out[f"{tag}/loss"] = _np(rec.ce[n_pgd][1])
    out[f"{tag}/moco_loss"] = _np(ret["moco_loss"])
    if backward:
        out[f"{tag}/dq_raw"] = _np(qf.grad)
    for name in ("pos_dist_attacked_img", "pos_cosine_attacked_img", "pos_dot_attacked_img",
                 "neg_dist_attacked_img", "neg_cosine_attacked_img", "neg_dot_attacked_img"):
        out[f"{tag}/diag/{name}"] = _np(ret[name])
    logged = pl_module.__dict__.get("_logged", {})
    if "moco_attack/PGD_success_rate" in logged:
        out[f"{tag}/pgd_success_rate"] = _np(logged["moco_attack/PGD_success_rate"])
    if "moco_attack/train/delta" in logged:
        out[f"{tag}/delta_range"] = _np(logged["moco_attack/train/delta"])
    return pk_before, pq, pk_after, rec


# ------------------------------------------------------------ tiny stand-in model
class _ToyBlock(nn.Module):
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


class _ToyTransformer(nn.Module):
    """Just enough of vision_transformer.VisionTransformer for PGDAttack.infer
    (attack/pgd_attack_vilt.py:29-106): visual_embed, blocks, norm."""

    def __init__(self, h, patch):
        super().__init__()
        self.patch_embed = nn.Conv2d(3, h, patch, patch)
        self.cls_token = nn.Parameter(torch.randn(1, 1, h) * 0.02)
        self.blocks = nn.ModuleList([_ToyBlock(h), _ToyBlock(h)])
        self.norm = nn.LayerNorm(h)

    def visual_embed(self, img, max_image_len=200, mask_it=False):
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        masks = torch.ones(x.shape[0], x.shape[1], dtype=torch.long)
        return x, masks, None, None


class _ToyPooler(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.dense = nn.Linear(h, h)

    def forward(self, x):
        return torch.tanh(self.dense(x[:, 0]))


def build_tiny_module(ref, B, C, K, hidden, n_pgd, lr, eps, T, m, seed):
    torch.manual_seed(seed)
    pl = sys.modules["pytorch_lightning"]
    PGDAttack = ref.pgd.PGDAttack

    class Tiny(pl.LightningModule):
        def __init__(self):
            super().__init__()
            h = hidden
            self.text_embeddings = nn.Embedding(50, h)
            self.token_type_embeddings = nn.Embedding(2, h)
            self.transformer = _ToyTransformer(h, 8)
            self.pooler = _ToyPooler(h)
            self.moco_head = ref.heads.MOCOHead(h, h, C)          # reference head
            self.k_text_embeddings = deepcopy(self.text_embeddings)
            self.k_token_type_embeddings = deepcopy(self.token_type_embeddings)
            self.k_transformer = deepcopy(self.transformer)
            self.k_moco_head = deepcopy(self.moco_head)
            for l in (self.k_text_embeddings, self.k_token_type_embeddings, self.k_transformer, self.k_moco_head):
                for p in l.parameters():
                    p.requires_grad = False
            self.momentum, self.temperature = m, T
            self.text_view, self.image_view, self.augmentation = False, True, False
            self.num_negative, self.per_step_bs = K, B
            self.cosine = nn.CosineSimilarity(dim=1, eps=1e-6)
            self.register_buffer("proj_queue", torch.randn(C, K))
            self.register_buffer("proj_queue_ptr", torch.zeros(1, dtype=torch.long))
            cfg = dict(adv_steps_img=n_pgd, adv_lr_img=lr, adv_max_norm_img=eps, max_image_len=200)
            self.pgd_attacker = ref.pgd.PGDAttack_moco(cfg)        # reference attacker
            self.max_image_len = 200
            self.train_moco_loss = lambda x: x
            self.val_moco_loss = lambda x: x

        # q / k forwards reuse the reference's mini-ViLT infer (pgd_attack_vilt.py:29-106)
        def infer(self, batch, mask_text=False, mask_image=False):
            return PGDAttack.infer(self, batch, mask_text, mask_image)

        def infer_k(self, batch, mask_text=False, mask_image=False):
            import types
            view = types.SimpleNamespace(
                text_embeddings=self.k_text_embeddings, token_type_embeddings=self.k_token_type_embeddings,
                transformer=self.k_transformer, pooler=self.pooler, max_image_len=200)
            return PGDAttack.infer(view, batch, mask_text, mask_image)

    mod = Tiny()
    mod.train()
    # make q and k params differ so the EMA is not a no-op
    with torch.no_grad():
        for p in list(mod.text_embeddings.parameters()) + list(mod.transformer.parameters()) + \
                list(mod.moco_head.parameters()) + list(mod.token_type_embeddings.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    return mod


def tiny_batch(B, img, L, seed):
    g = torch.Generator().manual_seed(seed)
    return {
        "image": [torch.randn(B, 3, img, img, generator=g)],
        "text": ["x"] * B,
        "text_ids": torch.randint(1, 50, (B, L), generator=g),
        "text_labels": torch.full((B, L), -100),
        "text_masks": torch.ones(B, L, dtype=torch.long),
    }


def make_tiny(ref, name, B, C, K, n_pgd, lr, eps, T=0.07, m=0.999, steps=2, seed=0):
    mod = build_tiny_module(ref, B, C, K, hidden=32, n_pgd=n_pgd, lr=lr, eps=eps, T=T, m=m, seed=seed)
    k_layers = [mod.k_text_embeddings, mod.k_token_type_embeddings, mod.k_transformer, mod.k_moco_head]
    q_layers = [mod.text_embeddings, mod.token_type_embeddings, mod.transformer, mod.moco_head]
    out = {"meta/B": np.int64(B), "meta/C": np.int64(C), "meta/K": np.int64(K), "meta/steps": np.int64(steps)}
    for s in range(steps):
        mod.zero_grad()
        batch = tiny_batch(B, 16, 6, seed * 100 + s)
        pkb, pq, pka, _ = run_reference_step(ref, mod, batch, k_layers, q_layers, f"step{s}", out)
        for i, (a, b, c) in enumerate(zip(pkb, pq, pka)):
            out[f"step{s}/ema/k_before/{i}"] = _np(a)
            out[f"step{s}/ema/q/{i}"] = _np(b)
            out[f"step{s}/ema/k_after/{i}"] = _np(c)
        out[f"step{s}/ema/n"] = np.int64(len(pkb))
        # nudge q-params like an optimiser would, so step 1's EMA sees new values
        with torch.no_grad():
            for l in q_layers:
                for p in l.parameters():
                    if p.grad is not None:
                        p.add_(-0.1 * p.grad)
    out = {k: v for k, v in out.items() if v is not None}
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB, loss step0 = {out['step0/loss']}")


# ---------------------------------------------------------------- cfg1 (real ViLT)
def cfg1_config():
    """vilt/config.py:24-116 defaults overlaid with task_moco (128-164) and the
    BASELINE cfg1 overrides."""
    loss_names = {k: 0 for k in ("itm", "mlm", "mpp", "vqa", "vqa_attacked", "nlvr2", "nlvr2_attacked",
                                 "irtr", "irtr_attacked", "moco", "barlowtwins")}
    loss_names["moco"] = 1
    return dict(
        exp_name="moco", seed=0, datasets=["coco"], loss_names=loss_names, batch_size=8,
        train_transform_keys=["pixelbert"], val_transform_keys=["pixelbert"], image_size=384, max_image_len=200,
        patch_size=32, draw_false_image=1, image_only=False, vqav2_label_size=3129, max_text_len=40,
        tokenizer="bert-base-uncased", vocab_size=30522, whole_word_masking=False, mlm_prob=0.15, draw_false_text=0,
        vit="vit_base_patch32_384", hidden_size=768, num_heads=12, num_layers=12, mlp_ratio=4, drop_rate=0.0,
        optim_type="adamw", learning_rate=1e-4, weight_decay=0.01, decay_power=1, max_epoch=1, max_steps=25000,
        warmup_steps=2500, end_lr=0, lr_mult=1, get_recall_metric=False, Multimodal=True, num_negative=4096,
        text_view=False, image_view=True, augmentation=False, num_beams=5, num_return_sequences=5,
        type_txt_augm=["PEGASUS", "EDA"], momentum=0.999, temperature=0.07, adv_lr=0.0051, TSNE_vizualisation=False,
        img_save_path="", adv_steps_img=1, adv_lr_img=0.05, adv_max_norm_img=8.0 / 255.0, attack_idx=[False, False],
        n_candidates=5, max_loops=10, sim_thred=0.5, cos_sim=True, synonym="cos_sim", embedding_path="", sim_path="",
        resume_from=None, fast_dev_run=False, val_check_interval=1.0, test_only=False, data_root="", log_dir="",
        per_gpu_batchsize=8, num_gpus=1, num_nodes=1, load_path="", num_workers=0, precision=32,
    )


def make_cfg1(ref):
    vilt_module, vilt_utils = ref_harness.load_reference_model_module()
    torch.manual_seed(0)
    cfg = cfg1_config()
    model = vilt_module.ViLTransformerSS(cfg)
    model.train()
    vilt_utils.set_task(model)
    B = 8
    g = torch.Generator().manual_seed(0)
    batch = {
        "image": [torch.randn(B, 3, 384, 384, generator=g)],
        "text": ["a photo"] * B,
        "text_ids": torch.randint(1000, 30000, (B, 40), generator=g),
        "text_labels": torch.full((B, 40), -100),
        "text_masks": torch.ones(B, 40, dtype=torch.long),
    }
    # let q drift from k so the EMA changes something (a fresh model has k == q)
    with torch.no_grad():
        for p in model.moco_head.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
        for p in model.token_type_embeddings.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    k_layers = [model.k_text_embeddings, model.k_token_type_embeddings, model.k_transformer, model.k_moco_head]
    q_layers = [model.text_embeddings, model.token_type_embeddings, model.transformer, model.moco_head]
    out = {"meta/B": np.int64(B), "meta/C": np.int64(128), "meta/K": np.int64(4096), "meta/steps": np.int64(1)}
    torch.manual_seed(1)  # visual_embed's multinomial patch permutation (F12)
    pkb, pq, pka, rec = run_reference_step(ref, model, batch, k_layers, q_layers, "step0", out)
    # EMA: the shape list of all 161 pairs + float64 checksums; full data only for the small tensors
    shapes = [tuple(p.shape) for p in pkb]
    out["step0/ema/n"] = np.int64(len(pkb))
    out["step0/ema/numels"] = np.array([p.numel() for p in pkb], dtype=np.int64)
    out["step0/ema/sum64_after"] = np.array([p.double().sum().item() for p in pka])
    keep = [i for i, p in enumerate(pkb) if p.numel() <= 768 * 128 and (pkb[i] != pq[i]).any()]
    out["step0/ema/kept"] = np.array(keep, dtype=np.int64)
    for i in keep:
        out[f"step0/ema/k_before/{i}"] = _np(pkb[i])
        out[f"step0/ema/q/{i}"] = _np(pq[i])
        out[f"step0/ema/k_after/{i}"] = _np(pka[i])
    # PGD tensors are 8x3x384x384 (14 MB each): keep sample 0 only (the update is per sample)
    keep_b = [0]
    out["step0/pgd_kept_samples"] = np.array(keep_b, dtype=np.int64)
    for s in range(int(out["step0/n_pgd"])):
        out[f"step0/pgd{s}/grad"] = out[f"step0/pgd{s}/grad"][keep_b]
        out[f"step0/pgd{s}/delta_after"] = out[f"step0/pgd{s}/delta_after"][keep_b].reshape(len(keep_b), -1)
    out = {k: v for k, v in out.items() if v is not None}
    path = os.path.join(GOLDEN_DIR, "ref_cfg1_vilt_b32.npz")
    np.savez_compressed(path, **out)
    with open(os.path.join(GOLDEN_DIR, "vilt_b32_key_encoder_shapes.txt"), "w") as f:
        f.write("# 161 key-encoder parameter shapes of ViLT-B/32 (+BertEmbeddings, token-type, MOCOHead), in\n"
                "# the order objectives.py:257-260 walks them. Generated by oracle/make_golden.py --cfg1.\n")
        for s in shapes:
            f.write("x".join(str(d) for d in s) + "\n")
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; loss={out['step0/loss']}; "
          f"{len(shapes)} tensors, {sum(int(np.prod(s)) for s in shapes)} params")


def make_pgd_direct(ref):
    """PGDAttack_moco.pgd_attack driven directly (no objective) — multi-step, eps=0
    (no clamp branch, pgd_attack_vilt.py:172) and an all-zero-gradient sample."""
    for name, n_pgd, lr, eps in (("ref_pgd_5step", 5, 0.05, 8.0 / 255.0), ("ref_pgd_noclamp", 3, 0.05, 0.0)):
        mod = build_tiny_module(ref, 4, 16, 64, hidden=32, n_pgd=n_pgd, lr=lr, eps=eps, T=0.07, m=0.999, seed=3)
        batch = tiny_batch(4, 16, 6, 77)
        with torch.no_grad():
            k = torch.nn.functional.normalize(mod.k_moco_head(mod.infer_k(batch)["cls_feats"]), dim=1)
        rec = Recorder()
        with rec.active():
            delta = mod.pgd_attacker.pgd_attack(mod, deepcopy(batch), k_modality=k)
        out = {"n_pgd": np.int64(n_pgd), "lr": np.float64(lr), "eps": np.float64(eps), "delta_final": _np(delta),
               "k_hat": _np(k), "queue": _np(mod.proj_queue), "temperature": np.float64(0.07)}
        for s in range(n_pgd):
            out[f"pgd{s}/grad"] = _np(rec.norm_in[s])
            out[f"pgd{s}/q_raw"] = _np(rec.normalize_in[s])
            out[f"pgd{s}/dq_raw"] = _np(rec.normalize_in[s].grad)
            out[f"pgd{s}/logits"] = _np(rec.ce[s][0])
            out[f"pgd{s}/loss"] = _np(rec.ce[s][1])
            if eps > 0:
                out[f"pgd{s}/delta_after"] = _np(rec.clamp_out[s])
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB")


def make_pgd_others(ref):
    """The other attackers of attack/pgd_attack_vilt.py (bartlowtwins 178-239, nlvr2 241-342, vqa 418-483)
    driven directly on the tiny module with stand-in heads; PGDAttack_irtr cannot run in the reference
    (undefined name at line 391) and has no vector."""
    B, hidden, D, n_ans = 4, 32, 24, 11
    mod = build_tiny_module(ref, B, 16, 64, hidden=hidden, n_pgd=3, lr=0.05, eps=8.0 / 255.0, T=0.07, m=0.999, seed=9)
    torch.manual_seed(91)
    mod.token_type_embeddings = nn.Embedding(3, hidden)          # nlvr2 uses image token type 2
    mod.barlowtwins_head = nn.Sequential(nn.Linear(hidden, D), nn.ReLU(), nn.Linear(D, D))
    mod.nlvr2_classifier = nn.Sequential(nn.Linear(2 * hidden, hidden), nn.GELU(), nn.Linear(hidden, 2))
    mod.vqa_classifier = nn.Sequential(nn.Linear(hidden, hidden), nn.GELU(), nn.Linear(hidden, n_ans))
    mod.adv_lr = 0.0051
    mod.hparams = type(mod.hparams)(config={"vqav2_label_size": n_ans}) if hasattr(mod, "hparams") else None
    if mod.hparams is None:
        class _H(dict):
            __getattr__ = dict.__getitem__
        mod.hparams = _H(config={"vqav2_label_size": n_ans})
    cfg = dict(adv_steps_img=3, adv_lr_img=0.05, adv_max_norm_img=8.0 / 255.0, max_image_len=200, attack_idx=[True, True])
    out = {"meta/B": np.int64(B), "meta/hidden": np.int64(hidden), "meta/D": np.int64(D), "meta/n_ans": np.int64(n_ans),
           "meta/adv_lr": np.float64(mod.adv_lr), "meta/n_pgd": np.int64(3), "meta/lr": np.float64(0.05),
           "meta/eps": np.float64(8.0 / 255.0)}
    for k, v in mod.state_dict().items():
        out[f"state/{k}"] = _np(v)
    batch = tiny_batch(B, 16, 6, 901)
    g = torch.Generator().manual_seed(902)
    # (PGDAttack.infer reads batch["image_0"] whenever that key exists, pgd_attack_vilt.py:38-41, so the
    #  two-image NLVR2 batch is a separate dict)
    batch2 = dict(batch)
    batch2["image_0"] = [torch.randn(B, 3, 16, 16, generator=g)]
    batch2["image_1"] = [torch.randn(B, 3, 16, 16, generator=g)]
    batch2["answers"] = [0, 1, 1, 0]
    batch["vqa_labels"] = [[1, 3], [0], [5, 6, 7], [10]]
    batch["vqa_scores"] = [[1.0, 0.3], [0.6], [0.3, 0.3, 1.0], [0.9]]
    k_bt = torch.randn(B, D, generator=g)
    for n in ("image", "image_0", "image_1"):
        out[f"batch/{n}"] = _np(batch2[n][0])
    out["batch/text_ids"] = _np(batch["text_ids"])
    out["k_barlowtwins"] = _np(k_bt)
    d = ref.pgd.PGDAttack_bartlowtwins(cfg).pgd_attack(mod, deepcopy(batch), k_modality=k_bt)
    out["delta/barlowtwins"] = _np(d)
    d0, d1 = ref.pgd.PGDAttack_nlvr2(cfg).pgd_attack(mod, deepcopy(batch2))
    out["delta/nlvr2_0"], out["delta/nlvr2_1"] = _np(d0), _np(d1)
    d0, d1 = ref.pgd.PGDAttack_nlvr2(dict(cfg, attack_idx=[False, True])).pgd_attack(mod, deepcopy(batch2))
    out["delta/nlvr2_only1_0"], out["delta/nlvr2_only1_1"] = _np(d0), _np(d1)
    d = ref.pgd.PGDAttack_vqa(cfg).pgd_attack(mod, deepcopy(batch))
    out["delta/vqa"] = _np(d)
    path = os.path.join(GOLDEN_DIR, "ref_pgd_others.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB; |delta| "
          + ", ".join(f"{k[6:]}={float(np.abs(v).mean()):.4f}" for k, v in out.items() if k.startswith("delta/")))


def make_facade(ref):
    """Whole-step golden for the drop-in facade: the tiny module's complete state, the batch, and
    everything the reference's compute_moco_contrastive + backward produced (loss, diagnostics,
    logged values, queue/pointer, EMA'd key params, gradients of every query param)."""
    for name, B, C, K, n_pgd, lr, eps, seed in (("ref_facade_c128", 8, 128, 256, 2, 0.05, 8.0 / 255.0, 5),
                                                   ("ref_facade_c16", 4, 16, 64, 1, 0.05, 0.005, 6)):
        mod = build_tiny_module(ref, B, C, K, hidden=32, n_pgd=n_pgd, lr=lr, eps=eps, T=0.07, m=0.999, seed=seed)
        out = {"meta/B": np.int64(B), "meta/C": np.int64(C), "meta/K": np.int64(K), "meta/hidden": np.int64(32),
               "meta/n_pgd": np.int64(n_pgd), "meta/lr": np.float64(lr), "meta/eps": np.float64(eps),
               "meta/T": np.float64(0.07), "meta/m": np.float64(0.999), "meta/steps": np.int64(2)}
        for k, v in mod.state_dict().items():
            out[f"state/{k}"] = _np(v)
        for s in range(2):
            mod.zero_grad()
            batch = tiny_batch(B, 16, 6, seed * 100 + s)
            out[f"step{s}/batch/image"] = _np(batch["image"][0])
            out[f"step{s}/batch/text_ids"] = _np(batch["text_ids"])
            ret = ref.objectives.compute_moco_contrastive(mod, deepcopy(batch))
            ret["moco_loss"].backward()
            for k, v in ret.items():
                out[f"step{s}/ret/{k}"] = _np(v)
            for k, v in mod.__dict__.get("_logged", {}).items():
                out[f"step{s}/log/{k}"] = _np(torch.as_tensor(v))
            out[f"step{s}/queue_after"] = _np(mod.proj_queue)
            out[f"step{s}/ptr_after"] = np.int64(int(mod.proj_queue_ptr))
            for k, v in mod.named_parameters():
                if k.startswith("k_"):
                    out[f"step{s}/k_after/{k}"] = _np(v)
                elif v.grad is not None:
                    out[f"step{s}/grad/{k}"] = _np(v.grad)
            with torch.no_grad():   # optimiser-like nudge so that step 1 differs
                for k, v in mod.named_parameters():
                    if v.grad is not None:
                        v.add_(-0.1 * v.grad)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB, losses "
              f"{out['step0/ret/moco_loss']}, {out['step1/ret/moco_loss']}")


def make_barlow(ref):
    """Whole-step golden of the Barlow-Twins objective (objectives.py:449-602, image view, PGD attacker
    PGDAttack_bartlowtwins): the tiny module with a stand-in projection head, two consecutive steps with
    backward, everything the reference returned and logged, plus the gradients of every parameter of
    sum(v for k, v in ret.items() if "loss" in k) — the training loss of vilt_module.py:475."""
    B, hidden, D, n_pgd, lr, eps = 16, 32, 96, 2, 0.05, 8.0 / 255.0
    mod = build_tiny_module(ref, B, 16, 64, hidden=hidden, n_pgd=n_pgd, lr=lr, eps=eps, T=0.07, m=0.999, seed=12)
    torch.manual_seed(121)
    mod.barlowtwins_head = nn.Sequential(nn.Linear(hidden, D), nn.ReLU(), nn.Linear(D, D))
    with torch.no_grad():                       # projections of O(1) spread so that diagonal(c) is O(1) as in training
        mod.barlowtwins_head[2].weight.mul_(6.0)
    mod.adv_lr = 0.0051
    cfg = dict(adv_steps_img=n_pgd, adv_lr_img=lr, adv_max_norm_img=eps, max_image_len=200)
    mod.pgd_attacker = ref.pgd.PGDAttack_bartlowtwins(cfg)
    for phase in ("train", "val"):
        for name in ("barlowtwins_loss", "barlowtwins_loss_invariance_img", "barlowtwins_loss_redundancy_img"):
            setattr(mod, f"{phase}_{name}", lambda x: x)
    out = {"meta/B": np.int64(B), "meta/hidden": np.int64(hidden), "meta/D": np.int64(D), "meta/n_pgd": np.int64(n_pgd),
           "meta/lr": np.float64(lr), "meta/eps": np.float64(eps), "meta/adv_lr": np.float64(mod.adv_lr),
           "meta/steps": np.int64(2)}
    for k, v in mod.state_dict().items():
        out[f"state/{k}"] = _np(v)
    for s in range(2):
        mod.zero_grad()
        batch = tiny_batch(B, 16, 6, 1200 + s)
        out[f"step{s}/batch/image"] = _np(batch["image"][0])
        out[f"step{s}/batch/text_ids"] = _np(batch["text_ids"])
        ret = ref.objectives.compute_barlowtwins_contrastive(mod, deepcopy(batch))
        total = sum(v for k, v in ret.items() if "loss" in k)
        total.backward()
        out[f"step{s}/total_loss"] = _np(total)
        for k, v in ret.items():
            out[f"step{s}/ret/{k}"] = _np(v)
        for k, v in mod.__dict__.get("_logged", {}).items():
            out[f"step{s}/log/{k}"] = _np(torch.as_tensor(v))
        for k, v in mod.named_parameters():
            if v.grad is not None:
                out[f"step{s}/grad/{k}"] = _np(v.grad)
        with torch.no_grad():
            for k, v in mod.named_parameters():
                if v.grad is not None:
                    v.add_(-0.02 * v.grad)
    path = os.path.join(GOLDEN_DIR, "ref_barlow_facade.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB, losses {out['step0/ret/barlowtwins_loss']}, "
          f"{out['step1/ret/barlowtwins_loss']} (invariance {out['step0/ret/barlowtwins_loss_invariance_img']}, "
          f"redundancy {out['step0/ret/barlowtwins_loss_redundancy_img']})")

    # kernel-level vectors: the reference expression chain itself (objectives.py:480-486) on random projections
    g = torch.Generator().manual_seed(122)
    vec = {}
    for name, Bv, Dv in (("a", 8, 64), ("b", 32, 200)):
        k = torch.randn(Bv, Dv, generator=g)
        q = (0.7 * k + 0.7 * torch.randn(Bv, Dv, generator=g)).requires_grad_(True)
        c = q.T @ k
        c.div_(Bv)
        torch.distributed.all_reduce(c)
        on_diag = torch.diagonal(c).add_(-1).pow_(2).sum()
        n = c.shape[0]
        off_diag = c.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten().pow_(2).sum()
        loss = on_diag + 0.0051 * off_diag
        loss.backward()
        for kk, vv in (("q", q), ("k", k), ("on_diag", on_diag), ("off_diag", off_diag), ("loss", loss), ("dq", q.grad)):
            vec[f"{name}/{kk}"] = _np(vv)
    path = os.path.join(GOLDEN_DIR, "ref_barlow_vectors.npz")
    np.savez_compressed(path, **vec)
    print(f"wrote {path}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg1", action="store_true")
    ap.add_argument("--only-cfg1", action="store_true")
    ap.add_argument("--only-facade", action="store_true")
    ap.add_argument("--only-pgd-others", action="store_true")
    ap.add_argument("--only-barlow", action="store_true")
    args = ap.parse_args()
    ref = ref_harness.load_reference()
    ref_harness.ensure_process_group()
    torch.set_num_threads(8)
    if args.only_facade:
        make_facade(ref)
        return
    if args.only_pgd_others:
        make_pgd_others(ref)
        return
    if args.only_barlow:
        make_barlow(ref)
        return
    if not args.only_cfg1:
        make_facade(ref)
        make_tiny(ref, "ref_tiny_c16", B=4, C=16, K=64, n_pgd=1, lr=0.05, eps=8.0 / 255.0, steps=3)
        make_tiny(ref, "ref_tiny_c128", B=8, C=128, K=256, n_pgd=3, lr=0.05, eps=0.005, steps=2, seed=1)
        make_pgd_direct(ref)
        make_pgd_others(ref)
        make_barlow(ref)
    if args.cfg1 or args.only_cfg1:
        make_cfg1(ref)


if __name__ == "__main__":
    main()
