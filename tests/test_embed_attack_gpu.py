"""GPU: the embedding-space PGD attacker (north_star (d): "a fused PGD step on patch/token embeddings";
SURVEY 8(f) N3: attack through the ``image_embeds`` hook of vilt_module.py:275-311).

ViLT's ``visual_embed`` selects and permutes patches at random whenever an image has at least ``max_image_len``
patches, so the perturbation only means something together with the embeddings and masks of the ONE call it was
optimised against.  The stand-in transformer below does the same (a fresh random patch subset per call) and
counts its calls; the tests check that
  * attack + attacked view run ``visual_embed`` once, and the attacked forward receives exactly base + delta;
  * delta equals an oracle loop (the reference's expressions, attack/pgd_attack_vilt.py:141-173, on CPU in fp32
    through torch autograd, update by oracle.pgd_update) started from the same embeddings;
  * "embed" perturbs text and image tokens with one [B, L_text+L_image, H] tensor, "embed_image" the image tokens.
"""
from copy import deepcopy

import pytest
import torch
import torch.nn.functional as F

import rmcl_oracle as O
from test_facade_gpu import TinyModule, ToyTransformer

pytestmark = pytest.mark.gpu
DEV = "cuda"
MAX_IMAGE_LEN = 9          # of 16 patches


class SamplingTransformer(ToyTransformer):
    """visual_embed with ViLT's random patch selection (vision_transformer.py: torch.multinomial over the valid
    patches when there are more than max_image_len): same parameters as the toy encoder, non-repeatable output."""
    calls = 0

    def visual_embed(self, img, max_image_len=200, mask_it=False):
        type(self).calls += 1
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        B, P, H = x.shape
        if 0 < max_image_len < P:
            sel = torch.stack([torch.randperm(P, device=x.device)[:max_image_len] for _ in range(B)])
            x = torch.gather(x, 1, sel[:, :, None].expand(-1, -1, H))
        x = torch.cat([self.cls_token.expand(B, -1, -1), x], dim=1)
        return x, torch.ones(B, x.shape[1], dtype=torch.long, device=x.device), None, None


def _module(golden, space, device=DEV):
    import rmcl_b200
    g = golden("ref_facade_c128")
    mod = TinyModule(g, "simt")
    sampling = SamplingTransformer(g.i("meta/hidden"), 8)
    sampling.load_state_dict(mod.transformer.state_dict())
    mod.transformer = sampling
    mod.k_transformer = deepcopy(sampling)
    mod.max_image_len = MAX_IMAGE_LEN

    class _H(dict):
        __getattr__ = dict.__getitem__
    mod.hparams = _H(config={"max_image_len": MAX_IMAGE_LEN})
    cfg = dict(adv_steps_img=3, adv_lr_img=g.f("meta/lr"), adv_max_norm_img=0.05, max_image_len=MAX_IMAGE_LEN)
    mod.pgd_attacker = rmcl_b200.PGDAttack_moco(cfg, space=space, infonce_path="simt")
    with torch.no_grad():                     # the golden state has a zero queue: give the loss real negatives
        mod.proj_queue.copy_(torch.randn(mod.proj_queue.shape, generator=torch.Generator().manual_seed(3)))
    return mod.to(device).train(), g, cfg


def _batch(g):
    gen = torch.Generator().manual_seed(11)
    ids = g.t("step0/batch/text_ids").to(DEV)
    img = torch.randn(ids.shape[0], 3, 32, 32, generator=gen).to(DEV)
    return {"image": [img], "text": ["x"] * img.shape[0], "text_ids": ids,
            "text_labels": torch.full_like(ids, -100), "text_masks": torch.ones_like(ids)}


def _oracle_loop(mod_cpu, batch_cpu, base, masks, k_hat, queue, T, cfg, space):
    """The reference loop (pgd_attack_vilt.py:136-173) with the perturbation moved onto the embeddings."""
    n_txt = batch_cpu["text_ids"].shape[1] if space == "embed" else 0
    delta = torch.zeros(base.shape[0], n_txt + base.shape[1], base.shape[2])
    n = cfg["adv_steps_img"]
    for _ in range(n):
        delta.requires_grad_(True)
        text = mod_cpu.text_embeddings(batch_cpu["text_ids"])
        if n_txt:
            text = text + delta[:, :n_txt]
        text = text + mod_cpu.token_type_embeddings(torch.zeros_like(batch_cpu["text_masks"]))
        image = base + delta[:, n_txt:] + mod_cpu.token_type_embeddings(torch.full_like(masks, 1))
        x = torch.cat([text, image], dim=1)
        for blk in mod_cpu.transformer.blocks:
            x, _ = blk(x)
        q_attacked = F.normalize(mod_cpu.moco_head(mod_cpu.pooler(mod_cpu.transformer.norm(x))), dim=1)
        logits = O.info_nce_logits(q_attacked, k_hat, queue, T)
        loss = F.cross_entropy(logits.float(), torch.zeros(logits.shape[0], dtype=torch.long)) / (1.0 * n)
        (grad,) = torch.autograd.grad(loss, delta)
        delta = O.pgd_update(delta.detach(), grad, cfg["adv_lr_img"], cfg["adv_max_norm_img"])
    return delta


@pytest.mark.parametrize("space", ["embed", "embed_image"])
def test_embed_space_attack_matches_oracle_loop(golden, space):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    mod, g, cfg = _module(golden, space)
    batch = _batch(g)
    k_hat = F.normalize(torch.randn(batch["image"][0].shape[0], g.i("meta/C"), generator=torch.Generator().manual_seed(5)), dim=1)
    SamplingTransformer.calls = 0
    delta = mod.pgd_attacker.pgd_attack(mod, batch, k_modality=k_hat.to(DEV))
    assert SamplingTransformer.calls == 1
    base, masks = mod.pgd_attacker.embed_base, mod.pgd_attacker.embed_masks
    n_txt = batch["text_ids"].shape[1]
    assert delta.shape == (base.shape[0], (n_txt if space == "embed" else 0) + MAX_IMAGE_LEN + 1, base.shape[2])
    mod_cpu = _module(golden, space, device="cpu")[0]          # same golden state, same queue seed
    batch_cpu = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in batch.items() if k != "image"}
    want = _oracle_loop(mod_cpu, batch_cpu, base.cpu(), masks.cpu(), k_hat, mod_cpu.proj_queue, mod.temperature, cfg, space)
    eps = cfg["adv_max_norm_img"]
    err = (delta.cpu() - want).abs().max().item()
    assert err <= 1e-3 * eps, f"max |delta - oracle| = {err:.3e} ({err / eps:.2e} eps)"
    assert 0 < delta.abs().max().item() <= eps + 1e-7
    if space == "embed":
        assert delta[:, :n_txt].abs().max().item() > 0 and delta[:, n_txt:].abs().max().item() > 0
    nz = want.abs() > 1e-2 * eps
    assert (torch.sign(delta.cpu())[nz] == torch.sign(want)[nz]).float().mean().item() >= 0.999


@pytest.mark.parametrize("space", ["embed", "embed_image"])
def test_embed_space_attacked_view_reuses_the_attacked_embeddings(golden, space):
    """compute_moco_contrastive with the embedding-space attacker: the attacked forward must see exactly
    base + delta with the masks of the attack's own visual_embed call (no re-sampling), text tokens perturbed
    through the text-embedding hook only during that forward."""
    import rmcl_b200
    mod, g, cfg = _module(golden, space)
    batch = _batch(g)
    seen = []
    plain_infer = mod.infer

    def spy(b, mask_text=False, mask_image=False, **kw):
        hooks = len(mod.text_embeddings._forward_hooks)
        seen.append((kw.get("image_embeds"), kw.get("image_masks"), hooks))
        return plain_infer(b, mask_text, mask_image, **kw)

    mod.infer = spy
    SamplingTransformer.calls = 0
    ret = rmcl_b200.compute_moco_contrastive(mod, batch)
    ret["moco_loss"].backward()
    torch.cuda.synchronize()
    # key forward, clean query forward, the attack: three calls; the attacked view: none
    assert SamplingTransformer.calls == 3
    att = mod.pgd_attacker
    clean_call, attacked_call = seen[0], seen[-1]
    assert clean_call[0] is None and clean_call[2] == 0
    assert attacked_call[1] is att.embed_masks
    assert attacked_call[2] == (1 if space == "embed" else 0)            # text hook active only for "embed"
    assert len(mod.text_embeddings._forward_hooks) == 0                  # ... and removed afterwards
    delta_norm = mod.logged["moco_attack/train/delta"]
    assert torch.isfinite(ret["moco_loss"]) and delta_norm.item() > 0
    n_txt = batch["text_ids"].shape[1] if space == "embed" else 0
    # recompute the attack's final delta from what the forward received
    d_img = attacked_call[0] - att.embed_base
    assert d_img.abs().max().item() <= cfg["adv_max_norm_img"] * (1 + 1e-5) and d_img.abs().max().item() > 0
    assert d_img.shape[1] == MAX_IMAGE_LEN + 1 and n_txt in (0, batch["text_ids"].shape[1])
    assert mod.moco_head.projector[0].weight.grad is not None
