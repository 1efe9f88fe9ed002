"""PGD image attack with the reference's call signature, running on the fused kernels.

Mirrors attack/pgd_attack_vilt.py: ``PGDAttack`` (7-106) and ``PGDAttack_moco`` (109-175).
Same constructor keys (``adv_steps_img``, ``adv_lr_img``, ``adv_max_norm_img``,
``max_image_len``), same ``pgd_attack(pl_module, batch, k_modality=None) -> delta``, same side
effect on ``batch['image'][0]``.  What changes underneath:
  * the inner loss is the fused InfoNCE (``ops.InfoNCE``), not einsum/cat/CrossEntropy;
  * the update is one ``rmcl_pgd_step`` launch instead of 7 elementwise/reduce kernels;
  * by default no 112 M-parameter deepcopy per call and no weight gradients: the gradient is
    taken w.r.t. the perturbation only (``torch.autograd.grad``).  ``copy_modules=True``
    restores the reference's deepcopy (identical numbers, just slower).
Extensions (defaults = reference behaviour): ``mode`` in {"ref_linf","sign_linf","l2"},
``space`` in {"pixel","embed"}.
"""
from copy import deepcopy

import torch

from . import ops


class PGDAttack:
    def __init__(self, config, contrastive_framework):
        self.contrastive_framework = contrastive_framework
        self.adv_steps_img = config["adv_steps_img"]
        self.adv_lr_img = config["adv_lr_img"]
        self.adv_max_norm_img = config["adv_max_norm_img"]
        self.max_image_len = config["max_image_len"]
        self.pl_module = None
        self.text_embeddings = None
        self.transformer = None
        self.token_type_embeddings = None
        self.pooler = None

    def build_mini_vilt(self, pl_module):
        raise NotImplementedError(f"Build_mini_vilt of {self.contrastive_framework} isn't implemented.")

    def vilt_zero_grad(self):
        raise NotImplementedError(f"vilt_zero_grad of {self.contrastive_framework} isn't implemented.")

    def pgd_attack(self, pl_module, batch, k_image):
        raise NotImplementedError(f"pgd_attack of {self.contrastive_framework} isn't implemented.")

    def infer(self, batch, mask_text=False, mask_image=False, image_token_type_idx=1, image_embeds=None,
              image_masks=None):
        """The mini-ViLT forward the attack differentiates through (same contract as the
        reference's PGDAttack.infer: text+image embeddings -> blocks -> norm -> pooler)."""
        imgkey = f"image_{image_token_type_idx - 1}"
        if imgkey not in batch:
            imgkey = "image"
        suffix = "_mlm" if mask_text else ""
        text_ids, text_labels, text_masks = batch[f"text_ids{suffix}"], batch[f"text_labels{suffix}"], batch["text_masks"]
        text_embeds = self.text_embeddings(text_ids)
        patch_index = image_labels = None
        if image_embeds is None and image_masks is None:
            image_embeds, image_masks, patch_index, image_labels = self.transformer.visual_embed(
                batch[imgkey][0], max_image_len=self.max_image_len, mask_it=mask_image)
        text_embeds = text_embeds + self.token_type_embeddings(torch.zeros_like(text_masks))
        image_embeds = image_embeds + self.token_type_embeddings(torch.full_like(image_masks, image_token_type_idx))
        x = torch.cat([text_embeds, image_embeds], dim=1)
        co_masks = torch.cat([text_masks, image_masks], dim=1)
        for blk in self.transformer.blocks:
            x, _ = blk(x, mask=co_masks)
        x = self.transformer.norm(x)
        n_text = text_embeds.shape[1]
        return {
            "text_feats": x[:, :n_text], "image_feats": x[:, n_text:],
            "cls_feats": self.pooler(x) if self.pooler is not None else None, "raw_cls_feats": x[:, 0],
            "image_labels": image_labels, "image_masks": image_masks, "text_labels": text_labels,
            "text_ids": text_ids, "text_masks": text_masks, "patch_index": patch_index,
        }


class PGDAttack_moco(PGDAttack):
    def __init__(self, config, mode="ref_linf", space="pixel", copy_modules=False, infonce_path="auto"):
        super().__init__(config, "moco")
        self.moco_head = None
        self.mode, self.space, self.copy_modules, self.infonce_path = mode, space, copy_modules, infonce_path

    def build_mini_vilt(self, pl_module):
        grab = deepcopy if self.copy_modules else (lambda m: m)
        self.pl_module = pl_module
        self.text_embeddings = grab(pl_module.text_embeddings)
        self.token_type_embeddings = grab(pl_module.token_type_embeddings)
        self.transformer = grab(pl_module.transformer)
        self.moco_head = grab(pl_module.moco_head)
        self.pooler = grab(pl_module.pooler)

    def vilt_zero_grad(self):
        if not self.copy_modules:
            return  # nothing accumulates: gradients are taken w.r.t. the perturbation only
        for m in (self.text_embeddings, self.transformer, self.token_type_embeddings, self.moco_head, self.pooler):
            m.zero_grad()

    def pgd_attack(self, pl_module, batch, k_modality=None):
        self.build_mini_vilt(pl_module)
        self.vilt_zero_grad()
        queue, temperature = pl_module.proj_queue, pl_module.temperature
        img_init = batch["image"][0]
        if self.space == "embed":
            with torch.no_grad():
                base, image_masks, _, _ = self.transformer.visual_embed(
                    img_init, max_image_len=self.max_image_len, mask_it=False)
        else:
            base, image_masks = img_init, None
        delta = torch.zeros_like(base)
        for _ in range(self.adv_steps_img):
            delta.requires_grad_(True)
            with torch.autocast("cuda", enabled=False), torch.enable_grad():
                if self.space == "embed":
                    infer = self.infer(batch, image_embeds=base + delta, image_masks=image_masks)
                else:
                    batch["image"][0] = img_init + delta  # reference side effect (pgd_attack_vilt.py:144)
                    infer = self.infer(batch)
                q_raw = self.moco_head(infer["cls_feats"])
                loss, _ = ops.infonce_loss(q_raw.float(), k_modality, queue, temperature, self.infonce_path)
                loss = loss / (1.0 * self.adv_steps_img)
                (grad,) = torch.autograd.grad(loss, delta)
            delta = delta.detach()
            ops.pgd_step_(delta, grad.contiguous(), self.adv_lr_img, self.adv_max_norm_img, self.mode)
        return delta
