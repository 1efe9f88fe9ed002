"""rmcl_b200 — B200-native kernels for the RMCL contrastive-adversarial training step.

Public surface (reference names):
  compute_moco_contrastive, compute_pgd          vilt/modules/objectives.py:217-447, 160-188
  compute_barlowtwins_contrastive                vilt/modules/objectives.py:449-602 (fused cross-correlation loss)
  PGDAttack, PGDAttack_moco                      attack/pgd_attack_vilt.py:7-175
  PGDAttack_{bartlowtwins,nlvr2,irtr,vqa}        attack/pgd_attack_vilt.py:178-483 (same update kernel)
  MoCo, concat_all_gather                        MoCo/MoCo_RMCL.py
  ops.{ema_multi_, infonce_fwd_bwd, infonce_loss, enqueue_, pgd_step_, barlow_fwd_bwd, barlow_twins_loss}   the kernels themselves
"""
from . import greedy, ops  # noqa: F401
from .dist import concat_all_gather  # noqa: F401
from .moco import MoCo  # noqa: F401
from .objectives import (compute_barlowtwins_contrastive, compute_moco_contrastive, compute_pgd,  # noqa: F401
                         dequeue_and_enqueue, momentum_update_key_encoder, shadow_layer)
from .pgd_attack import (PGDAttack, PGDAttack_bartlowtwins, PGDAttack_irtr, PGDAttack_moco,  # noqa: F401
                         PGDAttack_nlvr2, PGDAttack_vqa)

__version__ = "0.1.0"
