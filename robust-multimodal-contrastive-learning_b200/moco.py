"""``MoCo`` — the module API sketched in MoCo/MoCo_RMCL.py, made executable on the fused kernels.

The reference file is not valid Python; it fixes the *names*:
``MoCo(config, dim, K, m, T)``, ``_momentum_update_key_encoder()``,
``_dequeue_and_enqueue(keys_txt, keys_img, dist)``, ``forward(batch, pgd_parameters, opts, device)``
and the module-level ``concat_all_gather``; buffers ``txt_img_queue`` [dim,K] and
``txt_img_queue_ptr`` int64[1] (MoCo_RMCL.py:24,49-51,65-94,96-175,268-279).  Numerics follow the
live implementation (objectives.py), which is what the parity tests pin.

Since the sketch's own encoder construction cannot run, encoders are injected: any pair of
modules whose ``forward(batch)`` returns ``(txt_rep, img_rep)`` raw projections [B,dim].
"""
import torch
import torch.nn as nn

from . import ops
from .dist import concat_all_gather


class MoCo(nn.Module):
    def __init__(self, config, dim=128, K=65536, m=0.999, T=0.07, encoder_q=None, encoder_k=None,
                 queue_dtype=torch.float32, infonce_path="auto"):
        super().__init__()
        self.config, self.K, self.m, self.T = config, K, m, T
        self.infonce_path = infonce_path
        self.encoder_q, self.encoder_k = encoder_q, encoder_k
        if encoder_q is not None and encoder_k is not None:
            for param_q, param_k in zip(encoder_q.parameters(), encoder_k.parameters()):
                param_k.data.copy_(param_q.data)   # MoCo_RMCL.py:43-46
                param_k.requires_grad = False
        self.register_buffer("txt_img_queue", torch.randn(dim, K).to(queue_dtype))   # un-normalised, like the sketch
        self.register_buffer("txt_img_queue_ptr", torch.zeros(1, dtype=torch.long))
        self._ema_plan = None

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        if self._ema_plan is None:
            self._ema_plan = ops.EmaPlan(list(self.encoder_k.parameters()), list(self.encoder_q.parameters()))
        ops.ema_multi_(self._ema_plan, self.m)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys_txt, keys_img, dist=True):
        """Text keys then image keys, pointer advancing twice (MoCo_RMCL.py:74-94)."""
        if dist:
            keys_txt, keys_img = concat_all_gather(keys_txt), concat_all_gather(keys_img)
        assert self.K % keys_txt.shape[0] == 0  # MoCo_RMCL.py:83
        ops.enqueue_(self.txt_img_queue, keys_txt, self.txt_img_queue_ptr)
        ops.enqueue_(self.txt_img_queue, keys_img, self.txt_img_queue_ptr)

    def forward(self, batch, pgd_parameters=None, opts=None, device=None, materialize_logits=False):
        """Cross-modal InfoNCE of the sketch (text query vs image key and vice versa,
        MoCo_RMCL.py:150-175).  Returns ``(out, labels, logs, None)`` where ``out`` holds fused
        ``loss_txt``/``loss_img`` (and, with ``materialize_logits=True``, the debug-only
        ``txt``/``img`` logits the sketch returned)."""
        with torch.no_grad():
            self._momentum_update_key_encoder()
            txt_k_raw, img_k_raw = self.encoder_k(batch)
        txt_q, img_q = self.encoder_q(batch)
        queue = self.txt_img_queue
        # the raw key projections go straight into the loss launches, which normalise them on the way
        # (nn.functional.normalize of MoCo_RMCL.py:126-127) and hand the unit keys back for the enqueue
        loss_txt, _, img_k = ops.infonce_loss(txt_q, img_k_raw.float(), queue, self.T, self.infonce_path,
                                              normalize_k=True, return_k_hat=True)
        loss_img, _, txt_k = ops.infonce_loss(img_q, txt_k_raw.float(), queue, self.T, self.infonce_path,
                                              normalize_k=True, return_k_hat=True)
        out = {"loss_txt": loss_txt, "loss_img": loss_img}
        labels = {n: torch.zeros(txt_q.shape[0], dtype=torch.long, device=txt_q.device) for n in ("txt", "img")}
        if materialize_logits:  # debug / parity only: this is exactly what the fused path avoids
            with torch.no_grad():
                for name, q, k in (("txt", txt_q, img_k), ("img", img_q, txt_k)):
                    qh = nn.functional.normalize(q.float(), dim=1)
                    out[name] = torch.cat([(qh * k).sum(1, keepdim=True), qh @ queue.float()], dim=1) / self.T
        self._dequeue_and_enqueue(txt_k, img_k, dist=torch.distributed.is_available() and torch.distributed.is_initialized())
        return out, labels, {}, None
