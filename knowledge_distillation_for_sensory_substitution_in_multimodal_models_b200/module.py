"""A Lightning-compatible KD module shaped like the reference's ``OnlineKnowledgeDistillationLLavaOneVision``.

What it mirrors (paths relative to the reference root, P1 = .../phase1/OnlineKnowledgeDistillationLLavaOneVision.py,
TR2 = .../phase2/train_online_kd.py):

* the batch keys and the two model calls of ``forward`` (P1:204-240), the hooks on the vision towers'
  ``post_layernorm`` (P1:110-121), the phase switch (P1:250-260) and the hyper-parameters (P1:67-71);
* ``training_step`` / ``validation_step`` / ``configure_optimizers`` (P1:122-140, P1:196-202: AdamW + CosineAnnealingLR(T_max=10));
* ``freeze_student_vision_layers`` as called by TR2:105.

What it does differently, on purpose:

* every loss term runs on the B200 kernels (``losses.KDLossMethods``); the hooked ``[B', 729, 1152]`` features go to the fused
  pool + normalise + InfoNCE op un-pooled (no ``.mean(dim=1)`` / ``F.normalize`` pass in torch);
* models are INJECTED (any ``nn.Module`` with the HF calling convention — returns ``.logits`` and ``.loss``) instead of being
  downloaded in ``__init__`` (P1:30-46): no network here, and tests drive it with small stand-ins;
* label sanitisation: the data module pads labels with ``-100`` (DM:145-146), which the reference's ``compute_loca_loss`` cannot
  index with (P1:166 raises; its README runs use ``--batch_size 1``).  Rows whose label is outside ``[0, V)`` are dropped from
  the logit-KD term (``sanitize_labels=True``) — exactly the reference's result when nothing is padded;
* ``group=``: tokens sharded over a ``torch.distributed`` group (one process per GPU) — the loss of the concatenated batch.

It subclasses ``pytorch_lightning.LightningModule`` when that package is importable and ``torch.nn.Module`` otherwise, with the
same hook names, so ``pl.Trainer(...).fit(module, datamodule)`` and the plain loop of ``train_online_kd.py`` both work.
"""
from __future__ import annotations

import types

import torch

from . import functional as F_kd
from .losses import KDLossMethods

try:  # Lightning is optional (absent in this image)
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
    HAVE_LIGHTNING = isinstance(_Base, type) and _Base is not torch.nn.Module
except Exception:  # noqa: BLE001
    _Base = torch.nn.Module
    HAVE_LIGHTNING = False


def _resolve(module: torch.nn.Module, dotted: str):
    for part in dotted.split("."):
        module = getattr(module, part)
    return module


class OnlineKnowledgeDistillationModule(_Base, KDLossMethods):
    VISION_HOOK = "vision_tower.vision_model.post_layernorm"          # P1:110-111

    def __init__(self, student_model: torch.nn.Module, teacher_model: torch.nn.Module, processor=None, phase: int = 2,
                 learning_rate: float = 1e-5, T: float = 0.8, soft_target_loss_weight: float = 0.1, ce_loss_weight: float = 0.5,
                 gamma: float = 0.8, loca_mode: str = "reference", sanitize_labels: bool = True, vision_hook: str | None = VISION_HOOK,
                 group=None, exchange: str = "nccl", fuse_ce: bool = False):
        super().__init__()
        self.phase, self.learning_rate, self.processor = phase, learning_rate, processor
        self.student_model, self.teacher_model = student_model, teacher_model
        self.teacher_model.eval()
        for param in self.teacher_model.parameters():                  # P1:49-52
            param.requires_grad = False
        self.soft_target_loss_weight, self.ce_loss_weight, self.gamma, self.T = soft_target_loss_weight, ce_loss_weight, gamma, T
        self.loca_mode, self.sanitize_labels = loca_mode, sanitize_labels
        self.kd_group, self.kd_exchange = group, exchange
        # fuse_ce: the student is called WITHOUT labels and its cross-entropy comes out of the logit-KD kernel's pass over the
        # student logits — as-written LoCa mode, phases 2 / 3; with a group: kd_loca_emit_ce on the token shards
        self.fuse_ce = bool(fuse_ce) and loca_mode == "reference"
        self.student_representation_output = None
        self.teacher_representation_output = None
        if vision_hook:
            self._register_hooks(vision_hook)

    # ---- P1:98-121 -----------------------------------------------------------------------------------------------
    def _register_hooks(self, dotted: str) -> None:
        try:
            teacher_layer, student_layer = _resolve(self.teacher_model, dotted), _resolve(self.student_model, dotted)
        except AttributeError:
            return                                  # stand-in models without a vision tower: phases 2 / 3 need no features
        teacher_layer.register_forward_hook(self.hook_fn_teacher)
        student_layer.register_forward_hook(self.hook_fn_student)

    def hook_fn_student(self, module, input, output):  # noqa: A002 - signature of a forward hook
        self.student_representation_output = output

    def hook_fn_teacher(self, module, input, output):  # noqa: A002
        self.teacher_representation_output = output

    def setup(self, stage=None):                                        # P1:82-85
        self.student_model.train()
        self.teacher_model.eval()

    # ---- P1:204-262 ----------------------------------------------------------------------------------------------
    def forward(self, batch):
        labels = batch["labels"]
        extra = {"image_sizes": batch["image_sizes"]} if "image_sizes" in batch else {}
        with torch.no_grad():
            teacher_outputs = self.teacher_model(input_ids=batch["rgb_input_ids"], pixel_values=batch["rgb_pixel_values"],
                                                 labels=labels, **extra)
        fused = self.fuse_ce and self.phase in (2, 3)
        student_outputs = self.student_model(input_ids=batch["depth_input_ids"], pixel_values=batch["depth_pixel_values"],
                                             labels=None if fused else labels, **extra)
        teacher_logits = teacher_outputs.logits
        if fused:
            return self._fused_logit_phase(teacher_logits, student_outputs.logits, labels)
        if self.phase == 1:
            # un-pooled [B', tokens, D] features: pooling + normalisation + loss + gradient broadcast are one fused op
            return self.compute_vision_loss(self.student_representation_output, self.teacher_representation_output.detach(),
                                            teacher_logits, student_outputs)
        s_logits, t_logits, lab = self._valid_rows(student_outputs.logits, teacher_logits, labels)
        loca_loss = self.compute_loca_loss(t_logits, s_logits, student_outputs.loss, lab)
        if self.phase == 2:
            return loca_loss
        return self.gamma * loca_loss + (1 - self.gamma) * student_outputs.loss        # P1:257-260

    def _fused_logit_phase(self, teacher_logits, student_logits, labels):
        """Phases 2 / 3 with the cross-entropy fused: loca + CE (P1:254), or gamma * (loca + CE) + (1 - gamma) * CE
        = gamma * (loca + CE / gamma) (P1:257-260), from one pass over the logits.  The CE is the mean over ALL valid next-token
        targets; rows LoCa cannot use (label outside [0, V), e.g. the -100 of a prompt token) but whose NEXT token is a target
        are few (one per answer start) and go through a plain cross-entropy on just those rows.  With ``group=``: KD and CE are
        those of the batch concatenated over the ranks (the valid-target count is summed on the device); the value returned is that
        global loss plus this rank's share of the few left-over rows, the gradient is the local rows'."""
        V = student_logits.shape[-1]
        ce_labels = F_kd.shift_labels(labels).reshape(-1)
        flat = labels.reshape(-1)
        n_valid = (ce_labels != F_kd.CE_IGNORE).sum()
        group = self.kd_group
        if group is not None:
            import torch.distributed as dist
            n_valid = n_valid.to(torch.float32)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(n_valid, op=dist.ReduceOp.SUM, group=group)
        gamma = self.gamma if self.phase == 3 else 1.0
        s2, t2 = student_logits.reshape(-1, V), teacher_logits.reshape(-1, teacher_logits.shape[-1])
        extra = None
        if self.sanitize_labels:
            ok = (flat >= 0) & (flat < V)
            if not bool(ok.all()):
                idx = ok.nonzero(as_tuple=True)[0]
                rest = ((~ok) & (ce_labels != F_kd.CE_IGNORE)).nonzero(as_tuple=True)[0]
                if rest.numel():
                    extra = torch.nn.functional.cross_entropy(s2.index_select(0, rest).float(), ce_labels.index_select(0, rest),
                                                              reduction="sum") / n_valid
                s2, t2, flat, ce_labels = s2.index_select(0, idx), t2.index_select(0, idx), flat.index_select(0, idx), ce_labels.index_select(0, idx)
        up = gamma if gamma != 1.0 else None
        if group is not None:
            from . import sharded
            total, ce = sharded.loca_kl_ce_token_sharded_autograd(t2, s2, flat, ce_labels, self.T, ce_weight=1.0 / gamma, group=group,
                                                                  upstream=up, exchange=self.kd_exchange, ce_valid_count=n_valid,
                                                                  phases=getattr(self, "kd_phases", None))
        else:
            total, ce = F_kd.loca_kl_ce(t2, s2, flat, ce_labels, self.T, ce_weight=1.0 / gamma, upstream=up, ce_valid_count=n_valid)
        self.last_student_loss = ce if extra is None else ce + extra.detach()
        loss = gamma * total
        return loss if extra is None else loss + extra

    def _valid_rows(self, student_logits, teacher_logits, labels):
        """Drop the token rows whose label cannot be indexed (pads are -100, DM:145-146).  No copy when every label is valid —
        the reference's own setting (batch_size 1)."""
        V = student_logits.shape[-1]
        flat = labels.reshape(-1)
        if not self.sanitize_labels:
            return student_logits, teacher_logits, labels
        ok = (flat >= 0) & (flat < V)
        if bool(ok.all()):                      # one host read-back per step; Lightning's logging reads the loss anyway
            return student_logits, teacher_logits, labels
        idx = ok.nonzero(as_tuple=True)[0]
        s2 = student_logits.reshape(-1, V).index_select(0, idx)
        t2 = teacher_logits.reshape(-1, teacher_logits.shape[-1]).index_select(0, idx)
        return s2.unsqueeze(0), t2.unsqueeze(0), flat.index_select(0, idx).unsqueeze(0)

    # ---- P1:122-140, 196-202 -------------------------------------------------------------------------------------
    def _log(self, name, value, **kw):
        if HAVE_LIGHTNING and getattr(self, "_trainer", None) is not None:
            self.log(name, value, **kw)

    def training_step(self, batch, batch_idx):
        loss = self(batch)
        self._log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True, logger=True)
        return loss

    def validation_step(self, batch, batch_idx):
        loss = self(batch)
        self._log("val_loss", loss, on_step=False, on_epoch=True, prog_bar=True, logger=True)
        return loss

    def configure_optimizers(self):
        params = [p for p in self.student_model.parameters() if p.requires_grad]
        optimizer = torch.optim.AdamW(params, lr=self.learning_rate)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=10)
        return [optimizer], [scheduler]

    def freeze_student_vision_layers(self):                                  # called by TR2:105
        tower = getattr(self.student_model, "vision_tower", None)
        if tower is not None:
            for param in tower.parameters():
                param.requires_grad = False


class TinyCausalLM(torch.nn.Module):
    """Stand-in with the HF calling convention (``.logits`` / ``.loss`` with the shifted, ignore_index=-100 CE of
    P1:238-240) and a ``vision_tower.vision_model.post_layernorm`` to hook — for tests and for ``train_online_kd.py --synthetic``."""

    def __init__(self, vocab: int = 2048, hidden: int = 64, vision_dim: int = 64, dtype=torch.bfloat16, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.embed = torch.nn.Embedding(vocab, hidden)
        self.mix = torch.nn.Linear(hidden, hidden)
        self.lm_head = torch.nn.Linear(hidden, vocab, bias=False)
        vm = torch.nn.Module()
        vm.post_layernorm = torch.nn.LayerNorm(vision_dim)
        self.vision_tower = torch.nn.Module()
        self.vision_tower.vision_model = vm
        self.vision_in = torch.nn.Linear(vision_dim, vision_dim)
        self.vision_to_text = torch.nn.Linear(vision_dim, hidden)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / max(p.shape[-1], 1) ** 0.5) + (1.0 if p.dim() == 1 and p is vm.post_layernorm.weight else 0.0))
        self.out_dtype = dtype

    def forward(self, input_ids, pixel_values, labels=None, **_):
        feats = self.vision_tower.vision_model.post_layernorm(self.vision_in(pixel_values.float()))      # [B', tokens, D] (hooked)
        ctx = self.vision_to_text(feats.mean(dim=1)).mean(dim=0)
        h = torch.tanh(self.mix(self.embed(input_ids) + ctx))
        logits = self.lm_head(h).to(self.out_dtype)
        loss = None
        if labels is not None:
            shift = logits[..., :-1, :].float().reshape(-1, logits.shape[-1])
            loss = torch.nn.functional.cross_entropy(shift, labels[..., 1:].reshape(-1), ignore_index=-100)
        return types.SimpleNamespace(logits=logits, loss=loss)
