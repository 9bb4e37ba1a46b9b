"""Host-side mirror of the reference's loss-method interface.

The reference has no plugin layer: the boundary is a set of methods on the
``OnlineKnowledgeDistillationLLavaOneVision`` LightningModule, called from ``forward``
(P1:250-260, LG:165, FB:162-165, EV:236).  :class:`KDLossMethods` carries methods with the SAME
names, argument order and return conventions, backed by the CUDA kernels; :func:`patch_reference`
rebinds them onto a reference class (or instance) so its unmodified call sites run on the B200 path.
"""
from __future__ import annotations

import torch

from . import functional as F_kd


class KDLossMethods:
    """Mixin: expects the attributes the reference sets in ``__init__`` (P1:67-71, LG:73-75, FB:72-74):
    ``T``, ``soft_target_loss_weight``, ``ce_loss_weight`` (and ``gamma``/``phase`` for ``combine_phase``).
    ``loca_mode`` ("reference" | "per_row") selects the LoCa semantics; the default is the as-written one."""

    loca_mode = "reference"
    kd_group = None          # torch.distributed group the tokens are sharded over (None = this process holds the whole batch)
    kd_exchange = "nccl"     # record exchange of the as-written mode: "nccl" (all-gather) or "peer" (stores from phase A over NVLink)

    # P1:141-194 / LG:208-261 — returns loca + student_loss like the reference (P1:194)
    def compute_loca_loss(self, teacher_logits, student_logits, student_loss, labels, alpha=0.8):
        """Same signature as the reference.  ``student_loss=None`` asks for the FUSED form (:meth:`compute_loca_loss_fused`)."""
        mode = getattr(self, "loca_mode", "reference")
        group = getattr(self, "kd_group", None)
        if student_loss is None:
            return self.compute_loca_loss_fused(teacher_logits, student_logits, labels, alpha=alpha)
        if group is not None:
            # tokens sharded over the ranks of `group`: the loss is the one of the CONCATENATED batch (one 16 B/token exchange in
            # the as-written mode, none otherwise), the gradient is this rank's rows'
            from . import sharded
            if mode == "reference":
                loca = sharded.loca_kl_token_sharded_autograd(teacher_logits, student_logits, labels, self.T, alpha=alpha, group=group,
                                                              exchange=getattr(self, "kd_exchange", "nccl"),
                                                              phases=getattr(self, "kd_phases", None))
            else:
                loca = sharded.per_row_token_sharded(F_kd.loca_kl, teacher_logits, student_logits, labels, self.T, alpha=alpha,
                                                     mode="per_row", group=group)
        else:
            loca = F_kd.loca_kl(teacher_logits, student_logits, labels, self.T, alpha=alpha, mode=mode)
        return loca + student_loss

    # SURVEY §8(f) rank 1 — P1:194 with the student's own loss (P1:238-240) computed in the same pass over its logits
    def compute_loca_loss_fused(self, teacher_logits, student_logits, labels, alpha=0.8, ce_labels=None, ce_weight=1.0, upstream=None):
        """``loca + ce_weight * CE`` where CE is what ``student_outputs.loss`` would have been — call the student with
        ``labels=None`` so it is not computed twice.  ``ce_labels`` defaults to the next-token shift of ``labels``; the CE alone
        is left in ``self.last_student_loss`` (detached) for logging / the phase-3 mix."""
        if getattr(self, "loca_mode", "reference") != "reference":
            raise NotImplementedError("the fused cross-entropy rides the as-written LoCa kernel; pass student_loss for per_row calls")
        if ce_labels is None:
            ce_labels = F_kd.shift_labels(labels)
        group = getattr(self, "kd_group", None)
        if group is not None:
            # token shards: KD and CE are those of the concatenated batch (the CE's divisor is summed over the ranks on the device)
            from . import sharded
            total, ce = sharded.loca_kl_ce_token_sharded_autograd(teacher_logits, student_logits, labels, ce_labels, self.T,
                                                                  alpha=alpha, ce_weight=ce_weight, group=group, upstream=upstream,
                                                                  exchange=getattr(self, "kd_exchange", "nccl"),
                                                                  phases=getattr(self, "kd_phases", None))
            self.last_student_loss = ce
            return total
        total, ce = F_kd.loca_kl_ce(teacher_logits, student_logits, labels, ce_labels, self.T, alpha=alpha, ce_weight=ce_weight,
                                    upstream=upstream)
        self.last_student_loss = ce
        return total

    # P1:316-354 / EV:246-283 — features arrive pooled + normalised (P1:243-248)
    def compute_vision_loss(self, student_features, teacher_features, teacher_logits, student_outputs):
        kl = F_kd.temperature_kl(teacher_logits, student_outputs.logits, self.T)
        contrastive = self.contrastive_loss(student_features, teacher_features)
        return self.soft_target_loss_weight * kl + self.ce_loss_weight * contrastive

    # FB:191-230 — KL with the log_target quirk + CE + contrastive
    def compute_loss(self, teacher_logits, student_outputs, contrastive_loss_value):
        kl = F_kd.temperature_kl(teacher_logits, student_outputs.logits, self.T, log_target_quirk=True)
        return (self.soft_target_loss_weight * kl) + (self.ce_loss_weight * student_outputs.loss) + contrastive_loss_value

    # P1:356-390 — sum / size(0) normaliser (no live caller in the reference; kept for completeness)
    def compute_language_loss(self, teacher_logits, student_outputs):
        kl = F_kd.temperature_kl(teacher_logits, student_outputs.logits, self.T, reduction="batchmean")
        return (self.soft_target_loss_weight * kl) + (self.ce_loss_weight * student_outputs.loss)

    # P1:393-416 (copies LG:307-330, FB:288-311, EV:327-350)
    def contrastive_loss(self, student_features, teacher_features, temperature=0.07):
        return F_kd.info_nce(student_features, teacher_features, temperature)

    # EV:160-165
    def mse_loss(self, input, target):  # noqa: A002 - name kept from the reference
        return F_kd.feature_mse(input, target)

    # P1:250-260 — the phase switch of forward(), given model outputs
    def combine_phase(self, teacher_logits, student_outputs, labels, student_features=None, teacher_features=None):
        phase = getattr(self, "phase", 2)
        if phase == 1:
            return self.compute_vision_loss(student_features, teacher_features, teacher_logits, student_outputs)
        if phase == 2:
            return self.compute_loca_loss(teacher_logits, student_outputs.logits, student_outputs.loss, labels)
        if phase == 3:
            loca = self.compute_loca_loss(teacher_logits, student_outputs.logits, student_outputs.loss, labels)
            return self.gamma * loca + (1 - self.gamma) * student_outputs.loss
        raise ValueError(f"unknown phase {phase}")


class KDLoss(torch.nn.Module, KDLossMethods):
    """Stand-alone holder of the reference's hyper-parameters (defaults of P1:67-71)."""

    def __init__(self, T: float = 0.8, soft_target_loss_weight: float = 0.1, ce_loss_weight: float = 0.5,
                 gamma: float = 0.8, phase: int = 2, loca_mode: str = "reference"):
        super().__init__()
        self.T, self.soft_target_loss_weight, self.ce_loss_weight = T, soft_target_loss_weight, ce_loss_weight
        self.gamma, self.phase, self.loca_mode = gamma, phase, loca_mode

    def forward(self, teacher_logits, student_outputs, labels, student_features=None, teacher_features=None):
        return self.combine_phase(teacher_logits, student_outputs, labels, student_features, teacher_features)


_PATCHABLE = ("compute_loca_loss", "compute_vision_loss", "compute_loss", "compute_language_loss",
              "contrastive_loss", "mse_loss")


def patch_reference(target, loca_mode: str = "reference", only_existing: bool = True, group=None, exchange: str = "nccl"):
    """Rebind the reference's loss methods on ``target`` (a class or an instance) to the CUDA-backed ones.

    Only methods the target already defines are replaced unless ``only_existing=False``.  Returns the
    list of patched names.  The reference's call sites (P1:251/254/258, LG:165, FB:162/165, EV:236) stay untouched.
    ``group``: a ``torch.distributed`` process group over which the TOKENS of a step are sharded (one process per GPU):
    ``compute_loca_loss`` then returns the loss of the concatenated batch and differentiates w.r.t. the local rows
    (``sharded.loca_kl_token_sharded_autograd``); ``exchange`` picks the record exchange of the as-written mode.
    """
    import types

    patched = []
    is_class = isinstance(target, type)
    for name in _PATCHABLE:
        if only_existing and not hasattr(target, name):
            continue
        fn = getattr(KDLossMethods, name)
        if is_class:
            setattr(target, name, fn)
        else:
            setattr(target, name, types.MethodType(fn, target))
        patched.append(name)
    # helper the patched compute_loca_loss(student_loss=None) forwards to (not a reference method: always attached)
    fused = KDLossMethods.compute_loca_loss_fused
    setattr(target, "compute_loca_loss_fused", fused if is_class else types.MethodType(fused, target))
    setattr(target, "loca_mode", loca_mode)
    setattr(target, "kd_group", group)
    setattr(target, "kd_exchange", exchange)
    return patched
