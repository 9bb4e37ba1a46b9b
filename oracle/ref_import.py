"""TEST INFRASTRUCTURE ONLY — import the reference's own loss methods.

Works only where ``/root/reference`` exists (the dev container; never the GPU
box).  Used by ``oracle/make_goldens.py`` to freeze golden vectors and by the
``not gpu`` tests (skipped when the reference tree is absent) to pin
``oracle/kd_oracle.py`` and ``oracle/reference_port.py``.

The KD LightningModules import ``pytorch_lightning`` (absent here); the loss
methods only touch ``self.T`` & friends, so a stub module whose
``LightningModule`` is ``torch.nn.Module`` is enough and the methods are
called unbound on a ``SimpleNamespace`` (recipe: SURVEY.md Appendix B).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("KD_REFERENCE_ROOT", "/root/reference")

_FILES = {
    # abbreviations follow SURVEY.md
    "P1": "distillation/knowledge_distillation7b_double_trouble/phase1/OnlineKnowledgeDistillationLLavaOneVision.py",
    "LG": "distillation/knowledge_distillation7b_logit_based/OnlineKnowledgeDistillationLLavaOneVision.py",
    "FB": "distillation/knowledge_distillation7b_feature_based/OnlineKnowledgeDistillationLLavaOneVision.py",
    "EV": "evaluation/onevisionv3/OnlineKnowledgeDistillationLLavaOneVision.py",
}
_cache: dict[str, type] = {}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in _FILES.values())


def _install_stub() -> None:
    if "pytorch_lightning" in sys.modules:
        return
    stub = types.ModuleType("pytorch_lightning")
    stub.LightningModule = torch.nn.Module
    stub.LightningDataModule = object
    sys.modules["pytorch_lightning"] = stub


def reference_class(which: str = "P1") -> type:
    """Return the reference ``OnlineKnowledgeDistillationLLavaOneVision`` class of file ``which``."""
    if which in _cache:
        return _cache[which]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, _FILES[which])
    spec = importlib.util.spec_from_file_location(f"_kd_reference_{which}", path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):  # the module prints its cwd on import
        spec.loader.exec_module(mod)
    cls = mod.OnlineKnowledgeDistillationLLavaOneVision
    _cache[which] = cls
    return cls


def fake_self(**attrs) -> types.SimpleNamespace:
    """A stand-in for the LightningModule instance: the loss methods read only attributes."""
    ns = types.SimpleNamespace(**attrs)
    return ns


# ---- thin callables: (tensors, scalars) -> loss, exactly through the reference code ----

def loca_loss(teacher, student, labels, T, alpha=0.8, student_loss=None, which="P1"):
    """P1:141-194 / LG:208-261. ``teacher``/``student`` are [B,S,V]; ``labels`` [B,S] int64."""
    K = reference_class(which)
    if student_loss is None:
        student_loss = torch.zeros((), dtype=student.dtype)
    return K.compute_loca_loss(fake_self(T=T), teacher, student, student_loss, labels, alpha)


def vision_loss(student_feats, teacher_feats, teacher_logits, student_logits, T,
                soft_w=0.1, ce_w=0.5, which="P1"):
    """P1:316-354 / EV:246-283 (plain temperature KL + contrastive)."""
    K = reference_class(which)
    me = fake_self(T=T, soft_target_loss_weight=soft_w, ce_loss_weight=ce_w)
    me.contrastive_loss = types.MethodType(K.contrastive_loss, me)
    outs = types.SimpleNamespace(logits=student_logits, loss=None)
    return K.compute_vision_loss(me, student_feats, teacher_feats, teacher_logits, outs)


def feature_based_loss(teacher_logits, student_logits, student_ce, contrastive_value, T,
                       soft_w=0.1, ce_w=0.8):
    """FB:191-230 (``log_target=True`` KL quirk + CE + contrastive)."""
    K = reference_class("FB")
    me = fake_self(T=T, soft_target_loss_weight=soft_w, ce_loss_weight=ce_w)
    outs = types.SimpleNamespace(logits=student_logits, loss=student_ce)
    return K.compute_loss(me, teacher_logits, outs, contrastive_value)


def contrastive_loss(student_feats, teacher_feats, temperature=0.07, which="P1"):
    """P1:393-416 and copies."""
    K = reference_class(which)
    return K.contrastive_loss(fake_self(), student_feats, teacher_feats, temperature)


def mse_loss(inp, target):
    """EV:160-165."""
    K = reference_class("EV")
    return K.mse_loss(fake_self(), inp, target)
