#!/usr/bin/env python
"""Training entry point mirroring the reference's ``distillation/knowledge_distillation7b_double_trouble/phase{1,2,3}/train_online_kd.py``
(TR2:62-73 flags, TR2:140-153 trainer settings) on the B200 KD-loss kernels.

    python train_online_kd.py --phase 2 --batch_size 1 --max_epochs 10 --accumulate_grad_batches 64 [--synthetic]

With real checkpoints on disk pass ``--student`` / ``--teacher`` (HF ``LlavaOnevisionForConditionalGeneration`` directories; the
reference hard-codes the hub names at TR2:77-79) and a datamodule factory ``--datamodule pkg.mod:factory`` yielding the
reference's batch keys (``rgb_input_ids``, ``rgb_pixel_values``, ``depth_input_ids``, ``depth_pixel_values``, ``labels``,
``image_sizes``).  ``--synthetic`` runs the same loop on small stand-in models and random batches (no network in this image).
If ``pytorch_lightning`` is installed the module is handed to ``pl.Trainer`` with the reference's settings; otherwise the
equivalent plain loop below runs (AdamW + cosine schedule from ``configure_optimizers``, gradient accumulation, bf16 autocast).
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.module import (  # noqa: E402
    HAVE_LIGHTNING, OnlineKnowledgeDistillationModule, TinyCausalLM)


def synthetic_batches(n, batch_size, seq, vocab, tokens, vdim, device, pad_last=True):
    g = torch.Generator().manual_seed(0)
    for _ in range(n):
        ids = torch.randint(0, vocab, (batch_size, seq), generator=g)
        labels = ids.clone()
        if pad_last and batch_size > 1:
            labels[-1, -seq // 4:] = -100                      # padded tail, as DM:145-146 produces for batch_size > 1
        px = torch.randn(batch_size, tokens, vdim, generator=g)
        yield {"rgb_input_ids": ids.to(device), "depth_input_ids": ids.to(device), "rgb_pixel_values": px.to(device),
               "depth_pixel_values": (px + 0.3 * torch.randn(px.shape, generator=g)).to(device), "labels": labels.to(device)}


def main(argv=None):
    ap = argparse.ArgumentParser(description="Online KD training on the B200 KD-loss kernels")
    ap.add_argument("--batch_size", type=int, default=1)                       # TR2:66
    ap.add_argument("--max_epochs", type=int, default=10)                      # TR2:67
    ap.add_argument("--subset_percentage", type=float, default=1)              # TR2:68
    ap.add_argument("--accumulate_grad_batches", type=int, default=64)         # TR2:71
    ap.add_argument("--phase", type=int, default=2, choices=[1, 2, 3])
    ap.add_argument("--loca_mode", default="reference", choices=["reference", "per_row"])
    ap.add_argument("--learning_rate", type=float, default=1e-5)
    ap.add_argument("--fuse_ce", action="store_true",
                    help="phases 2 / 3: call the student without labels and take its cross-entropy from the logit-KD kernel's own pass")
    ap.add_argument("--student", default=None)
    ap.add_argument("--teacher", default=None)
    ap.add_argument("--datamodule", default=None, help="pkg.module:factory returning an object with train_dataloader()")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--steps_per_epoch", type=int, default=8, help="--synthetic only")
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("the KD-loss kernels are CUDA-only (sm_100a)")
    dev = torch.device("cuda")

    if args.synthetic or not (args.student and args.teacher):
        vocab = 2048
        student, teacher = TinyCausalLM(vocab, 64, seed=1).to(dev), TinyCausalLM(vocab, 128, seed=2).to(dev)
        loaders = lambda: synthetic_batches(args.steps_per_epoch, args.batch_size, 64, vocab, 9, 64, dev)  # noqa: E731
        processor = None
    else:
        from transformers import AutoProcessor, LlavaOnevisionForConditionalGeneration
        processor = AutoProcessor.from_pretrained(args.teacher)
        student = LlavaOnevisionForConditionalGeneration.from_pretrained(args.student, low_cpu_mem_usage=True).to(dev)
        teacher = LlavaOnevisionForConditionalGeneration.from_pretrained(args.teacher, low_cpu_mem_usage=True, torch_dtype=torch.bfloat16).to(dev)
        if not args.datamodule:
            raise SystemExit("--datamodule pkg.module:factory is required with real checkpoints")
        mod, fn = args.datamodule.split(":")
        dm = getattr(importlib.import_module(mod), fn)(processor=processor, batch_size=args.batch_size, subset_percentage=args.subset_percentage)
        loaders = dm.train_dataloader

    model = OnlineKnowledgeDistillationModule(student, teacher, processor, phase=args.phase, learning_rate=args.learning_rate,
                                              loca_mode=args.loca_mode, fuse_ce=args.fuse_ce)
    model.freeze_student_vision_layers()                                        # TR2:105

    if HAVE_LIGHTNING and not args.synthetic:
        import pytorch_lightning as pl
        trainer = pl.Trainer(max_epochs=args.max_epochs, accelerator="gpu", precision="bf16-mixed", devices=1, num_sanity_val_steps=0,
                             accumulate_grad_batches=args.accumulate_grad_batches)      # TR2:140-153, bf16 instead of fp16 AMP
        trainer.fit(model, dm)
        return None

    (optimizer,), (scheduler,) = model.configure_optimizers()
    model.setup()
    history = []
    for epoch in range(args.max_epochs):
        optimizer.zero_grad(set_to_none=True)
        for i, batch in enumerate(loaders()):
            loss = model.training_step(batch, i) / args.accumulate_grad_batches
            loss.backward()
            if (i + 1) % args.accumulate_grad_batches == 0:
                optimizer.step()
                optimizer.zero_grad(set_to_none=True)
            history.append(loss.item() * args.accumulate_grad_batches)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
        scheduler.step()
        print(f"epoch {epoch}: train_loss {sum(history[-(i + 1):]) / (i + 1):.6f}", flush=True)
    return history


if __name__ == "__main__":
    main()
