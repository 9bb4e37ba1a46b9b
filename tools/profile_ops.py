"""Driver for ncu: runs ONE op of the path a few times at its benchmark shape (the ncu command line picks the kernel).

usage: python tools/profile_ops.py <op> [rows]      op in: loca_reference loca_reference_ce loca_per_row loca_reference_T08 kl kl_log_target regrad
                                                          mse proj_mse info_nce info_nce_proj nce_tc gemm_dh gemm_dw lm_head vshard vshard_phases
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd  # noqa: E402
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded, synthetic  # noqa: E402

op = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
V = synthetic.VOCAB
dev = "cuda"
REPS = 3


def logits(peaked=False, boost=1.0):
    t, s, lab = synthetic.logit_batch(rows, V, seed=1234, device=dev, peaked=peaked)
    if boost != 1.0:
        s = (s.float() * boost).to(torch.bfloat16)
    return t, s.requires_grad_(True), lab


if op in ("loca_reference", "loca_reference_ce", "loca_per_row", "loca_reference_T08", "kl", "kl_log_target", "regrad"):
    peaked = op == "loca_reference_T08"
    t, s, lab = logits(peaked, 4.0 if peaked else 1.0)
    for _ in range(REPS):
        s.grad = None
        if op == "loca_reference":
            loss = kd.loca_kl(t, s, lab, 2.0)
        elif op == "loca_reference_ce":           # fused student cross-entropy, half of the rows carry a target
            g = torch.Generator().manual_seed(77)
            cel = torch.randint(0, V, (rows,), generator=g)
            cel[torch.rand(rows, generator=g) < 0.5] = -100
            loss = kd.loca_kl_ce(t, s, lab, cel.cuda(), 2.0)[0]
        elif op == "loca_per_row":
            loss = kd.loca_kl(t, s, lab, 2.0, mode="per_row")
        elif op == "loca_reference_T08":
            loss = kd.loca_kl(t, s, lab, 0.8)
        elif op == "kl":
            loss = kd.temperature_kl(t, s, 2.0)
        elif op == "kl_log_target":
            loss = kd.temperature_kl(t, s, 0.8, log_target_quirk=True)
        else:
            loss = kd.loca_kl(t, s, lab, 2.0) * 3.0          # grad_output != 1: the backward re-emission runs
        loss.backward()
elif op == "mse":
    a = torch.randn(65536, 3584, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    b = torch.randn(65536, 3584, device=dev, dtype=torch.bfloat16)
    for _ in range(REPS):
        a.grad = None
        loss = kd.feature_mse(a, b)
        loss.backward()
elif op == "proj_mse":
    hs, ht, w = synthetic.hidden_state_batch(65536, 896, 3584, device=dev)
    hs.requires_grad_(True)
    w.requires_grad_(True)
    for _ in range(REPS):
        hs.grad = None
        w.grad = None
        loss = kd.feature_mse(hs, ht, proj=w)
        loss.backward()
elif op == "info_nce":
    xs, xt = synthetic.vision_feature_batch(64, 729, 1152, device=dev)
    xs.requires_grad_(True)
    for _ in range(REPS):
        xs.grad = None
        loss = kd.info_nce(xs, xt)
        loss.backward()
elif op == "info_nce_proj":                       # BASELINE configs[2] with its learned 1152 -> 896 projection (one library call)
    xs, _ = synthetic.vision_feature_batch(64, 729, 1152, device=dev)
    xs.requires_grad_(True)
    wp = (torch.randn(1152, 896, device=dev) * 0.03).to(torch.bfloat16).requires_grad_(True)
    xt = torch.randn(64, 729, 896, device=dev).to(torch.bfloat16)
    for _ in range(REPS):
        xs.grad = None
        wp.grad = None
        loss = kd.info_nce(xs, xt, proj=wp)
        loss.backward()
elif op == "vshard_phases":                       # the streaming KD phases of one configs[4] shard, both tensors from HBM
    n, vl = 65536, 18992
    ts = (torch.randn(n, vl, device=dev) * 2.0).to(torch.bfloat16)
    ss = (torch.randn(n, vl, device=dev) * 2.0).to(torch.bfloat16)
    lab = torch.randint(0, vl, (n,), device=dev)
    for _ in range(2):
        loss, _g = sharded.logit_kd_vocab_sharded(ts, ss, lab, 0, V, 2.0)
elif op == "nce_tc":                              # token-level contrastive loss: 64 x 729 un-pooled SigLIP tokens as rows
    M = rows if len(sys.argv) > 2 else 46656
    xs, xt = synthetic.vision_feature_batch(M, 1, 1152, seed=7, device=dev)
    xs, xt = xs[:, 0].contiguous().requires_grad_(True), xt[:, 0].contiguous()
    for _ in range(REPS):
        xs.grad = None
        loss = kd.info_nce(xs, xt)
        loss.backward()
elif op in ("gemm_dh", "gemm_dw", "lm_head"):
    n, vl, H = 65536, 18992, 896
    g = (torch.randn(n, vl, device=dev) * 1e-3).to(torch.bfloat16)
    W = (torch.randn(vl, H, device=dev) * 0.02).to(torch.bfloat16)
    hid = torch.randn(n, H, device=dev).to(torch.bfloat16)
    for _ in range(REPS):
        if op == "gemm_dh":
            loss = kd.gemm(g, W, b_mn_major=True, out_dtype=torch.float32).sum()
        elif op == "gemm_dw":
            loss = kd.gemm(g, hid, a_mn_major=True, b_mn_major=True).float().sum()
        else:
            loss = sharded.lm_head_fwd(hid, W, T=2.0)[0].float().sum()
elif op == "vshard":
    n, vl, H = 65536, 18992, 896
    hid = torch.randn(n, H, device=dev).to(torch.bfloat16).requires_grad_(True)
    W = (torch.randn(vl, H, device=dev) * 0.02).to(torch.bfloat16).requires_grad_(True)
    ts = (torch.randn(n, vl, device=dev) * 2.0).to(torch.bfloat16)
    lab = torch.randint(0, vl, (n,), device=dev)
    for _ in range(2):
        hid.grad = None
        W.grad = None
        loss = sharded.lm_head_kd_vocab_sharded(hid, W, ts, lab, 0, V, 2.0)
        loss.backward()
else:
    raise SystemExit(f"unknown op {op}")
torch.cuda.synchronize()
print(op, "loss", float(loss))
