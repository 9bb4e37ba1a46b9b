"""8-GPU timing of the configs[4] step (vocabulary-parallel lm_head + KD, 65 536 tokens) with the d hidden all-reduce (a) blocking,
(b) started asynchronously under the d weight product, (c) the same with a few SMs kept free of the persistent GEMM.
   torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/vshard_overlap_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
K, V, T, n = 896, 151936, 2.0, 65536
lo, hi = rank * (V // world), (rank + 1) * (V // world) if rank < world - 1 else V
H = torch.randn(n, K, device=dev, dtype=torch.bfloat16, requires_grad=True)
Ws = (torch.randn(hi - lo, K, device=dev) * 0.02).to(torch.bfloat16).requires_grad_(True)
ts = (torch.randn(n, hi - lo, device=dev) * 2.0).to(torch.bfloat16)
lab = torch.randint(0, V, (n,), device=dev)
dist.broadcast(lab, 0)


def step():
    H.grad = None
    Ws.grad = None
    sharded.lm_head_kd_vocab_sharded(H, Ws, ts, lab, lo, V, T).backward()


def timed(label):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / 5], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:48s} {ms.item():7.3f} ms/step", flush=True)


real_async = sharded.DistCollectives.sum_async
sharded.DistCollectives.sum_async = lambda self, x: (self.sum_(x), None)[1]      # blocking all-reduce
timed("blocking all-reduce of d hidden")
sharded.DistCollectives.sum_async = real_async
for reserve in (0, 16, 28, 32):
    sharded._MAX_RESERVED_SMS = reserve
    timed(f"async all-reduce, up to {reserve} SMs kept free of the dW product (wave-neutral)")
dist.destroy_process_group()
