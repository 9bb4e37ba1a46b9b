"""2+ GPU check of the vocabulary-sharded path over NCCL: every rank holds V/G columns, results against the fp64 oracle.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/vshard_nccl_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded, synthetic
from oracle import kd_oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
n, V, T = 64, 151936, 2.0
t, s, lab = synthetic.logit_batch(n, V, seed=5)
lo, hi = rank * (V // world), (rank + 1) * (V // world) if rank < world - 1 else V
ts, ss = t[:, lo:hi].contiguous().cuda(), s[:, lo:hi].contiguous().cuda()
ok = True
for mode in ("reference", "per_row", "kl"):
    loss, grad = sharded.logit_kd_vocab_sharded(ts, ss, lab.cuda(), lo, V, T, mode=mode)
    ref = O.loca_kl(t, s, lab, T, 0.8, mode=mode, chunk_rows=8) if mode != "kl" else O.temperature_kl(t, s, T, chunk_rows=8)
    e_loss = abs(loss.item() - ref.loss.item()) / abs(ref.loss.item())
    e_grad = ((grad.float().cpu().double() - ref.grad[:, lo:hi]).norm() / ref.grad[:, lo:hi].norm()).item()
    good = e_loss <= 1e-5 and e_grad <= 2.5e-3
    ok &= good
    print(f"rank {rank} mode {mode}: loss rel {e_loss:.2e} grad rel {e_grad:.2e} {'OK' if good else 'FAIL'}", flush=True)
# timing at a larger row count (64k tokens / G shards of V), as-written mode
n = 4096
t, s, lab = synthetic.logit_batch(n, V // world, seed=6, device="cuda")
lab = torch.randint(0, V, (n,), device="cuda")
for _ in range(2):
    sharded.logit_kd_vocab_sharded(t, s, lab, lo, V, T, mode="reference")
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    sharded.logit_kd_vocab_sharded(t, s, lab, lo, V, T, mode="reference")
torch.cuda.synchronize()
if rank == 0:
    print(f"vocab-sharded reference mode, {n} rows x {V} cols over {world} GPUs: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
