"""2+ GPU check of BASELINE configs[4]: vocabulary-parallel student lm_head (896 x 151 936; the tcgen05 kernel, or the
library GEMM with `library` as the first argument) feeding the vocabulary-sharded LoCa KD.  Correctness at 64 rows against the fp64 oracle, then the 64k-token step time.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/lmhead_vshard_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded, synthetic
from oracle import kd_oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
K, V, T = 896, 151936, 2.0
GEMM = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
lo, hi = rank * (V // world), (rank + 1) * (V // world) if rank < world - 1 else V
g = torch.Generator().manual_seed(11)
n = 64
H = (torch.randn(n, K, generator=g) * 1.0).to(torch.bfloat16)
W = (torch.randn(V, K, generator=g) * 0.02).to(torch.bfloat16)
t, _, lab = synthetic.logit_batch(n, V, seed=12)
logits = (H.float() @ W.float().t()).to(torch.bfloat16)                 # what the bf16 GEMM hands to the loss
ref = O.loca_kl(t, logits, lab, T, 0.8, chunk_rows=8)
dH_ref = ref.grad @ W.double()
dW_ref = ref.grad.t() @ H.double()
h = H.to(dev).requires_grad_(True)
w = W[lo:hi].to(dev).requires_grad_(True)
loss = sharded.lm_head_kd_vocab_sharded(h, w, t[:, lo:hi].contiguous().to(dev), lab.to(dev), lo, V, T, gemm=GEMM)
loss.backward()
e_loss = abs(loss.item() - ref.loss.item()) / abs(ref.loss.item())
e_dh = ((h.grad.double().cpu() - dH_ref).norm() / dH_ref.norm()).item()
e_dw = ((w.grad.double().cpu() - dW_ref[lo:hi]).norm() / dW_ref[lo:hi].norm()).item()
ok = e_loss <= 2e-4 and e_dh <= 1e-2 and e_dw <= 1e-2       # the bf16 GEMM output is re-rounded: looser than the loss-only gates
print(f"rank {rank}: loss rel {e_loss:.2e} dH rel {e_dh:.2e} dW rel {e_dw:.2e} {'OK' if ok else 'FAIL'}", flush=True)
del h, w, loss
n = 65536
H = torch.randn(n, K, device=dev, dtype=torch.bfloat16, requires_grad=True)
Ws = (torch.randn(hi - lo, K, device=dev) * 0.02).to(torch.bfloat16).requires_grad_(True)
ts = (torch.randn(n, hi - lo, device=dev) * 2.0).to(torch.bfloat16)
lab = torch.randint(0, V, (n,), device=dev)
dist.broadcast(lab, 0)
def step():
    H.grad = None; Ws.grad = None
    sharded.lm_head_kd_vocab_sharded(H, Ws, ts, lab, lo, V, T, gemm=GEMM).backward()
for _ in range(2): step()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): step()
b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / 3], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[{GEMM}] lm_head (896 x {V}) + LoCa KD, {n} tokens over {world} GPUs: {ms.item():.1f} ms/step = {n / ms.item() * 1e3:.0f} tokens/s", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
