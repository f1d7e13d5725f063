"""torchrun --nproc-per-node N tools/trial_shard_check.py : ONE RRR session with its trials sharded over N GPUs (replicated
parameters, one all-reduce of the gradient per closure evaluation) against the same fit on one GPU.  Also times both at
BASELINE size when FULL=1."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch, torch.distributed as dist
import bench
from model.rrr import train_model_from_frames
from parallel import pack_trial_shard, train_trial_sharded

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda")
dist.init_process_group("nccl")
full = os.environ.get("FULL") == "1"
K, Kt, F, N = (400, 80, 110 * 166, 144) if full else (48, 16, 500, 20)
mode = os.environ.get("MODE", "classic")                     # "exact": the default exact-operand layout (2 half planes, float64 epilogues)
planes = None if mode != "classic" else (1 if full else 3)
ftr, ctr, fte, cte = bench.rrr_inputs(K, Kt, F, N, 0, pinned=True)
sidx = bench.sorted_idx_42()
cut = lambda n: (n * rank // world, n * (rank + 1) // world)
(a, b), (c, d) = cut(K), cut(Kt)
ok = True
for it in range(3 if full else 1):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    entry = pack_trial_shard(ftr[a:b], ctr[a:b], fte[c:d], cte[c:d], sidx, 3, planes=planes or 1, mode=mode)
    model, res = train_trial_sharded(entry, 100.0, 3, planes=planes)
    val = float(res["mse_val_mean"]); dt = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        print(f"sharded over {world}: val SSE {val:.6f}  e2e {dt:.1f} ms  evals {model.n_closure_evals}")
if rank == 0:
    for it in range(3 if full else 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ref, r1, _ = train_model_from_frames(ftr, ctr, fte, cte, sidx, l2=100.0, n_comp=3, planes=planes, mode=None if mode == "classic" else mode)
        v1 = float(r1["mse_val_mean"]); dt = (time.perf_counter() - t0) * 1e3
        print(f"one GPU          : val SSE {v1:.6f}  e2e {dt:.1f} ms")
    rel = abs(val - v1) / v1
    dU = float((model.model["session_U"] - ref.model["session_U"]).abs().max())
    print(f"rel diff val SSE {rel:.3e}  |dU|max {dU:.3e}")
    # sharding regroups fp32 partial sums (1e-7); the un-line-searched fit amplifies that ~150x (classic, 1 plane); the exact
    # mode stays within the 1e-3 parity tolerance
    ok = rel < (5e-3 if (full and mode == "classic") else 1e-3)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
