"""torchrun --nproc-per-node 2 tools/joint_2gpu_check.py : the session-sharded joint RRR model (shared V all-reduced
over NCCL) against the same joint model fitted on ONE GPU.  Prints max deviations; exits non-zero on mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "video-spike_b200"), ROOT]
import numpy as np, torch, torch.distributed as dist
from tests.helpers import small_rrr_problem
from model.rrr import RRRGD, train_model
from parallel import shard_sessions, train_joint_model

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda")
dist.init_process_group("nccl")
td = {}
for i, (K, F, N) in enumerate([(16, 70, 6), (12, 90, 11), (20, 50, 8), (14, 64, 9)]):
    td.update(small_rrr_problem(seed=10 + i, K=K, Kt=5, F=F, N=N, eid=f"s{i}"))
plan = [(e, td[e]["y"][0].shape[2], td[e]["X"][0].shape[2], td[e]["y"][0].shape[1]) for e in td]
mine = shard_sessions(list(td), rank, world)
local = {e: td[e] for e in mine}
m = RRRGD(local, 3, l2=100.0, planes=3, init_plan=plan); m.to(dev)
_, res = train_joint_model(m, local)
ok = True
if rank == 0:
    ref = RRRGD(td, 3, l2=100.0, planes=3); ref.to(dev)
    _, rr = train_model(ref, td, ref.make_optimizer(), "tmp", save=False)
    dv = float((m.model["V"] - ref.model["V"]).abs().max())
    du = max(float((m.model[f"{e}_U"] - ref.model[f"{e}_U"]).abs().max()) for e in mine)
    rel = abs(float(res["mse_val_mean"]) - float(rr["mse_val_mean"])) / float(rr["mse_val_mean"])
    print(f"world={world} sessions/rank={len(mine)} evals={m.n_closure_evals} |dV|max={dv:.3e} |dU|max={du:.3e} val SSE rel diff={rel:.3e}")
    ok = dv < 1e-7 and du < 1e-7 and rel < 1e-8
vs_all = [None] * world
dist.all_gather_object(vs_all, m.model["V"].detach().cpu())
if rank == 0:
    same = all(torch.equal(vs_all[0], v) for v in vs_all)
    print("V replicas bit-identical across ranks:", same)
    ok = ok and same
dist.destroy_process_group()
sys.exit(0 if ok else 1)
