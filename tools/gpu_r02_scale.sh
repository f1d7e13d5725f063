#!/bin/bash
# round 2: the lines the driver's scaling run produces at N GPUs (default flags = joint shared-V model + row-parallel Linear),
# plus the trial-sharded single session.   usage: bash tools/gpu_r02_scale.sh N
N=${1:-4}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 $2 \
      > gpurun_out/r02s_$1_$N.json 2> gpurun_out/r02s_$1_$N.err; echo "$1 rc=$?"
  tail -c 300 gpurun_out/r02s_$1_$N.err
}
run default ""
run strong "--workload rrr --strong --dropin-e2e 0"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 --ref-budget 40 \
      > gpurun_out/r02s_ref_$N.json 2> gpurun_out/r02s_ref_$N.err; echo "reference arm rc=$?"
python - <<PY
import json
for nm in ("default", "strong", "ref"):
    try:
        d = json.loads([l for l in open("gpurun_out/r02s_%s_$N.json" % nm) if l.startswith("{")][-1])
        print(nm, "n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value", round(d["value"]), "e2e ms", round(d["e2e"].get("ms_per_step", 0), 1), str(d["config"].get("parallelism"))[:60], "scaling", d["scaling"],
              "parity", (d.get("parity") or {}).get("fit_rel_diff"), "| linear", (d.get("linear") or {}).get("ms_per_step"))
    except Exception as e:
        print(nm, "unreadable", e)
PY
