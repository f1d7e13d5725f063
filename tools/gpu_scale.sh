#!/bin/bash
# scaling lines at N ranks: independent-session RRR (weak), joint shared-V RRR, row-parallel Linear (strong)
N=${1:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $T --master-port 29531 bench.py --gpus $N --steps 3 --no-cpu-baseline --no-parity > gpurun_out/scale_rrr_$N.json 2> gpurun_out/scale_rrr_$N.err; echo "rc=$?" >> gpurun_out/scale_rrr_$N.err
timeout 400 $T --master-port 29532 bench.py --gpus $N --workload linear --steps 30 --no-cpu-baseline > gpurun_out/scale_linear_$N.json 2> gpurun_out/scale_linear_$N.err; echo "rc=$?" >> gpurun_out/scale_linear_$N.err
timeout 400 $T --master-port 29533 bench.py --gpus $N --steps 3 --joint --no-cpu-baseline > gpurun_out/joint_rrr_$N.json 2> gpurun_out/joint_rrr_$N.err; echo "rc=$?" >> gpurun_out/joint_rrr_$N.err
echo done
