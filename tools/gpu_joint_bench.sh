#!/bin/bash
# joint shared-V model bench (BASELINE configs[2]) at N ranks, next to the independent-session line at the same N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --joint --no-cpu-baseline > gpurun_out/joint_rrr_$N.json 2> gpurun_out/joint_rrr_$N.err; echo "rc=$?" >> gpurun_out/joint_rrr_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 3 --no-cpu-baseline --no-parity > gpurun_out/indep_rrr_$N.json 2> gpurun_out/indep_rrr_$N.err; echo "rc=$?" >> gpurun_out/indep_rrr_$N.err
echo done
