#!/bin/bash
# multi-GPU regression of parallel.py after kernel/layout changes: joint shared-V model and trial-sharded session vs one GPU
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 tools/joint_2gpu_check.py > gpurun_out/joint_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/joint_${N}gpu.log
timeout 300 $T --master-port 29542 tools/trial_shard_check.py > gpurun_out/trial_shard_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/trial_shard_${N}gpu.log
echo done
