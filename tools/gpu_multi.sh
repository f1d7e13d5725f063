#!/bin/bash
# multi-GPU round (gpurun --gpus N): joint-model check over NCCL, bench at N ranks, plus the single-GPU gpu test suite
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/joint_2gpu_check.py > gpurun_out/joint_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/joint_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/trial_shard_check.py > gpurun_out/trial_shard_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/trial_shard_${N}gpu.log
FULL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 tools/trial_shard_check.py > gpurun_out/trial_shard_full_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/trial_shard_full_${N}gpu.log
python bench.py --gpus 1 --steps 3 --no-cpu-baseline > gpurun_out/scale_rrr_1.json 2> gpurun_out/scale_rrr_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 > gpurun_out/scale_rrr_$N.json 2> gpurun_out/scale_rrr_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --impl reference > gpurun_out/scale_ref_$N.json 2> gpurun_out/scale_ref_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload linear --steps 30 > gpurun_out/scale_linear_$N.json 2> gpurun_out/scale_linear_$N.err
echo done
