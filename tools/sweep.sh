#!/bin/bash
# SURVEY 8d configs 2/4/5 (single-GPU cells): RRR over K x N, Linear over frame size x batch.  One JSON line per cell.
out=gpurun_out/sweep.jsonl; : > $out
for K in 200 400; do for N in 144 436; do
  python bench.py --workload rrr --trials $K --neurons $N --steps 3 --no-cpu-baseline --no-parity >> $out 2>> gpurun_out/sweep.err
done; done
python bench.py --workload rrr --trials 400 --neurons 144 --planes 2 --steps 3 --no-cpu-baseline --no-parity >> $out 2>> gpurun_out/sweep.err
python bench.py --workload rrr --trials 400 --neurons 144 --planes 3 --steps 3 --no-cpu-baseline --no-parity >> $out 2>> gpurun_out/sweep.err
for D in $((120*128*128)) $((120*256*256)) $((240*256*256)); do for B in 16 32; do
  python bench.py --workload linear --input-dim $D --batch $B --steps 20 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
done; done
python bench.py --workload linear --neurons 436 --steps 20 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
echo done
