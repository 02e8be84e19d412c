#!/bin/bash
# 8-GPU lines: P3 (male2female bs 8/GPU) and P4 (glasses-removal bs 16/GPU, focus branch on); torchrun, NCCL over NVLink
mkdir -p gpurun_out
N=${NGPU:-8}
for spec in "male2female.yaml 8 p3" "glasses-removal.yaml 16 p4"; do
  set -- $spec
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
      --steps 10 --warmup 3 --config $1 --batch $2 --no-cpu-baseline --no-library-bar > gpurun_out/bench_$3_n$N.json 2> gpurun_out/bench_$3_n$N.err
  tail -c 1800 gpurun_out/bench_$3_n$N.json; echo; tail -2 gpurun_out/bench_$3_n$N.err
done
