#!/bin/bash
# 2-GPU check of the data-parallel path (NCCL all-reduce of both gradient arenas, 1/world folded into the tiled Adam kernel) next to
# the 1-GPU line of the same box
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
python -c "
import json; d = json.loads(open('gpurun_out/bench_n1_same_box.json').read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/bench_p3_n2.json 2> gpurun_out/bench_p3_n2.err
python -c "
import json; d = json.loads(open('gpurun_out/bench_p3_n2.json').read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d.get('e2e', {}).get('value'))"
tail -2 gpurun_out/bench_p3_n2.err
