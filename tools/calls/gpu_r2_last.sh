#!/bin/bash
# last call of round 2: every -m gpu test on the committed build, the P2 line (selfie2anime, parity precision fp32x3)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/last_tests.log
python bench.py --steps 10 --warmup 3 --config selfie2anime.yaml --precision fp32x3 --no-cpu-baseline --no-library-bar > gpurun_out/last_bench_p2.json 2> gpurun_out/last_bench_p2.err
tail -c 1500 gpurun_out/last_bench_p2.json; tail -2 gpurun_out/last_bench_p2.err
