#!/bin/bash
# round 2, call A (GPU box): smoke, all -m gpu tests (incl. the new 256x256 / optimizer tests), default bench line, fold-mode bench
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_igemm.py > gpurun_out/tests_step.log 2>&1
grep -E "^\[|passed|failed|FAILED|Error|error|assert" gpurun_out/tests_step.log | cut -c1-1500 | tail -60
timeout 900 python -m pytest tests/test_gpu_igemm.py -m gpu -q > gpurun_out/tests_igemm.log 2>&1; tail -5 gpurun_out/tests_igemm.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ACLGAN_FOLD=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err
python - <<'PY'
import json
for f in ("bench", "bench_fold"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("schedule_2to1"), d.get("library_bar"))
    except Exception as e:
        print(f, "unparsed", e)
PY
ACLGAN_FOLD=1 timeout 600 python -m pytest tests/test_gpu_step256.py -m gpu -q -s -k "fixture" > gpurun_out/tests_fold256.log 2>&1; tail -4 gpurun_out/tests_fold256.log
