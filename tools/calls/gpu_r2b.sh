#!/bin/bash
# round 2, call B (GPU box): smoke, all -m gpu tests (no -x), default bench line
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_igemm.py > gpurun_out/tests_step.log 2>&1
grep -E "^\[|passed|failed|FAILED|Error|error|^E  " gpurun_out/tests_step.log | cut -c1-1200 | tail -${TAILN:-70}
timeout 900 python -m pytest tests/test_gpu_igemm.py -m gpu -q > gpurun_out/tests_igemm.log 2>&1; tail -3 gpurun_out/tests_igemm.log
python bench.py --steps 10 --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}), d.get("schedule_2to1", {}).get("value"), d.get("library_bar"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "unparsed", e)
PY
