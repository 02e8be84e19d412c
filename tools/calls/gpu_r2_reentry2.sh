#!/bin/bash
# round 2, re-entry call 4: the whole -m gpu suite (no -x) incl. the new nsgan / full-size / drop-in tests, then a short bench line
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -s --tb=short -rf 2>&1 | grep -E "^\[batch|^\[step parity nsgan|passed|failed|^FAILED|^E  |Error" | cut -c1-700 | tail -60 | tee gpurun_out/suite2.log
timeout 120 python bench.py --no-cpu-baseline --no-library-bar > gpurun_out/bench_reentry2.json 2> gpurun_out/bench_reentry2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_reentry2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"])
PY
