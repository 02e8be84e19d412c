#!/bin/bash
# round 2, re-entry call (GPU budget ~15 min incl. box set-up): the reference's unmodified train.py / test.py drop-in test, smoke(),
# the default bench line of the rebuilt library, then as much of the -m gpu suite as the remaining time allows.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_dropin_drivers.py -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/dropin.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout ${1:-300} python -m pytest tests -m gpu -x -q --deselect tests/test_dropin_drivers.py 2>&1 | tail -6 | tee gpurun_out/suite.log
