"""Runs every collected -m gpu test of the given files in its OWN process (a device trap in one test
does not poison the CUDA context of the next).  Prints one line per test + the tail of each failure.
usage: python tools/gpu_isolated.py tests/test_gpu_igemm.py [...]  (results also in gpurun_out/isolated.log)"""
import os
import subprocess
import sys
import time

files = sys.argv[1:]
os.makedirs("gpurun_out", exist_ok=True)
r = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + files,
                   capture_output=True, text=True)
ids = [l.strip() for l in r.stdout.splitlines() if "::" in l]
log = open("gpurun_out/isolated.log", "w")
summary = []
for tid in ids:
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-m", "pytest", tid, "-q", "-m", "gpu", "-x", "-s"],
                           capture_output=True, text=True, timeout=240)
        ok = p.returncode == 0
        out = p.stdout + p.stderr
    except subprocess.TimeoutExpired as e:
        ok, out = False, "TIMEOUT\n" + str(e.stdout)[-2000:]
    line = "%s %s (%.1fs)" % ("PASS" if ok else "FAIL", tid, time.time() - t0)
    print(line, flush=True)
    log.write(line + "\n")
    if not ok:
        tail = "\n".join(out.splitlines()[-25:])
        print(tail, flush=True)
        log.write(out[-6000:] + "\n")
    else:
        for l in out.splitlines():
            if l.startswith("["):
                print("   ", l, flush=True)
                log.write("    " + l + "\n")
    summary.append(ok)
print("isolated: %d/%d passed" % (sum(summary), len(summary)))
log.close()
sys.exit(0 if all(summary) else 1)
