"""Random-geometry sweep of the host-side plan builders (csrc/plans.cu) through the CPU emulation of their TMA / MMA / epilogue
contract (tests/emul.py, the functions of tests/test_plans.py): forward, data-gradient and weight-gradient plans of every
convolution kind of the networks (3x3 s1, 4x4 s2, 5x5 s1, 7x7 s1 full-width and pixel-window, 4x4 s2 pixel-window) at random
channel counts, batch sizes, heights and widths (ragged widths up to 280, hi/lo planes) against torch fp64 convolutions.
Runs on the CPU (no GPU needed).  usage: python tools/fuzz_plans.py [cases=150] [seed=1]
Last run (round 2 re-entry): 150 geometries x 3 plan kinds, 0 failures."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "acl-gan_b200")]
import test_plans as TP  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
random.seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
kinds = [(3, 1, 1, 0), (4, 2, 1, 0), (5, 1, 2, 0), (7, 1, 3, 0), (7, 1, 3, 1), (4, 2, 1, 1)]       # k, stride, pad, window
fails, count, t0 = 0, 0, time.time()
while count < cases:
    k, s, pad, window = random.choice(kinds)
    if window:
        cin, cout = (random.choice([3, 6]) if k == 4 else 3), random.choice([16, 64])
    elif k == 7:
        cin, cout = 64, random.choice([3, 4])
    else:
        cin, cout = random.choice([64, 128, 256]), random.choice([16, 32, 64, 128, 256])
    n = random.choice([1, 1, 2, 3])
    lo = 4 if s == 2 else pad + 1
    h = random.randint(lo, 20)
    w = random.choice([random.randint(lo, 40), random.randint(100, 280)])
    if s == 2:
        h, w = h + h % 2, w + w % 2
    planes = random.choice([1, 1, 2])
    if cin * cout * h * w * n > 3e8:
        continue
    count += 1
    runs = [("fwd", TP.test_conv_fwd_plan, (cin, cout, k, s, pad, window, n, h, w, planes)),
            ("wgrad", TP.test_conv_wgrad_plan, (cin, cout, k, s, pad, window, n, h, w, planes))]
    if not window or k == 7:
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        runs.append(("dgrad", TP.test_conv_dgrad_plan, (cin, cout, k, s, pad, n, ho, wo, planes)))
    for name, fn, args in runs:
        try:
            fn(*args)
        except Exception as e:          # noqa: BLE001 - report and keep sweeping
            fails += 1
            print(name, "FAIL", args, type(e).__name__, str(e)[:160])
print("cases", count, "fails", fails, "seconds %.0f" % (time.time() - t0))
sys.exit(1 if fails else 0)
