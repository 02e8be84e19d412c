"""tcgen05.mma issue-rate probe (see tools/probe/probe.cu, built into tools/probe/libaclgan_probe.so): cycles per 128 x N x 16 bf16 MMA for several configurations."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import aclgan_native as N  # noqa: E402

L = C.CDLL(N.build_probe())
L.aclgan_umma_probe.argtypes = [C.c_int] * 9 + [C.c_uint64, C.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 2048


def run(n, commit_every=4, swizzle=2, k_advance=1, rotate=4, a_mn=0, b_mn=0, ctas=148):
    rc = L.aclgan_umma_probe(n, iters, commit_every, swizzle, k_advance, rotate, a_mn, b_mn, ctas, out.data_ptr(), sp)
    assert rc == 0, rc
    torch.cuda.synchronize()
    cyc = out[:ctas].double()
    return float(cyc.mean()) / iters, float(cyc.max()) / iters


print("cycles per MMA (mean over CTAs / max), M=128, K=16, ideal = N/2")
for n in (256, 128, 64, 32, 16):
    print("N=%3d SW128 K-major commit/4      : %6.1f %6.1f" % ((n,) + run(n, rotate=1)))
for n in (256, 64):
    print("N=%3d no intermediate commits     : %6.1f %6.1f" % ((n,) + run(n, commit_every=2048, rotate=1)))
    print("N=%3d two accumulators alternating : %6.1f %6.1f" % ((n,) + run(n, rotate=2)))
    print("N=%3d no K advance                : %6.1f %6.1f" % ((n,) + run(n, k_advance=0, rotate=1)))
for sw, name in ((2, "128B"), (4, "64B"), (6, "32B"), (0, "none")):
    print("N=256 swizzle %-5s               : %6.1f %6.1f" % ((name,) + run(256, swizzle=sw, rotate=1)))
print("N=256 A MN-major                   : %6.1f %6.1f" % run(256, a_mn=1, rotate=1))
print("N=256 A,B MN-major                 : %6.1f %6.1f" % run(256, a_mn=1, b_mn=1, rotate=1))
print("N=256 single CTA on the chip       : %6.1f %6.1f" % run(256, ctas=1, rotate=1))
print("N= 64 single CTA on the chip       : %6.1f %6.1f" % run(64, ctas=1, rotate=1))

L.aclgan_tmem_ld_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
out4 = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
for mode, name, per in ((0, "x32 + wait", 1), (1, "2 x x32 + wait", 2), (2, "x16 + wait", 1)):
    rc = L.aclgan_tmem_ld_probe(1024, mode, 148, out4.data_ptr(), sp)
    assert rc == 0
    torch.cuda.synchronize()
    print("tcgen05.ld %-16s: %7.1f cycles per load (per warp, 4 warps / SM)" % (name, float(out4.double().mean()) / 1024 / per))
