// Hardware probe (perf triage tool, not on the product path): issue rate of tcgen05.mma kind::f16 with both operands in
// shared memory as a function of N, swizzle mode, K-advance and commit frequency.  One CTA per SM; shared memory holds
// zeros; reports clock cycles per MMA measured by the issuing thread (start -> completion of the last commit).
#include "common.cuh"

namespace aclgan {

struct ProbeArgs {
    int n;            // UMMA N
    int iters;        // MMAs issued
    int commit_every; // commit + barrier wait granularity: commit after this many MMAs (the wait is only at the end)
    int swizzle;      // descriptor layout type: 2 = 128B, 4 = 64B, 6 = 32B, 0 = none
    int k_advance;    // 1: cycle through the 4 K slices of the 64-wide block; 0: always slice 0
    int rotate;       // number of distinct A/B tile pairs cycled through (1..4)
    int a_mn, b_mn;   // operand major-ness (1 = MN-major)
    int pair;         // unused here
    long long* out;   // [gridDim.x] cycles
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar_mid;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 4 * 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_mid, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)a.n, (uint32_t)a.a_mn, (uint32_t)a.b_mn);
        const uint32_t sa = smem_u32(smem);
        const uint32_t sb = sa + 16384;
        uint64_t da = make_smem_desc_sw128(sa, a.a_mn ? 8192 : 16, 1024);
        uint64_t db = make_smem_desc_sw128(sb, a.b_mn ? 8192 : 16, 1024);
        const uint64_t mask = ~(static_cast<uint64_t>(7) << 61);
        da = (da & mask) | (static_cast<uint64_t>(a.swizzle) << 61);
        db = (db & mask) | (static_cast<uint64_t>(a.swizzle) << 61);
        const uint64_t adv_a = a.k_advance ? (a.a_mn ? 128 : 2) : 0, adv_b = a.k_advance ? (a.b_mn ? 128 : 2) : 0;
        const bool mid = a.commit_every <= 4;
        const uint32_t acc_cols = (a.rotate > 1) ? 256u : 0u;       // rotate > 1: alternate between two accumulators
        const long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < a.iters; i += 4) {
            const uint32_t d = tmem_base + ((i >> 2) & 1) * acc_cols;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d, da + adv_a * kk, db + adv_b * kk, idesc, i >= 8 ? 1u : (kk ? 1u : 0u));
            if (mid) umma_commit(&bar_mid);      // (never waited on)
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        a.out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// TMEM read probe: 4 warps, each reading its own 32 lanes: cycles per tcgen05.ld.32x32b.x32 (+ wait)
__global__ void __launch_bounds__(128, 1) tmem_ld_probe_kernel(int iters, int mode, long long* out) {
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) {            // one x32 load, wait, consume
            uint32_t r[32];
            tmem_ld_32x32(t_row + (i & 7) * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += r[k];
        } else if (mode == 1) {     // two x32 loads in flight
            uint32_t r[32], q[32];
            tmem_ld_32x32(t_row + (i & 3) * 64, r);
            tmem_ld_32x32(t_row + (i & 3) * 64 + 32, q);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += r[k] + q[k];
        } else {                    // x16
            uint32_t r[16];
            tmem_ld_32x16(t_row + (i & 15) * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += r[k];
        }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = (t1 - t0) + (acc == 0x12345678u ? 1 : 0);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace aclgan

extern "C" int aclgan_tmem_ld_probe(int iters, int mode, int ctas, uint64_t out, void* stream) {
    aclgan::tmem_ld_probe_kernel<<<ctas, 128, 0, (cudaStream_t)stream>>>(iters, mode, reinterpret_cast<long long*>(out));
    return (int)cudaGetLastError();
}

extern "C" int aclgan_umma_probe(int n, int iters, int commit_every, int swizzle, int k_advance, int rotate, int a_mn, int b_mn,
                                 int ctas, uint64_t out, void* stream) {
    using namespace aclgan;
    static bool attr = false;
    const int smem = 4 * 49152 + 2048;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    ProbeArgs a;
    a.n = n; a.iters = iters; a.commit_every = commit_every < 1 ? 1 : commit_every; a.swizzle = swizzle;
    a.k_advance = k_advance; a.rotate = rotate < 1 ? 1 : (rotate > 4 ? 4 : rotate); a.a_mn = a_mn; a.b_mn = b_mn; a.pair = 0;
    a.out = reinterpret_cast<long long*>(out);
    umma_probe_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}
