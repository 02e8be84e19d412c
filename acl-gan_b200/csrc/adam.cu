// Multi-tensor Adam + weight re-packing in one pass over the parameters.
// torch.optim.Adam semantics (reference trainer.py:39-42): g' = g*scale + wd*p ; m = b1 m + (1-b1) g' ;
// v = b2 v + (1-b2) g'^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  The step counter lives on the
// device (advanced by a 1-thread kernel) so the whole update is CUDA-graph capturable; `scale` folds the
// 1/world_size of the data-parallel gradient all-reduce.  While the fp32 master value is on chip the
// bf16 hi(/lo) planes of both packed layouts (forward K-major, transposed) are rewritten, so the next step's
// tensor-core kernels need no separate repack pass.
//
// HBM-bound (16 B read + 12 B written per parameter, + 2-8 B of packed bf16): every stream must be coalesced.  The
// master / moments are OIHW (tap fastest), the gradient arena and the forward packing are [co][tap][ci] (ci fastest),
// the transposed packing is [ci][tap][co] (co fastest).  A conv tensor is therefore processed in tiles of
// (tco output channels) x (tci input channels) x (all taps) staged in shared memory: the gradient tile is read in ITS
// fastest order, the update runs in OIHW order, and each packing is written in its own fastest order - the three
// permutations happen in shared memory (odd strides: conflict-free).  Dense tensors (biases, Linear weights, 1x1 heads)
// take the flat path.
#include "common.cuh"

namespace aclgan {

constexpr int kAdamThreads = 256;
constexpr int kAdamFlat = 1024;            // elements per CTA of a flat tensor
constexpr int kAdamTileFloats = 13000;     // shared-memory tile budget (52 KB: four CTAs of 256 threads per SM)

struct AdamTile {
    int flat;          // 1: contiguous gradient, no packed planes
    int taps, tp;      // filter taps, padded to odd (shared-memory stride of ci)
    int tco, tci;      // tile extent
    int n_ct;          // tiles along ci
    int units;         // CTAs for this tensor
    int co_stride;     // shared-memory stride of co (odd)
};

__host__ __device__ inline AdamTile adam_tile(const aclgan_adam_tensor& T) {
    AdamTile t;
    const int64_t numel = (int64_t)T.d[0] * T.d[1] * T.d[2] * T.d[3];
    const bool contiguous = T.gs[3] == 1 && T.gs[2] == T.d[3] && T.gs[1] == (int64_t)T.d[2] * T.d[3] &&
                            T.gs[0] == (int64_t)T.d[1] * T.d[2] * T.d[3];
    t.flat = (T.pk[0][0] == 0 && T.pk[1][0] == 0 && contiguous) ? 1 : 0;
    t.taps = T.d[2] * T.d[3];
    t.tp = t.taps | 1;
    t.tci = T.d[1] < 32 ? T.d[1] : 32;
    t.tco = T.d[0] < 32 ? T.d[0] : 32;
    while (t.tco > 8 && t.tco * (t.tci * t.tp + 1) > kAdamTileFloats) t.tco /= 2;
    if (t.tco * (t.tci * t.tp + 1) > kAdamTileFloats || t.taps > 64) t.flat = 2;      // (no shipped layer: > 64 taps) scalar fallback
    t.co_stride = t.tci * t.tp + 1;
    t.n_ct = (T.d[1] + t.tci - 1) / t.tci;
    if (t.flat) t.units = (int)((numel + kAdamFlat - 1) / kAdamFlat);
    else t.units = ((T.d[0] + t.tco - 1) / t.tco) * t.n_ct;
    return t;
}

__global__ void adam_advance_kernel(float* hyper) {
    if (threadIdx.x == 0 && blockIdx.x == 0) hyper[6] += 1.0f;
}

struct AdamCoef {
    float lr_bc1, bc2s, b1, b2, eps, wd, gscale;
};

__device__ __forceinline__ float adam_update(const AdamCoef& c, float pv, float graw, float& mv, float& vv) {
    const float gv = graw * c.gscale + c.wd * pv;
    mv = c.b1 * mv + (1.f - c.b1) * gv;
    vv = c.b2 * vv + (1.f - c.b2) * gv * gv;
    return pv - c.lr_bc1 * mv / (sqrtf(vv) / c.bc2s + c.eps);
}

__global__ void __launch_bounds__(kAdamThreads, 4) adam_kernel(const aclgan_adam_tensor* __restrict__ table,
                                                            const int* __restrict__ chunks, const float* __restrict__ hyper) {
    extern __shared__ float tile[];
    __shared__ AdamCoef coef_s;
    const int tid = chunks[2 * blockIdx.x];
    const int unit = chunks[2 * blockIdx.x + 1];
    const aclgan_adam_tensor T = table[tid];
    const AdamTile G = adam_tile(T);
    if (threadIdx.x == 0) {
        const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2];
        const double step = hyper[6];
        const float bc1 = (float)(1.0 - pow((double)b1, step));
        AdamCoef c;
        c.lr_bc1 = lr / bc1;
        c.bc2s = (float)sqrt(1.0 - pow((double)b2, step));
        c.b1 = b1; c.b2 = b2; c.eps = hyper[3]; c.wd = hyper[4]; c.gscale = hyper[5];
        coef_s = c;
    }
    __syncthreads();
    const AdamCoef c = coef_s;
    float* __restrict__ p = reinterpret_cast<float*>(T.p);
    float* __restrict__ m = reinterpret_cast<float*>(T.m);
    float* __restrict__ v = reinterpret_cast<float*>(T.v);
    const float* __restrict__ g = reinterpret_cast<const float*>(T.g) + T.goff;
    const int64_t numel = (int64_t)T.d[0] * T.d[1] * T.d[2] * T.d[3];

    if (G.flat == 1) {
        const int64_t first = (int64_t)unit * kAdamFlat;
        float pv[kAdamFlat / kAdamThreads], gv[kAdamFlat / kAdamThreads], mv[kAdamFlat / kAdamThreads], vv[kAdamFlat / kAdamThreads];
#pragma unroll
        for (int i = 0; i < kAdamFlat / kAdamThreads; ++i) {
            const int64_t e = first + i * kAdamThreads + threadIdx.x;
            if (e < numel) { pv[i] = p[e]; gv[i] = g[e]; mv[i] = m[e]; vv[i] = v[e]; }
        }
#pragma unroll
        for (int i = 0; i < kAdamFlat / kAdamThreads; ++i) {
            const int64_t e = first + i * kAdamThreads + threadIdx.x;
            if (e < numel) {
                const float np = adam_update(c, pv[i], gv[i], mv[i], vv[i]);
                m[e] = mv[i]; v[e] = vv[i]; p[e] = np;
            }
        }
        return;
    }
    if (G.flat == 2) {
        // scalar fallback (tiles that do not fit shared memory): element order = OIHW
        for (int64_t e = (int64_t)unit * kAdamFlat + threadIdx.x; e < numel && e < (int64_t)(unit + 1) * kAdamFlat; e += kAdamThreads) {
            int64_t r = e;
            const int kw = (int)(r % T.d[3]); r /= T.d[3];
            const int kh = (int)(r % T.d[2]); r /= T.d[2];
            const int ci = (int)(r % T.d[1]);
            const int co = (int)(r / T.d[1]);
            float mv = m[e], vv = v[e];
            const float np = adam_update(c, p[e], g[co * T.gs[0] + ci * T.gs[1] + kh * T.gs[2] + kw * T.gs[3]], mv, vv);
            m[e] = mv; v[e] = vv; p[e] = np;
            for (int k = 0; k < 2; ++k) {
                if (T.pk[k][0] == 0) continue;
                const int64_t o = T.aff[k][0] + co * T.aff[k][1] + ci * T.aff[k][2] + kh * T.aff[k][3] + kw * T.aff[k][4];
                const __nv_bfloat16 hi = __float2bfloat16_rn(np);
                reinterpret_cast<__nv_bfloat16*>(T.pk[k][0])[o] = hi;
                if (T.planes == 2) reinterpret_cast<__nv_bfloat16*>(T.pk[k][1])[o] = __float2bfloat16_rn(np - __bfloat162float(hi));
            }
        }
        return;
    }

    const int co0 = (unit / G.n_ct) * G.tco, ci0 = (unit % G.n_ct) * G.tci;
    const int nco = min(G.tco, T.d[0] - co0), nci = min(G.tci, T.d[1] - ci0);      // valid extent of this tile
    const int taps = G.taps;

    // The kernel is bound by instruction issue unless the per-element index arithmetic is kept to a handful of 32-bit
    // operations (the first tiled version spent 360 thread instructions per parameter, ncu: issue slots 73 % busy at 2.3 TB/s):
    //  * per-tap offsets (kh * s_kh + kw * s_kw) of the gradient and of both packings come from small shared-memory tables;
    //  * a warp walks (slow index, tap) rows with incremental (no division) updates, its lanes are the fast index;
    //  * all offsets are 32-bit (every buffer of a parameter group is far below 2^31 elements);
    //  * loops are unrolled with the loads of a batch issued before the first use (memory-level parallelism).
    __shared__ int tab[3][64];
    if (threadIdx.x < taps) {
        const int kh = threadIdx.x / T.d[3], kw = threadIdx.x - kh * T.d[3];
        tab[0][threadIdx.x] = (int)(kh * T.gs[2] + kw * T.gs[3]);
        tab[1][threadIdx.x] = (int)(kh * T.aff[0][3] + kw * T.aff[0][4]);
        tab[2][threadIdx.x] = (int)(kh * T.aff[1][3] + kw * T.aff[1][4]);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kAdamThreads / 32;
    constexpr int U = 8;
    const int q16 = kWarps / taps, r16 = kWarps % taps;       // advancing a (slow, tap) row index by kWarps

    // phase 1: gradient tile -> shared memory, read in the gradient layout's fastest order
    {
        const bool co_fast = (T.gs[0] == 1 || T.gs[0] == -1);
        const int nf = co_fast ? nco : nci, ns = co_fast ? nci : nco;            // fast / slow extents
        const int gf = (int)(co_fast ? T.gs[0] : T.gs[1]), gsl = (int)(co_fast ? T.gs[1] : T.gs[0]);
        const int sf = co_fast ? G.co_stride : G.tp, ss = co_fast ? G.tp : G.co_stride;
        const float* __restrict__ gl = g + ((int64_t)co0 * T.gs[0] + (int64_t)ci0 * T.gs[1]) + lane * gf;
        float* tl = tile + lane * sf;
        const int rows = ns * taps;
        const bool lane_ok = lane < nf;
        int sl = warp / taps, tap = warp - (warp / taps) * taps;                   // row r = sl * taps + tap
        for (int r0 = warp; r0 < rows; r0 += U * kWarps) {
            float val[U];
            int so[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = (r0 + u * kWarps < rows) && lane_ok;
                so[u] = ok ? sl * ss + tap : -1;
                if (ok) val[u] = gl[sl * gsl + tab[0][tap]];
                sl += q16; tap += r16;
                if (tap >= taps) { tap -= taps; ++sl; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (so[u] >= 0) tl[so[u]] = val[u];
        }
    }
    __syncthreads();
    // phase 2: the update in OIHW order: one warp per output channel, lanes along the contiguous (ci, tap) run of p / m / v
    {
        const int run = nci * taps;
        const int q32 = 32 / taps, r32 = 32 % taps;
        constexpr int V = 4;
        for (int co_l = warp; co_l < nco; co_l += kWarps) {
            const int64_t e0 = ((int64_t)(co0 + co_l) * T.d[1] + ci0) * taps;
            float* __restrict__ pp = p + e0;
            float* __restrict__ mm = m + e0;
            float* __restrict__ vp = v + e0;
            float* trow = tile + co_l * G.co_stride;
            int ci_l = lane / taps, tap = lane - (lane / taps) * taps;
            for (int i0 = lane; i0 < run; i0 += V * 32) {
                float pv[V], mv[V], vv[V];
                int so[V];
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const int i = i0 + u * 32;
                    so[u] = i < run ? ci_l * G.tp + tap : -1;
                    if (i < run) { pv[u] = pp[i]; mv[u] = mm[i]; vv[u] = vp[i]; }
                    ci_l += q32; tap += r32;
                    if (tap >= taps) { tap -= taps; ++ci_l; }
                }
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const int i = i0 + u * 32;
                    if (so[u] >= 0) {
                        const float np = adam_update(c, pv[u], trow[so[u]], mv[u], vv[u]);
                        mm[i] = mv[u]; vp[i] = vv[u]; pp[i] = np;
                        trow[so[u]] = np;
                    }
                }
            }
        }
    }
    __syncthreads();
    // phase 3: packed bf16 planes, each written in its own fastest order
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        if (T.pk[k][0] == 0) continue;
        const bool co_fast = (T.aff[k][1] == 1 || T.aff[k][1] == -1);
        const int nf = co_fast ? nco : nci, ns = co_fast ? nci : nco;
        const int af = (int)(co_fast ? T.aff[k][1] : T.aff[k][2]), asl = (int)(co_fast ? T.aff[k][2] : T.aff[k][1]);
        const int sf = co_fast ? G.co_stride : G.tp, ss = co_fast ? G.tp : G.co_stride;
        const int64_t o0 = T.aff[k][0] + (int64_t)co0 * T.aff[k][1] + (int64_t)ci0 * T.aff[k][2] + (int64_t)lane * af;
        __nv_bfloat16* __restrict__ hi_p = reinterpret_cast<__nv_bfloat16*>(T.pk[k][0]) + o0;
        __nv_bfloat16* __restrict__ lo_p = reinterpret_cast<__nv_bfloat16*>(T.pk[k][1]) + o0;
        const float* tl = tile + lane * sf;
        const int* tb = tab[1 + k];
        const int rows = ns * taps;
        const bool two = T.planes == 2;
        int sl = warp / taps, tap = warp - (warp / taps) * taps;
        if (lane < nf) {
#pragma unroll 4
            for (int r = warp; r < rows; r += kWarps) {
                const float np = tl[sl * ss + tap];
                const int o = sl * asl + tb[tap];
                const __nv_bfloat16 hi = __float2bfloat16_rn(np);
                hi_p[o] = hi;
                if (two) lo_p[o] = __float2bfloat16_rn(np - __bfloat162float(hi));
                sl += q16; tap += r16;
                if (tap >= taps) { tap -= taps; ++sl; }
            }
        }
    }
}

}  // namespace aclgan

extern "C" int aclgan_adam_units(const aclgan_adam_tensor* t) { return aclgan::adam_tile(*t).units; }

extern "C" int aclgan_adam_advance(uint64_t hyper, void* stream) {
    aclgan::adam_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(hyper));
    return (int)cudaGetLastError();
}

extern "C" int aclgan_adam_step(uint64_t table, uint64_t chunks, int32_t n_chunks, uint64_t hyper, void* stream) {
    if (n_chunks <= 0) return ACLGAN_OK;
    static bool attr_set = false;
    const int smem = aclgan::kAdamTileFloats * (int)sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(aclgan::adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    aclgan::adam_kernel<<<n_chunks, aclgan::kAdamThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const aclgan_adam_tensor*>(table), reinterpret_cast<const int*>(chunks),
        reinterpret_cast<const float*>(hyper));
    return (int)cudaGetLastError();
}
