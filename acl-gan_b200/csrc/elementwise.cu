// HBM-bound kernels around the tensor-core convolutions: image (un)packing, InstanceNorm / AdaIN / LayerNorm
// statistics + apply (fused with activation, residual add, nearest 2x upsample and the consumer's reflect pad),
// and their backward passes (fold of the reflect/upsample gather, activation mask, two-moment norm backward).
// All work on 8-channel (16 B bf16 / 32 B fp32) vectors of NHWC planes; reductions go warp/CTA-local first and
// then to fp64 atomics so the statistics do not suffer from fp32 cancellation.
#include <cstdlib>

#include "common.cuh"
#include "elementwise_rows.cuh"

namespace aclgan {

struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 f8_zero() {
    F8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
    return r;
}

__device__ __forceinline__ F8 load8(uint64_t base, int kind, int64_t idx) {
    F8 r;
    if (kind == 0) {
        const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r.v[2 * i] = __uint_as_float(w[i] << 16);
            r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
        const float4 a = p[0], b = p[1];
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    }
    return r;
}

__device__ __forceinline__ uint32_t pk2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// hi = bf16(v), lo = bf16(v - hi) into plane 0 / plane 1
__device__ __forceinline__ void store8_planes(const uint64_t (&pl)[2], int planes, int64_t idx, const F8& v) {
    uint4 q;
    q.x = pk2(v.v[0], v.v[1]); q.y = pk2(v.v[2], v.v[3]); q.z = pk2(v.v[4], v.v[5]); q.w = pk2(v.v[6], v.v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pl[0]) + idx) = q;
    if (planes == 2) {
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = v.v[i] - __bfloat162float(__float2bfloat16_rn(v.v[i]));
        q.x = pk2(r[0], r[1]); q.y = pk2(r[2], r[3]); q.z = pk2(r[4], r[5]); q.w = pk2(r[6], r[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pl[1]) + idx) = q;
    }
}

// value represented by the plane(s): hi (+ lo)
__device__ __forceinline__ F8 load8_planes(const uint64_t (&pl)[2], int planes, int64_t idx) {
    F8 r = load8(pl[0], 0, idx);
    if (planes == 2) {
        const F8 l = load8(pl[1], 0, idx);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] += l.v[i];
    }
    return r;
}

// 8 consecutive fp32 per-channel coefficients (32-byte aligned: channel groups start at multiples of 8)
__device__ __forceinline__ F8 load_coef8(uint64_t base, int64_t idx) {
    F8 r;
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ int reflect_idx(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

__device__ __forceinline__ int mirrors_of(int c, int L, int p, int (&out)[3]) {
    int n = 0;
    out[n++] = c;
    if (p > 0) {
        if (c >= 1 && c <= p) out[n++] = -c;
        if (c >= L - 1 - p && c <= L - 2) out[n++] = 2 * (L - 1) - c;
    }
    return n;
}

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
    if (act == ACLGAN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACLGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == ACLGAN_ACT_TANH) return tanhf(v);
    return v;
}

// ------------------------------------------------------------------------------------------ pack_img
__global__ void pack_img_kernel(aclgan_pack_img_args a) {
    const int p = a.dst.pad, hp = a.h + 2 * p, wp = a.w + 2 * p;
    const int64_t total = (int64_t)a.n * hp * wp;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int X = (int)(t % wp);
    const int Y = (int)((t / wp) % hp);
    const int n = (int)(t / ((int64_t)wp * hp));
    const int y = reflect_idx(Y - p, a.h), x = reflect_idx(X - p, a.w);
    const float* s0 = reinterpret_cast<const float*>(a.src0);
    const float* s1 = reinterpret_cast<const float*>(a.src1);
    const int64_t hw = (int64_t)a.h * a.w;
    const int cs = a.dst.c;
    for (int g = 0; g < cs / 8; ++g) {
        F8 v = f8_zero();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            if (c < a.c0) v.v[j] = __ldg(s0 + ((int64_t)n * a.c0 + c) * hw + (int64_t)y * a.w + x);
            else if (c < a.c0 + a.c1) v.v[j] = __ldg(s1 + ((int64_t)n * a.c1 + (c - a.c0)) * hw + (int64_t)y * a.w + x);
        }
        store8_planes(a.dst.data, a.dst.planes, t * cs + g * 8, v);
    }
}

// ------------------------------------------------------------------------------------------ generic NCHW <-> plane
// fp32 NCHW tensor (any channel count) -> reflect-padded NHWC plane(s); stored channels beyond c are zero.
// One thread per (padded pixel, 8-channel group).  Used by the inference API (AdaINGen.decode takes the content code as an
// NCHW tensor: reference networks.py:147-152, test.py:96-106).
__global__ void pack_nchw_kernel(const float* __restrict__ src, int c, int n, int h, int w, aclgan_act dst) {
    const int p = dst.pad, hp = h + 2 * p, wp = w + 2 * p, cg = dst.c / 8;
    const int64_t total = (int64_t)n * hp * wp * cg;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int g = (int)(t % cg);
    const int64_t pix = t / cg;
    const int X = (int)(pix % wp), Y = (int)((pix / wp) % hp), ni = (int)(pix / ((int64_t)wp * hp));
    const int y = reflect_idx(Y - p, h), x = reflect_idx(X - p, w);
    F8 v = f8_zero();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = g * 8 + j;
        if (ch < c) v.v[j] = __ldg(src + (((int64_t)ni * c + ch) * h + y) * w + x);
    }
    store8_planes(dst.data, dst.planes, pix * dst.c + g * 8, v);
}

// plane(s) -> fp32 NCHW [n][c][h][w] (interior, first c channels, hi + lo): thread per (n, c, y, x), x fastest
__global__ void unpack_plane_kernel(aclgan_act src, int c, float* __restrict__ dst) {
    const int64_t total = (int64_t)src.n * c * src.h * src.w;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int x = (int)(t % src.w), y = (int)((t / src.w) % src.h);
    const int ch = (int)((t / ((int64_t)src.w * src.h)) % c), ni = (int)(t / ((int64_t)src.w * src.h * c));
    const int p = src.pad, wp = src.w + 2 * p, hp = src.h + 2 * p;
    const int64_t i = (((int64_t)ni * hp + y + p) * wp + x + p) * src.c + ch;
    float v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src.data[0])[i]);
    if (src.planes == 2) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src.data[1])[i]);
    dst[t] = v;
}

// ------------------------------------------------------------------------------------------ norm_stats
// CTA = 256 threads = (C/8 channel groups) x (pixel lanes); fp32 partials per thread, CTA reduce, fp64 atomics
constexpr int kStatThreads = 256;
constexpr int kStatIters = 8;      // pixels per thread: small ranges -> many CTAs (these kernels are latency / HBM bound)

__global__ void __launch_bounds__(kStatThreads) norm_stats_kernel(aclgan_tensor4 y, double* sums) {
    extern __shared__ float red[];  // [lanes][C][2]
    const int cg = y.c / 8;
    const int lanes = kStatThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const int64_t hw = (int64_t)y.h * y.w;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (lane < lanes) {
        const int64_t p0 = (int64_t)blockIdx.x * lanes * kStatIters + lane;
        F8 buf[kStatIters];
#pragma unroll
        for (int it = 0; it < kStatIters; ++it) {       // all loads first (memory-level parallelism), then the math
            const int64_t pix = p0 + (int64_t)it * lanes;
            buf[it] = pix < hw ? load8(y.ptr, y.kind, ((int64_t)n * hw + pix) * y.c + g * 8) : f8_zero();
        }
#pragma unroll
        for (int it = 0; it < kStatIters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += buf[it].v[i]; q[i] += buf[it].v[i] * buf[it].v[i]; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[((lane * y.c) + g * 8 + i) * 2] = s[i];
            red[((lane * y.c) + g * 8 + i) * 2 + 1] = q[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < y.c; c += kStatThreads) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < lanes; ++l) { a += red[(l * y.c + c) * 2]; b += red[(l * y.c + c) * 2 + 1]; }
        atomicAdd(&sums[((int64_t)n * y.c + c) * 2], a);
        atomicAdd(&sums[((int64_t)n * y.c + c) * 2 + 1], b);
    }
}

// ------------------------------------------------------------------------------------------ norm_finalize
__device__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t += sh[i];
    return t;
}

__global__ void norm_finalize_kernel(aclgan_norm_finalize_args a) {
    __shared__ double sh[32];
    const int c_valid = a.c_valid;
    const int n = blockIdx.x;
    const double* sums = reinterpret_cast<const double*>(a.sums) + (int64_t)n * a.c * 2;
    float* scale = reinterpret_cast<float*>(a.scale) + (int64_t)n * a.c;
    float* shift = reinterpret_cast<float*>(a.shift) + (int64_t)n * a.c;
    float* mean = reinterpret_cast<float*>(a.mean) + (int64_t)n * a.c;
    float* inv = reinterpret_cast<float*>(a.inv) + (int64_t)n * a.c;
    if (a.mode == ACLGAN_NORM_LN) {
        double s1 = 0.0, s2 = 0.0;
        if (a.stat_groups > 1) {        // phase groups of the sub-pixel up-convolution: [n][groups * c_valid][2]
            const int tot = a.stat_groups * c_valid;
            const double* sg = reinterpret_cast<const double*>(a.sums) + (int64_t)n * tot * 2;
            for (int c = threadIdx.x; c < tot; c += blockDim.x) { s1 += sg[2 * c]; s2 += sg[2 * c + 1]; }
        } else {
            for (int c = threadIdx.x; c < c_valid; c += blockDim.x) { s1 += sums[2 * c]; s2 += sums[2 * c + 1]; }
        }
        s1 = block_sum(s1, sh);
        s2 = block_sum(s2, sh);
        const double M = (double)c_valid * a.hw;
        const double mu = s1 / M;
        double var = (s2 - M * mu * mu) / (M - 1.0);
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var);
        const double r = 1.0 / (sd + (double)a.eps);
        if (threadIdx.x == 0) reinterpret_cast<float*>(a.sigma)[n] = (float)sd;
        const float* gam = reinterpret_cast<const float*>(a.w);
        const float* bet = reinterpret_cast<const float*>(a.b);
        for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
            if (c < c_valid) {
                const double sc = (double)gam[c] * r;
                scale[c] = (float)sc;
                shift[c] = (float)((double)bet[c] - mu * sc);
                mean[c] = (float)mu;
                inv[c] = (float)r;
            } else { scale[c] = shift[c] = mean[c] = inv[c] = 0.f; }
        }
        return;
    }
    for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
        if (c >= c_valid) { scale[c] = shift[c] = mean[c] = inv[c] = 0.f; continue; }
        const double mu = sums[2 * c] / a.hw;
        double var = sums[2 * c + 1] / a.hw - mu * mu;
        if (var < 0.0) var = 0.0;
        const double r = 1.0 / sqrt(var + (double)a.eps);
        double sc = r, sf = -mu * r;
        if (a.mode == ACLGAN_NORM_ADAIN) {
            const int64_t ld = a.wb_stride > 0 ? a.wb_stride : c_valid;
            const double w = reinterpret_cast<const float*>(a.w)[(int64_t)n * ld + c];
            const double b = reinterpret_cast<const float*>(a.b)[(int64_t)n * ld + c];
            sc = r * w;
            sf = b - mu * sc;
        }
        scale[c] = (float)sc; shift[c] = (float)sf; mean[c] = (float)mu; inv[c] = (float)r;
    }
}

// ------------------------------------------------------------------------------------------ norm_apply
constexpr int kApplyPix = 4;     // padded pixels per thread (same image, same channel group -> coefficients stay in registers)

template <int K, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_apply_kernel(aclgan_apply_args a) {
    const int u = a.upsample, p = a.dst.pad;
    const int hd = a.y.h * u, wd = a.y.w * u, hp = hd + 2 * p, wp = wd + 2 * p;
    const int cg = a.y.c / 8;
    const int lanes = 256 / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    if (lane >= lanes) return;
    const int npix = hp * wp;                 // (the launcher rejects planes of 2^30 pixels or more)
    F8 sc, sf;
    if (a.scale != 0) {
        sc = load_coef8(a.scale, (int64_t)n * a.y.c + g * 8);
        sf = load_coef8(a.shift, (int64_t)n * a.y.c + g * 8);
    }
    const int rp = a.res.pad, rwp = a.res.w + 2 * rp, rhp = a.res.h + 2 * rp;
    const int batch = lanes * K;
    // CTAs stride over pixel batches of their image; per batch all loads of the thread's pixels are issued first
    // (memory-level parallelism), then the arithmetic and the stores
    for (int p0 = blockIdx.x * batch + lane; p0 - lane < npix; p0 += gridDim.x * batch) {
        F8 yv[K], rv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int pix = p0 + j * lanes;
            if (pix >= npix) pix = npix - 1;
            const int Y = pix / wp, X = pix - Y * wp;
            int y = reflect_idx(Y - p, hd), x = reflect_idx(X - p, wd);
            if (u == 2) { y >>= 1; x >>= 1; }
            yv[j] = load8(a.y.ptr, a.y.kind, (((int64_t)n * a.y.h + y) * a.y.w + x) * a.y.c + g * 8);
            if (a.has_res)
                rv[j] = load8_planes(a.res.data, a.res.planes, (((int64_t)n * rhp + y + rp) * rwp + x + rp) * a.res.c + g * 8);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int pix = p0 + j * lanes;
            if (pix >= npix) break;
            F8 v = yv[j];
            if (a.scale != 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] = v.v[i] * sc.v[i] + sf.v[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] = act_fn(v.v[i], a.act, a.slope);
            if (a.has_res) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] += rv[j].v[i];
            }
            store8_planes(a.dst.data, a.dst.planes, ((int64_t)n * npix + pix) * a.dst.c + g * 8, v);
        }
    }
}

// ------------------------------------------------------------------------------------------ block backward
// gradient w.r.t. the logical block output at (n, y, x), channel group g, times the activation derivative;
// also returns yhat when the block has a norm
struct BwdCoef {
    F8 scale, shift, mean, inv;
};

__device__ __forceinline__ BwdCoef load_bwd_coef(const aclgan_block_bwd_args& a, int n, int g) {
    BwdCoef c;
    const int64_t i = (int64_t)n * a.c + g * 8;
    if (a.mask_mode == ACLGAN_MASK_FROM_Z) { c.scale = load_coef8(a.scale, i); c.shift = load_coef8(a.shift, i); }
    if (a.norm) { c.mean = load_coef8(a.mean, i); c.inv = load_coef8(a.inv, i); }
    return c;
}

__device__ __forceinline__ void block_dz(const aclgan_block_bwd_args& a, const BwdCoef& cf, int n, int y, int x, int g,
                                         F8& dz, F8& yhat) {
    F8 acc = f8_zero();
    if (a.gp != 0) {
        const int u = a.upsample, p = a.gp_pad;
        const int hu = a.h * u, wu = a.w * u, hpp = hu + 2 * p, wpp = wu + 2 * p;
        for (int da = 0; da < u; ++da) {
            int my[3];
            const int ny = mirrors_of(y * u + da, hu, p, my);
            for (int db = 0; db < u; ++db) {
                int mx[3];
                const int nx = mirrors_of(x * u + db, wu, p, mx);
                for (int iy = 0; iy < ny; ++iy)
                    for (int ix = 0; ix < nx; ++ix) {
                        const F8 t = load8(a.gp, a.g_kind,
                                           (((int64_t)n * hpp + my[iy] + p) * wpp + mx[ix] + p) * a.c + g * 8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
                    }
            }
        }
    }
    if (a.gr != 0) {
        const F8 t = load8(a.gr, a.g_kind, (((int64_t)n * a.h + y) * a.w + x) * a.c + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
    }
    F8 yv = f8_zero();
    if (a.mask_mode == ACLGAN_MASK_FROM_Z || a.norm)
        yv = load8(a.y.ptr, a.y.kind, (((int64_t)n * a.h + y) * a.w + x) * a.c + g * 8);
    if (a.mask_mode == ACLGAN_MASK_FROM_Z) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = yv.v[i] * cf.scale.v[i] + cf.shift.v[i];
            if (!(z > 0.f)) acc.v[i] *= a.slope;
        }
    } else if (a.mask_mode == ACLGAN_MASK_FROM_OUT) {
        const int po = a.out.pad, wo = a.out.w + 2 * po, ho = a.out.h + 2 * po;
        const F8 o = load8(a.out.data[0], 0, (((int64_t)n * ho + y + po) * wo + x + po) * a.out.c + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (!(o.v[i] > 0.f)) acc.v[i] *= a.slope;
    }
    dz = acc;
    if (a.norm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) yhat.v[i] = (yv.v[i] - cf.mean.v[i]) * cf.inv.v[i];
    }
}

__global__ void __launch_bounds__(kStatThreads) block_bwd_reduce_kernel(aclgan_block_bwd_args a) {
    extern __shared__ float red[];
    const int cg = a.c / 8;
    const int lanes = kStatThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const int64_t hw = (int64_t)a.h * a.w;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (lane < lanes) {
        const BwdCoef cf = load_bwd_coef(a, n, g);
        const int64_t p0 = (int64_t)blockIdx.x * lanes * kStatIters + lane;
#pragma unroll 4
        for (int it = 0; it < kStatIters; ++it) {
            const int64_t pix = p0 + (int64_t)it * lanes;
            if (pix >= hw) break;
            F8 dz, yh = f8_zero();
            block_dz(a, cf, n, (int)(pix / a.w), (int)(pix % a.w), g, dz, yh);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += dz.v[i]; q[i] += dz.v[i] * yh.v[i]; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[((lane * a.c) + g * 8 + i) * 2] = s[i];
            red[((lane * a.c) + g * 8 + i) * 2 + 1] = q[i];
        }
    }
    __syncthreads();
    double* sums = reinterpret_cast<double*>(a.sums);
    for (int c = threadIdx.x; c < a.c; c += kStatThreads) {
        double x = 0.0, y = 0.0;
        for (int l = 0; l < lanes; ++l) { x += red[(l * a.c + c) * 2]; y += red[(l * a.c + c) * 2 + 1]; }
        atomicAdd(&sums[((int64_t)n * a.c + c) * 2], x);
        atomicAdd(&sums[((int64_t)n * a.c + c) * 2 + 1], y);
    }
}

__global__ void __launch_bounds__(256) block_bwd_apply_kernel(aclgan_block_bwd_args a) {
    const int pz = a.dy.pad, hz = a.h + 2 * pz, wz = a.w + 2 * pz;
    const int cg = a.c / 8;
    const int lanes = 256 / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    if (lane >= lanes) return;
    const int64_t npix = (int64_t)hz * wz;
    const BwdCoef cf = load_bwd_coef(a, n, g);
    F8 ca, cb, cc;
    if (a.norm) {
        const int64_t ci = (int64_t)n * a.c + g * 8;
        ca = load_coef8(a.ca, ci); cb = load_coef8(a.cb, ci); cc = load_coef8(a.cc, ci);
    }
    const int64_t p0 = (int64_t)blockIdx.x * lanes * kApplyPix + lane;
#pragma unroll
    for (int j = 0; j < kApplyPix; ++j) {
        const int64_t pix = p0 + (int64_t)j * lanes;
        if (pix >= npix) break;
        const int y = (int)(pix / wz) - pz, x = (int)(pix % wz) - pz;
        F8 out = f8_zero();
        if (y >= 0 && y < a.h && x >= 0 && x < a.w) {
            F8 dz, yh = f8_zero();
            block_dz(a, cf, n, y, x, g, dz, yh);
            if (a.norm) {
#pragma unroll
                for (int i = 0; i < 8; ++i) out.v[i] = ca.v[i] * dz.v[i] + cb.v[i] * yh.v[i] + cc.v[i];
            } else {
                out = dz;
            }
        }
        store8_planes(a.dy.data, a.dy.planes, ((int64_t)n * npix + pix) * a.c + g * 8, out);
    }
}

// ---- fast variants (every plane extent > 2 * pad + 1, so a coordinate has at most ONE reflect image) ----------------
// All unconditional loads of a thread's K pixels are issued before any arithmetic (memory-level parallelism: these
// kernels are pure HBM/L2 streams); the reflect images of border pixels are added under a rare, mostly warp-uniform branch.
__device__ __forceinline__ bool mirror_of1(int c, int L, int p, int& m) {
    if (c >= 1 && c <= p) { m = -c; return true; }
    if (c >= L - 1 - p && c <= L - 2) { m = 2 * (L - 1) - c; return true; }
    return false;
}

template <int U>
__device__ __forceinline__ int64_t gp_index(const aclgan_block_bwd_args& a, int n, int cy, int cx, int g) {
    const int p = a.gp_pad, hpp = a.h * U + 2 * p, wpp = a.w * U + 2 * p;
    return (((int64_t)n * hpp + cy + p) * wpp + cx + p) * a.c + g * 8;
}

// adds the reflect images of the U x U source pixels of (y, x); returns false when there are none
template <int U>
__device__ __forceinline__ void add_mirrors(const aclgan_block_bwd_args& a, int n, int y, int x, int g, F8& acc) {
    const int p = a.gp_pad, hu = a.h * U, wu = a.w * U;
#pragma unroll
    for (int da = 0; da < U; ++da) {
        const int cy = y * U + da;
        int my = 0;
        const bool hy = mirror_of1(cy, hu, p, my);
#pragma unroll
        for (int db = 0; db < U; ++db) {
            const int cx = x * U + db;
            int mx = 0;
            const bool hx = mirror_of1(cx, wu, p, mx);
            if (hy) {
                const F8 t = load8(a.gp, a.g_kind, gp_index<U>(a, n, my, cx, g));
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
            }
            if (hx) {
                const F8 t = load8(a.gp, a.g_kind, gp_index<U>(a, n, cy, mx, g));
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
            }
            if (hy && hx) {
                const F8 t = load8(a.gp, a.g_kind, gp_index<U>(a, n, my, mx, g));
#pragma unroll
                for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
            }
        }
    }
}

template <int U>
__device__ __forceinline__ bool has_mirrors(const aclgan_block_bwd_args& a, int y, int x) {
    const int p = a.gp_pad, hu = a.h * U, wu = a.w * U;
    if (p == 0) return false;
    const int y0 = y * U, y1 = y * U + U - 1, x0 = x * U, x1 = x * U + U - 1;
    return (y0 <= p && y1 >= 1) || (y1 >= hu - 1 - p && y0 <= hu - 2) || (x0 <= p && x1 >= 1) ||
           (x1 >= wu - 1 - p && x0 <= wu - 2);
}

// raw 8-channel vectors as loaded (bf16: 16 B = 4 registers, fp32: 32 B), expanded to fp32 only when consumed, so the
// K pixels a thread keeps in flight stay cheap in registers
template <int KIND> struct Raw8;
template <> struct Raw8<0> { uint4 q; };
template <> struct Raw8<1> { float4 a, b; };

template <int KIND>
__device__ __forceinline__ Raw8<KIND> raw_load(uint64_t base, int64_t idx);
template <>
__device__ __forceinline__ Raw8<0> raw_load<0>(uint64_t base, int64_t idx) {
    Raw8<0> r;
    r.q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    return r;
}
template <>
__device__ __forceinline__ Raw8<1> raw_load<1>(uint64_t base, int64_t idx) {
    Raw8<1> r;
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    r.a = p[0]; r.b = p[1];
    return r;
}
__device__ __forceinline__ F8 raw_f8(const Raw8<0>& r) {
    F8 o;
    const uint32_t w[4] = {r.q.x, r.q.y, r.q.z, r.q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.v[2 * i] = __uint_as_float(w[i] << 16);
        o.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
    return o;
}
__device__ __forceinline__ F8 raw_f8(const Raw8<1>& r) {
    F8 o;
    o.v[0] = r.a.x; o.v[1] = r.a.y; o.v[2] = r.a.z; o.v[3] = r.a.w;
    o.v[4] = r.b.x; o.v[5] = r.b.y; o.v[6] = r.b.z; o.v[7] = r.b.w;
    return o;
}

// loads of one pixel that do not depend on anything else: gradient sources, raw conv output, forward output
template <int U, int KIND>
struct PixIn {
    Raw8<KIND> g[U * U];
    Raw8<KIND> gr, yv;
    Raw8<0> ov;
};

template <int U, int KIND>
__device__ __forceinline__ void pix_load(const aclgan_block_bwd_args& a, int n, int y, int x, int g, PixIn<U, KIND>& in) {
    if (a.gp != 0) {
#pragma unroll
        for (int da = 0; da < U; ++da)
#pragma unroll
            for (int db = 0; db < U; ++db)
                in.g[da * U + db] = raw_load<KIND>(a.gp, gp_index<U>(a, n, y * U + da, x * U + db, g));
    }
    const int64_t di = (((int64_t)n * a.h + y) * a.w + x) * a.c + g * 8;
    if (a.gr != 0) in.gr = raw_load<KIND>(a.gr, di);
    if (a.mask_mode == ACLGAN_MASK_FROM_Z || a.norm) in.yv = raw_load<KIND>(a.y.ptr, di);
    if (a.mask_mode == ACLGAN_MASK_FROM_OUT) {
        const int po = a.out.pad, wo = a.out.w + 2 * po, ho = a.out.h + 2 * po;
        in.ov = raw_load<0>(a.out.data[0], (((int64_t)n * ho + y + po) * wo + x + po) * a.out.c + g * 8);
    }
}

template <int U, int KIND>
__device__ __forceinline__ void pix_dz(const aclgan_block_bwd_args& a, const BwdCoef& cf, int n, int y, int x, int g,
                                       const PixIn<U, KIND>& in, F8& dz, F8& yhat) {
    F8 acc = f8_zero();
    if (a.gp != 0) {
#pragma unroll
        for (int k = 0; k < U * U; ++k) {
            const F8 t = raw_f8(in.g[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
        }
        if (has_mirrors<U>(a, y, x)) add_mirrors<U>(a, n, y, x, g, acc);
    }
    if (a.gr != 0) {
        const F8 t = raw_f8(in.gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
    }
    F8 yv = f8_zero();
    if (a.mask_mode == ACLGAN_MASK_FROM_Z || a.norm) yv = raw_f8(in.yv);
    if (a.mask_mode == ACLGAN_MASK_FROM_Z) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = yv.v[i] * cf.scale.v[i] + cf.shift.v[i];
            if (!(z > 0.f)) acc.v[i] *= a.slope;
        }
    } else if (a.mask_mode == ACLGAN_MASK_FROM_OUT) {
        const F8 ov = raw_f8(in.ov);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (!(ov.v[i] > 0.f)) acc.v[i] *= a.slope;
    }
    dz = acc;
    if (a.norm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) yhat.v[i] = (yv.v[i] - cf.mean.v[i]) * cf.inv.v[i];
    }
}

// CTA-level reduction of per-thread channel partials (s, q) -> fp64 atomics (stat != null) / fp32 atomics (dbias != null)
__device__ __forceinline__ void cta_channel_reduce(float* red, const float (&s)[8], const float (&q)[8], int c, int g, int lane,
                                                   int lanes, bool active, double* stat, float* dbias, int dbias_n) {
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[((lane * c) + g * 8 + i) * 2] = s[i];
            red[((lane * c) + g * 8 + i) * 2 + 1] = q[i];
        }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float x = 0.f, y = 0.f;
        for (int l = 0; l < lanes; ++l) { x += red[(l * c + ch) * 2]; y += red[(l * c + ch) * 2 + 1]; }
        if (stat != nullptr) {
            atomicAdd(&stat[ch * 2], (double)x);
            atomicAdd(&stat[ch * 2 + 1], (double)y);
        }
        if (dbias != nullptr && ch < dbias_n) atomicAdd(dbias + ch, x);
    }
}

constexpr int kBwdPix1 = 4;     // pixels per thread, no upsample
constexpr int kBwdPix2 = 2;     // pixels per thread, 2x upsample (4 gradient sources per pixel)

// per-(image, channel) coefficient rows staged in shared memory (one image per CTA): keeps ~56 registers per thread free
// for loads in flight.  Row r of `coef` holds c floats.
enum { CF_SCALE = 0, CF_SHIFT, CF_MEAN, CF_INV, CF_CA, CF_CB, CF_CC, CF_ROWS };

__device__ __forceinline__ void stage_coefs(const aclgan_block_bwd_args& a, int n, float* coef, bool apply) {
    const uint64_t src[CF_ROWS] = {a.mask_mode == ACLGAN_MASK_FROM_Z ? a.scale : 0, a.mask_mode == ACLGAN_MASK_FROM_Z ? a.shift : 0,
                                   a.norm ? a.mean : 0, a.norm ? a.inv : 0, (a.norm && apply) ? a.ca : 0,
                                   (a.norm && apply) ? a.cb : 0, (a.norm && apply) ? a.cc : 0};
#pragma unroll
    for (int r = 0; r < CF_ROWS; ++r) {
        if (src[r] == 0) continue;
        const float* p = reinterpret_cast<const float*>(src[r]) + (int64_t)n * a.c;
        for (int i = threadIdx.x; i < a.c; i += blockDim.x) coef[r * a.c + i] = __ldg(p + i);
    }
    __syncthreads();
}

__device__ __forceinline__ F8 coef8(const float* coef, int row, int c, int g) {
    F8 r;
    const float4* p = reinterpret_cast<const float4*>(coef + row * c + g * 8);
    const float4 x = p[0], y = p[1];
    r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w;
    r.v[4] = y.x; r.v[5] = y.y; r.v[6] = y.z; r.v[7] = y.w;
    return r;
}

// dz (and yhat) of one pixel from its raw loads; coefficients come from shared memory
template <int U, int KIND>
__device__ __forceinline__ void pix_dz_s(const aclgan_block_bwd_args& a, const float* coef, int n, int y, int x, int g,
                                         const PixIn<U, KIND>& in, F8& dz, F8& yhat) {
    F8 acc = f8_zero();
    if (a.gp != 0) {
#pragma unroll
        for (int k = 0; k < U * U; ++k) {
            const F8 t = raw_f8(in.g[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
        }
        if (has_mirrors<U>(a, y, x)) add_mirrors<U>(a, n, y, x, g, acc);
    }
    if (a.gr != 0) {
        const F8 t = raw_f8(in.gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
    }
    F8 yv = f8_zero();
    if (a.mask_mode == ACLGAN_MASK_FROM_Z || a.norm) yv = raw_f8(in.yv);
    if (a.mask_mode == ACLGAN_MASK_FROM_Z) {
        const F8 sc = coef8(coef, CF_SCALE, a.c, g), sf = coef8(coef, CF_SHIFT, a.c, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = yv.v[i] * sc.v[i] + sf.v[i];
            if (!(z > 0.f)) acc.v[i] *= a.slope;
        }
    } else if (a.mask_mode == ACLGAN_MASK_FROM_OUT) {
        const F8 ov = raw_f8(in.ov);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (!(ov.v[i] > 0.f)) acc.v[i] *= a.slope;
    }
    dz = acc;
    if (a.norm) {
        const F8 mu = coef8(coef, CF_MEAN, a.c, g), iv = coef8(coef, CF_INV, a.c, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) yhat.v[i] = (yv.v[i] - mu.v[i]) * iv.v[i];
    }
}

template <int U, int KIND, int MINB, int KP>
__global__ void __launch_bounds__(kStatThreads, MINB) block_bwd_reduce_fast(aclgan_block_bwd_args a) {
    constexpr int K = (U == 1) ? KP : (KP > 2 ? 2 : 1);
    extern __shared__ float smem_f[];
    float* coef = smem_f;                       // [CF_ROWS][c]
    float* red = smem_f + CF_ROWS * a.c;        // [lanes][c][2]
    const int cg = a.c / 8;
    const int lanes = kStatThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const int hw = a.h * a.w;
    stage_coefs(a, n, coef, false);
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const bool active = lane < lanes;
    if (active) {
        for (int p0 = blockIdx.x * lanes * K + lane; p0 - lane < hw; p0 += gridDim.x * lanes * K) {
            PixIn<U, KIND> in[K];
            int py[K], px[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                int pix = p0 + j * lanes;
                if (pix >= hw) pix = hw - 1;            // clamped duplicate: loaded, not counted
                py[j] = pix / a.w; px[j] = pix - py[j] * a.w;
                pix_load<U, KIND>(a, n, py[j], px[j], g, in[j]);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                F8 dz, yh = f8_zero();
                pix_dz_s<U, KIND>(a, coef, n, py[j], px[j], g, in[j], dz, yh);
                const float wgt = (p0 + j * lanes < hw) ? 1.f : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += wgt * dz.v[i]; q[i] += wgt * dz.v[i] * yh.v[i]; }
            }
        }
    }
    cta_channel_reduce(red, s, q, a.c, g, lane, lanes, active, reinterpret_cast<double*>(a.sums) + (int64_t)n * a.c * 2,
                       nullptr, 0);
}

template <int U, int KIND, int MINB, int KP>
__global__ void __launch_bounds__(kStatThreads, MINB) block_bwd_apply_fast(aclgan_block_bwd_args a) {
    constexpr int K = (U == 1) ? KP : (KP > 2 ? 2 : 1);
    extern __shared__ float smem_f[];
    float* coef = smem_f;
    float* red = smem_f + CF_ROWS * a.c;
    const int pz = a.dy.pad, hz = a.h + 2 * pz, wz = a.w + 2 * pz;
    const int cg = a.c / 8;
    const int lanes = kStatThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const int npix = hz * wz;
    const bool active = lane < lanes;
    stage_coefs(a, n, coef, true);
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (active) {
        for (int p0 = blockIdx.x * lanes * K + lane; p0 - lane < npix; p0 += gridDim.x * lanes * K) {
            PixIn<U, KIND> in[K];
            int py[K], px[K];
            bool inside[K];
            // loads are unconditional (border / tail pixels load a clamped in-range pixel and discard it): no divergent
            // control flow between the loads, so all of them are in flight before the first use
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int pix = p0 + j * lanes;
                const int Y = pix / wz;
                const int y = Y - pz, x = pix - Y * wz - pz;
                inside[j] = pix < npix && y >= 0 && y < a.h && x >= 0 && x < a.w;
                py[j] = min(max(y, 0), a.h - 1);
                px[j] = min(max(x, 0), a.w - 1);
                pix_load<U, KIND>(a, n, py[j], px[j], g, in[j]);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int pix = p0 + j * lanes;
                F8 dz, yh = f8_zero();
                pix_dz_s<U, KIND>(a, coef, n, py[j], px[j], g, in[j], dz, yh);
                F8 out;
                if (a.norm) {
                    const F8 ca = coef8(coef, CF_CA, a.c, g), cb = coef8(coef, CF_CB, a.c, g), cc = coef8(coef, CF_CC, a.c, g);
#pragma unroll
                    for (int i = 0; i < 8; ++i) out.v[i] = ca.v[i] * dz.v[i] + cb.v[i] * yh.v[i] + cc.v[i];
                } else {
                    out = dz;
                }
                const float wgt = inside[j] ? 1.f : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) { out.v[i] *= wgt; s[i] += wgt * dz.v[i]; }
                if (pix < npix) store8_planes(a.dy.data, a.dy.planes, ((int64_t)n * npix + pix) * a.c + g * 8, out);
            }
        }
    }
    if (a.dbias != 0)      // bias gradient of a block without norm: sum of dz over (n, h, w)
        cta_channel_reduce(red, s, q, a.c, g, lane, lanes, active, nullptr, reinterpret_cast<float*>(a.dbias), a.dbias_n);
}

__global__ void norm_bwd_finalize_kernel(aclgan_norm_bwd_finalize_args a) {
    __shared__ double sh[32];
    const int c_valid = a.c_valid;
    const int n = blockIdx.x;
    const double* sums = reinterpret_cast<const double*>(a.sums) + (int64_t)n * a.c * 2;
    const float* inv = reinterpret_cast<const float*>(a.inv) + (int64_t)n * a.c;
    float* ca = reinterpret_cast<float*>(a.ca) + (int64_t)n * a.c;
    float* cb = reinterpret_cast<float*>(a.cb) + (int64_t)n * a.c;
    float* cc = reinterpret_cast<float*>(a.cc) + (int64_t)n * a.c;
    if (a.mode == ACLGAN_NORM_LN) {
        const float* gam = reinterpret_cast<const float*>(a.w);
        double s1 = 0.0, s2 = 0.0;
        for (int c = threadIdx.x; c < c_valid; c += blockDim.x) {
            s1 += (double)gam[c] * sums[2 * c];
            s2 += (double)gam[c] * sums[2 * c + 1];
        }
        s1 = block_sum(s1, sh);
        s2 = block_sum(s2, sh);
        const double M = (double)c_valid * a.hw;
        const double sd = reinterpret_cast<const float*>(a.sigma)[n];
        for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
            if (c >= c_valid) { ca[c] = cb[c] = cc[c] = 0.f; continue; }
            const double r = inv[c];
            ca[c] = (float)(r * gam[c]);
            cb[c] = (float)(sd > 0.0 ? -s2 / ((M - 1.0) * sd) : 0.0);
            cc[c] = (float)(-r * s1 / M);
            atomicAdd(reinterpret_cast<float*>(a.dw) + c, (float)sums[2 * c + 1]);
            atomicAdd(reinterpret_cast<float*>(a.db) + c, (float)sums[2 * c]);
            if (a.dbias != 0) {
                // conv bias in front of LayerNorm: db[c] = sum over (n, hw) of dy = ca*T1 + cb*sum_hw(yhat) + cc*HW
                const double mu = reinterpret_cast<const float*>(a.mean)[(int64_t)n * a.c + c];
                double s1f;
                if (a.fstat_groups > 1) {
                    s1f = 0.0;
                    for (int gph = 0; gph < a.fstat_groups; ++gph)
                        s1f += reinterpret_cast<const double*>(a.fsums)[(((int64_t)n * a.fstat_groups + gph) * c_valid + c) * 2];
                } else {
                    s1f = reinterpret_cast<const double*>(a.fsums)[((int64_t)n * a.c + c) * 2];
                }
                const double syh = (s1f - (double)a.hw * mu) * r;
                const double dbv = (double)ca[c] * sums[2 * c] + (double)cb[c] * syh + (double)cc[c] * (double)a.hw;
                atomicAdd(reinterpret_cast<float*>(a.dbias) + c, (float)dbv);
            }
        }
        return;
    }
    for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
        if (c >= c_valid) { ca[c] = cb[c] = cc[c] = 0.f; continue; }
        double g = 1.0;
        if (a.mode == ACLGAN_NORM_ADAIN) {
            const int64_t ld = a.wb_stride > 0 ? a.wb_stride : c_valid;
            g = reinterpret_cast<const float*>(a.w)[(int64_t)n * ld + c];
            reinterpret_cast<float*>(a.dw)[(int64_t)n * ld + c] = (float)sums[2 * c + 1];
            reinterpret_cast<float*>(a.db)[(int64_t)n * ld + c] = (float)sums[2 * c];
        }
        const double r = inv[c];
        ca[c] = (float)(r * g);
        cb[c] = (float)(-r * g * sums[2 * c + 1] / a.hw);
        cc[c] = (float)(-r * g * sums[2 * c] / a.hw);
    }
}

// ------------------------------------------------------------------------------------------ image gradients
__global__ void img_grad_pack_kernel(aclgan_img_grad_pack_args a) {
    __shared__ float red[8][8];
    const int pz = a.dy.pad, hz = a.h + 2 * pz, wz = a.w + 2 * pz;
    const int64_t total = (int64_t)a.n * hz * wz;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    F8 v = f8_zero();
    if (t < total) {
        const int X = (int)(t % wz);
        const int Y = (int)((t / wz) % hz);
        const int n = (int)(t / ((int64_t)wz * hz));
        const int y = Y - pz, x = X - pz;
        if (y >= 0 && y < a.h && x >= 0 && x < a.w) {
            const float* d = reinterpret_cast<const float*>(a.dimg);
            const float* o = reinterpret_cast<const float*>(a.out_img);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < a.c) {
                    const int64_t i = (((int64_t)n * a.c + c) * a.h + y) * a.w + x;
                    float gv = __ldg(d + i);
                    if (o != nullptr) { const float ov = __ldg(o + i); gv *= (1.f - ov * ov); }
                    v.v[c] = gv;
                }
        }
        store8_planes(a.dy.data, a.dy.planes, t * 8, v);
    }
    if (a.dbias != 0) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float s = v.v[c];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[w][c] = s;
        }
        __syncthreads();
        if (threadIdx.x < a.c) {
            float s = 0.f;
            for (int i = 0; i < (int)blockDim.x / 32; ++i) s += red[i][threadIdx.x];
            atomicAdd(reinterpret_cast<float*>(a.dbias) + threadIdx.x, s);
        }
    }
}

__global__ void img_grad_unpack_kernel(aclgan_img_grad_unpack_args a) {
    const int64_t total = (int64_t)a.n * a.h * a.w;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int x = (int)(t % a.w);
    const int y = (int)((t / a.w) % a.h);
    const int n = (int)(t / ((int64_t)a.w * a.h));
    const int p = a.pad, hp = a.h + 2 * p, wp = a.w + 2 * p;
    int my[3], mx[3];
    const int ny = mirrors_of(y, a.h, p, my), nx = mirrors_of(x, a.w, p, mx);
    const float* s = reinterpret_cast<const float*>(a.src);
    float* d = reinterpret_cast<float*>(a.dst);
    for (int c = 0; c < a.c; ++c) {
        float acc = 0.f;
        for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix)
                acc += __ldg(s + (((int64_t)n * hp + my[iy] + p) * wp + mx[ix] + p) * a.cs + a.c_off + c);
        const int64_t o = (((int64_t)n * a.c + c) * a.h + y) * a.w + x;
        d[o] = a.accumulate ? d[o] + acc : acc;
    }
}

__global__ void pack_weight_kernel(aclgan_pack_weight_args a) {
    const int64_t total = (int64_t)a.co * a.ci * a.kh * a.kw;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int kw = (int)(t % a.kw);
    const int kh = (int)((t / a.kw) % a.kh);
    const int ci = (int)((t / ((int64_t)a.kw * a.kh)) % a.ci);
    const int co = (int)(t / ((int64_t)a.kw * a.kh * a.ci));
    const float v = reinterpret_cast<const float*>(a.w)[t];
    const int64_t o = a.base + co * a.s_co + ci * a.s_ci + kh * a.s_kh + kw * a.s_kw;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(a.dst[0])[o] = hi;
    if (a.planes == 2) reinterpret_cast<__nv_bfloat16*>(a.dst[1])[o] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

static inline int grid_for(int64_t total, int block) { return (int)((total + block - 1) / block); }

// CTAs per image of the grid-stride element-wise kernels: about four CTAs per SM over all images (two resident at a
// time), so a CTA lives long enough to amortise its prologue and the per-CTA reductions / atomics stay few
static inline int strided_grid(int batches, int n_images, int ctas_per_sm = 4) {
    int per = (ctas_per_sm * num_sms() + n_images - 1) / n_images;
    if (per < 1) per = 1;
    return batches < per ? batches : per;
}

}  // namespace aclgan

using namespace aclgan;

static int check_cg(int c) {
    if (c % 8 != 0 || c / 8 > kStatThreads) return ACLGAN_ERR_SHAPE;
    return 0;
}

extern "C" int aclgan_pack_img(const aclgan_pack_img_args* a, void* stream) {
    if (a->dst.c != 8 && a->dst.c != 16) return ACLGAN_ERR_SHAPE;
    if (a->c0 + a->c1 > a->dst.c) return ACLGAN_ERR_SHAPE;
    const int64_t total = (int64_t)a->n * (a->h + 2 * a->dst.pad) * (a->w + 2 * a->dst.pad);
    pack_img_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_pack_nchw(uint64_t src, int32_t c, const aclgan_act* dst, void* stream) {
    if (dst->c % 8 != 0 || c < 1 || c > dst->c || dst->h <= dst->pad || dst->w <= dst->pad) return ACLGAN_ERR_SHAPE;
    const int64_t total = (int64_t)dst->n * (dst->h + 2 * dst->pad) * (dst->w + 2 * dst->pad) * (dst->c / 8);
    pack_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(src), c, dst->n, dst->h,
                                                                           dst->w, *dst);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_unpack_plane(const aclgan_act* src, int32_t c, uint64_t dst, void* stream) {
    if (c < 1 || c > src->c) return ACLGAN_ERR_SHAPE;
    const int64_t total = (int64_t)src->n * c * src->h * src->w;
    unpack_plane_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, c, reinterpret_cast<float*>(dst));
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_stats(const aclgan_tensor4* y, uint64_t sums, void* stream) {
    if (check_cg(y->c)) return ACLGAN_ERR_SHAPE;
    const int lanes = kStatThreads / (y->c / 8);
    const int64_t hw = (int64_t)y->h * y->w;
    dim3 grid(grid_for(hw, lanes * kStatIters), y->n);
    const size_t smem = (size_t)lanes * y->c * 2 * sizeof(float);
    norm_stats_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>(*y, reinterpret_cast<double*>(sums));
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_finalize(const aclgan_norm_finalize_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->c) return ACLGAN_ERR_SHAPE;
    norm_finalize_kernel<<<a->n, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_apply(const aclgan_apply_args* a, void* stream) {
    if (a->y.c % 8 || a->dst.c != a->y.c || (a->upsample != 1 && a->upsample != 2)) return ACLGAN_ERR_SHAPE;
    if (a->has_res && (a->res.c != a->y.c || a->res.h != a->y.h || a->res.w != a->y.w)) return ACLGAN_ERR_SHAPE;
    if (check_cg(a->y.c)) return ACLGAN_ERR_SHAPE;
    const int u = a->upsample, p = a->dst.pad;
    const int64_t npix = (int64_t)(a->y.h * u + 2 * p) * (a->y.w * u + 2 * p);
    if (npix >= (1LL << 30)) return ACLGAN_ERR_SHAPE;
    {
        const int rc = rows_norm_apply(a, (cudaStream_t)stream);        // row-structured fast path (elementwise_rows.cu)
        if (rc != -100) return rc;
    }
    const int lanes = 256 / (a->y.c / 8);
    static int variant = -1;
    if (variant < 0) {
        const char* e = getenv("ACLGAN_APPLY_VARIANT");
        variant = e != nullptr ? atoi(e) : 1;
    }
    const int K = variant == 0 ? 4 : 2;
    dim3 grid(strided_grid(grid_for(npix, lanes * K), a->y.n, variant == 0 ? 4 : 8), a->y.n);
    if (variant == 0) norm_apply_kernel<4, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    else if (variant == 1) norm_apply_kernel<2, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    else norm_apply_kernel<2, 3><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_finalize_apply(const aclgan_norm_finalize_args* f, const aclgan_apply_args* a, void* stream) {
    if (f->c_valid < 1 || f->c_valid > f->c) return ACLGAN_ERR_SHAPE;
    if (a->y.c % 8 || a->dst.c != a->y.c || (a->upsample != 1 && a->upsample != 2)) return ACLGAN_ERR_SHAPE;
    if (a->has_res && (a->res.c != a->y.c || a->res.h != a->y.h || a->res.w != a->y.w)) return ACLGAN_ERR_SHAPE;
    if (a->scale != f->scale || a->shift != f->shift) return ACLGAN_ERR_SHAPE;      // the apply pass uses exactly the finalized coefficients
    if (check_cg(a->y.c) == 0) {
        const int rc = rows_norm_finalize_apply(f, a, (cudaStream_t)stream);       // ONE launch (elementwise_rows.cu)
        if (rc != -100) return rc;
    }
    const int rc = aclgan_norm_finalize(f, stream);
    if (rc) return rc;
    return aclgan_norm_apply(a, stream);
}

// variant 0: 4 pixels / thread, 2 CTAs / SM (128 registers); 1: 2 pixels, 3 CTAs (85); 2: 2 pixels, 4 CTAs (64)
static int bwd_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ACLGAN_BWD_VARIANT");
        v = e != nullptr ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}
static int bwd_pix(int upsample) {
    const int kp = bwd_variant() == 0 ? 4 : 2;
    return upsample == 1 ? kp : (kp > 2 ? 2 : 1);
}

// fast kernels: at most one reflect image per coordinate, 32-bit pixel indices
static bool bwd_fast_ok(const aclgan_block_bwd_args* a) {
    const int u = a->upsample, p = a->gp_pad;
    if (u != 1 && u != 2) return false;
    if ((a->mask_mode == ACLGAN_MASK_FROM_Z || a->norm) && a->y.kind != a->g_kind) return false;
    if (a->gp != 0 && (a->h * u <= 2 * p + 1 || a->w * u <= 2 * p + 1)) return false;
    if ((int64_t)(a->h + 2 * a->dy.pad) * (a->w + 2 * a->dy.pad) >= (1LL << 30)) return false;
    return getenv("ACLGAN_BWD_GENERIC") == nullptr;
}

extern "C" int aclgan_block_bwd_reduce(const aclgan_block_bwd_args* a, void* stream) {
    if (check_cg(a->c)) return ACLGAN_ERR_SHAPE;
    {
        const int rc = rows_bwd_reduce(a, (cudaStream_t)stream);
        if (rc != -100) return rc;
    }
    const int lanes = kStatThreads / (a->c / 8);
    const int64_t hw = (int64_t)a->h * a->w;
    if (bwd_fast_ok(a)) {
        const int K = bwd_pix(a->upsample);
        dim3 grid(strided_grid(grid_for(hw, lanes * K), a->n, 2 * (bwd_variant() + 2)), a->n);
        const size_t smem = ((size_t)lanes * a->c * 2 + (size_t)CF_ROWS * a->c) * sizeof(float);
        const int sel = (a->upsample == 1 ? 0 : 2) + (a->g_kind ? 1 : 0) + 4 * bwd_variant();
        cudaStream_t st = (cudaStream_t)stream;
        switch (sel) {
            case 0: block_bwd_reduce_fast<1, 0, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 1: block_bwd_reduce_fast<1, 1, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 2: block_bwd_reduce_fast<2, 0, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 3: block_bwd_reduce_fast<2, 1, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 4: block_bwd_reduce_fast<1, 0, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 5: block_bwd_reduce_fast<1, 1, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 6: block_bwd_reduce_fast<2, 0, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 7: block_bwd_reduce_fast<2, 1, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 8: block_bwd_reduce_fast<1, 0, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 9: block_bwd_reduce_fast<1, 1, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 10: block_bwd_reduce_fast<2, 0, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            default: block_bwd_reduce_fast<2, 1, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
        }
        return (int)cudaGetLastError();
    }
    dim3 grid(grid_for(hw, lanes * kStatIters), a->n);
    const size_t smem = (size_t)lanes * a->c * 2 * sizeof(float);
    block_bwd_reduce_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_block_bwd_apply(const aclgan_block_bwd_args* a, void* stream) {
    if (check_cg(a->c) || a->dy.c != a->c) return ACLGAN_ERR_SHAPE;
    {
        const int rc = rows_bwd_apply(a, (cudaStream_t)stream);
        if (rc != -100) return rc;
    }
    const int64_t npix = (int64_t)(a->h + 2 * a->dy.pad) * (a->w + 2 * a->dy.pad);
    const int lanes = 256 / (a->c / 8);
    if (bwd_fast_ok(a)) {
        const int K = bwd_pix(a->upsample);
        dim3 grid(strided_grid(grid_for(npix, lanes * K), a->n, 2 * (bwd_variant() + 2)), a->n);
        const size_t smem = ((a->dbias != 0 ? (size_t)lanes * a->c * 2 : 0) + (size_t)CF_ROWS * a->c) * sizeof(float);
        const int sel = (a->upsample == 1 ? 0 : 2) + (a->g_kind ? 1 : 0) + 4 * bwd_variant();
        cudaStream_t st = (cudaStream_t)stream;
        switch (sel) {
            case 0: block_bwd_apply_fast<1, 0, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 1: block_bwd_apply_fast<1, 1, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 2: block_bwd_apply_fast<2, 0, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 3: block_bwd_apply_fast<2, 1, 2, 4><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 4: block_bwd_apply_fast<1, 0, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 5: block_bwd_apply_fast<1, 1, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 6: block_bwd_apply_fast<2, 0, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 7: block_bwd_apply_fast<2, 1, 3, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 8: block_bwd_apply_fast<1, 0, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 9: block_bwd_apply_fast<1, 1, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            case 10: block_bwd_apply_fast<2, 0, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
            default: block_bwd_apply_fast<2, 1, 4, 2><<<grid, kStatThreads, smem, st>>>(*a); break;
        }
        return (int)cudaGetLastError();
    }
    if (a->dbias != 0) return ACLGAN_ERR_UNSUPPORTED;     // the generic kernel has no fused bias gradient
    dim3 grid(grid_for(npix, lanes * kApplyPix), a->n);
    block_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_bwd_finalize(const aclgan_norm_bwd_finalize_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->c) return ACLGAN_ERR_SHAPE;
    norm_bwd_finalize_kernel<<<a->n, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_bwd_finalize_apply(const aclgan_norm_bwd_finalize_args* f, const aclgan_block_bwd_args* a, void* stream) {
    if (f->c_valid < 1 || f->c_valid > f->c) return ACLGAN_ERR_SHAPE;
    if (check_cg(a->c) || a->dy.c != a->c) return ACLGAN_ERR_SHAPE;
    if (a->ca != f->ca || a->cb != f->cb || a->cc != f->cc || a->sums != f->sums) return ACLGAN_ERR_SHAPE;
    {
        const int rc = rows_bwd_finalize_apply(f, a, (cudaStream_t)stream);        // ONE launch (elementwise_rows.cu)
        if (rc != -100) return rc;
    }
    const int rc = aclgan_norm_bwd_finalize(f, stream);
    if (rc) return rc;
    return aclgan_block_bwd_apply(a, stream);
}

extern "C" int aclgan_img_grad_pack(const aclgan_img_grad_pack_args* a, void* stream) {
    if (a->dy.c != 8 || a->c > 8) return ACLGAN_ERR_SHAPE;
    const int64_t total = (int64_t)a->n * (a->h + 2 * a->dy.pad) * (a->w + 2 * a->dy.pad);
    img_grad_pack_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_img_grad_unpack(const aclgan_img_grad_unpack_args* a, void* stream) {
    const int64_t total = (int64_t)a->n * a->h * a->w;
    img_grad_unpack_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_pack_weight(const aclgan_pack_weight_args* a, void* stream) {
    const int64_t total = (int64_t)a->co * a->ci * a->kh * a->kw;
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}
