// HBM-bound kernels around the tensor-core convolutions: image (un)packing, InstanceNorm / AdaIN / LayerNorm
// statistics + apply (fused with activation, residual add, nearest 2x upsample and the consumer's reflect pad),
// and their backward passes (fold of the reflect/upsample gather, activation mask, two-moment norm backward).
// All work on 8-channel (16 B bf16 / 32 B fp32) vectors of NHWC planes; reductions go warp/CTA-local first and
// then to fp64 atomics so the statistics do not suffer from fp32 cancellation.
#include "common.cuh"

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
        for (int c = threadIdx.x; c < c_valid; c += blockDim.x) { s1 += sums[2 * c]; s2 += sums[2 * c + 1]; }
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
            const double w = reinterpret_cast<const float*>(a.w)[(int64_t)n * c_valid + c];
            const double b = reinterpret_cast<const float*>(a.b)[(int64_t)n * c_valid + c];
            sc = r * w;
            sf = b - mu * sc;
        }
        scale[c] = (float)sc; shift[c] = (float)sf; mean[c] = (float)mu; inv[c] = (float)r;
    }
}

// ------------------------------------------------------------------------------------------ norm_apply
constexpr int kApplyPix = 4;     // padded pixels per thread (same image, same channel group -> coefficients stay in registers)

__global__ void __launch_bounds__(256) norm_apply_kernel(aclgan_apply_args a) {
    const int u = a.upsample, p = a.dst.pad;
    const int hd = a.y.h * u, wd = a.y.w * u, hp = hd + 2 * p, wp = wd + 2 * p;
    const int cg = a.y.c / 8;
    const int lanes = 256 / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    if (lane >= lanes) return;
    const int64_t npix = (int64_t)hp * wp;
    F8 sc, sf;
    if (a.scale != 0) {
        sc = load_coef8(a.scale, (int64_t)n * a.y.c + g * 8);
        sf = load_coef8(a.shift, (int64_t)n * a.y.c + g * 8);
    }
    const int64_t p0 = (int64_t)blockIdx.x * lanes * kApplyPix + lane;
#pragma unroll
    for (int j = 0; j < kApplyPix; ++j) {
        const int64_t pix = p0 + (int64_t)j * lanes;
        if (pix >= npix) break;
        const int Y = (int)(pix / wp), X = (int)(pix % wp);
        const int y = reflect_idx(Y - p, hd) / u, x = reflect_idx(X - p, wd) / u;
        F8 v = load8(a.y.ptr, a.y.kind, (((int64_t)n * a.y.h + y) * a.y.w + x) * a.y.c + g * 8);
        if (a.scale != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] = v.v[i] * sc.v[i] + sf.v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = act_fn(v.v[i], a.act, a.slope);
        if (a.has_res) {
            const int rp = a.res.pad, rwp = a.res.w + 2 * rp, rhp = a.res.h + 2 * rp;
            const F8 rr = load8_planes(a.res.data, a.res.planes,
                                       (((int64_t)n * rhp + y + rp) * rwp + x + rp) * a.res.c + g * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] += rr.v[i];
        }
        store8_planes(a.dst.data, a.dst.planes, ((int64_t)n * npix + pix) * a.dst.c + g * 8, v);
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
        }
        return;
    }
    for (int c = threadIdx.x; c < a.c; c += blockDim.x) {
        if (c >= c_valid) { ca[c] = cb[c] = cc[c] = 0.f; continue; }
        double g = 1.0;
        if (a.mode == ACLGAN_NORM_ADAIN) {
            g = reinterpret_cast<const float*>(a.w)[(int64_t)n * c_valid + c];
            reinterpret_cast<float*>(a.dw)[(int64_t)n * c_valid + c] = (float)sums[2 * c + 1];
            reinterpret_cast<float*>(a.db)[(int64_t)n * c_valid + c] = (float)sums[2 * c];
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
    const int lanes = 256 / (a->y.c / 8);
    dim3 grid(grid_for(npix, lanes * kApplyPix), a->y.n);
    norm_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_block_bwd_reduce(const aclgan_block_bwd_args* a, void* stream) {
    if (check_cg(a->c)) return ACLGAN_ERR_SHAPE;
    const int lanes = kStatThreads / (a->c / 8);
    const int64_t hw = (int64_t)a->h * a->w;
    dim3 grid(grid_for(hw, lanes * kStatIters), a->n);
    const size_t smem = (size_t)lanes * a->c * 2 * sizeof(float);
    block_bwd_reduce_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_block_bwd_apply(const aclgan_block_bwd_args* a, void* stream) {
    if (check_cg(a->c) || a->dy.c != a->c) return ACLGAN_ERR_SHAPE;
    const int64_t npix = (int64_t)(a->h + 2 * a->dy.pad) * (a->w + 2 * a->dy.pad);
    const int lanes = 256 / (a->c / 8);
    dim3 grid(grid_for(npix, lanes * kApplyPix), a->n);
    block_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_norm_bwd_finalize(const aclgan_norm_bwd_finalize_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->c) return ACLGAN_ERR_SHAPE;
    norm_bwd_finalize_kernel<<<a->n, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
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
