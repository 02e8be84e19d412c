// Row-structured variants of the three plane-sized element-wise passes (norm apply forward, block backward reduce / apply).
//
// The round-1 kernels walk a flattened pixel index (one integer division and a reflect computation per pixel), reload their
// per-channel coefficients from shared memory for every pixel and branch on run-time flags inside the pixel loop:
// `ncu --set full` shows ~135 thread instructions per 8-channel vector at ~50 % issue-slot utilisation, i.e. they are bound by
// instruction issue at 0.35-0.55 of the HBM roofline (profiles/r2_elementwise_ncu.md).  Here a CTA owns whole plane ROWS of one
// image: the row base addresses are computed once per row, a thread keeps ONE 8-channel group for its whole life so the
// (pre-combined) coefficients live in registers, the pixel loop is unrolled with all loads issued first, and every mode flag
// is a template parameter.  Same arithmetic as elementwise.cu (which stays as the generic / fallback path: 2x upsample, tanh,
// degenerate plane sizes).
//
// Replaces (reference): InstanceNorm2d networks.py:333, AdaptiveInstanceNorm2d.forward :490-503, LayerNorm.forward :520-536,
// ReLU / LeakyReLU :345-347, residual add :309, the consumer's ReflectionPad2d :319 - and autograd's backward of them.
#include "common.cuh"
#include "elementwise_rows.cuh"

namespace aclgan {

namespace {

constexpr int kThreads = 256;

struct V8 {
    float v[8];
};

__device__ __forceinline__ V8 ldg8(uint64_t base, int kind, int64_t idx) {
    V8 r;
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

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int PLANES>
__device__ __forceinline__ void st8_planes(const uint64_t (&pl)[2], int64_t idx, const V8& v) {
    uint4 q;
    q.x = pack2(v.v[0], v.v[1]); q.y = pack2(v.v[2], v.v[3]); q.z = pack2(v.v[4], v.v[5]); q.w = pack2(v.v[6], v.v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pl[0]) + idx) = q;
    if (PLANES == 2) {
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = v.v[i] - __bfloat162float(__float2bfloat16_rn(v.v[i]));
        q.x = pack2(r[0], r[1]); q.y = pack2(r[2], r[3]); q.z = pack2(r[4], r[5]); q.w = pack2(r[6], r[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pl[1]) + idx) = q;
    }
}

template <int PLANES>
__device__ __forceinline__ V8 ld8_planes(const uint64_t (&pl)[2], int64_t idx) {
    V8 r = ldg8(pl[0], 0, idx);
    if (PLANES == 2) {
        const V8 l = ldg8(pl[1], 0, idx);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] += l.v[i];
    }
    return r;
}

__device__ __forceinline__ V8 coef8(uint64_t base, int64_t idx) {
    V8 r;
    if (base == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
        return r;
    }
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ int reflect1(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

// mirror image of coordinate c under reflect padding of width p (at most one exists when L > 2p + 1), or -1000
__device__ __forceinline__ int mirror1(int c, int L, int p) {
    if (c >= 1 && c <= p) return -c;
    if (c >= L - 1 - p && c <= L - 2) return 2 * (L - 1) - c;
    return -1000;
}

// ------------------------------------------------------------------------------------------ forward apply
// statistics -> scale / shift of this thread's 8 channels of image n: the arithmetic of norm_finalize_kernel (elementwise.cu),
// done by every CTA for itself so that no separate launch sits between the convolution and this pass; CTA 0 of the image also
// stores the coefficients the backward pass reads (scale, shift, mean, inv, sigma)
constexpr int kFinMaxC = 512;       // channels per plane the fused finalize steps share through shared memory

__device__ __forceinline__ void finalize8(const aclgan_norm_finalize_args& f, int n, int g, bool store, V8& sc, V8& sf) {
    __shared__ double sh[kThreads / 32];
    __shared__ float sh_sc[kFinMaxC], sh_sf[kFinMaxC];
    // only the first C / 8 threads (one per channel group) do the fp64 arithmetic - 256 threads x 8 channels x (2 divisions + 1
    // square root) in fp64 cost every CTA ~2.5 us - and hand the coefficients to the CTA's other pixel lanes through shared memory
    const bool compute = threadIdx.x < (f.c >> 3);
    const int C = f.c, c_valid = f.c_valid;
    double mu_ln = 0.0, r_ln = 0.0;
    if (f.mode == ACLGAN_NORM_LN) {
        const int tot = f.stat_groups > 1 ? f.stat_groups * c_valid : c_valid;
        const double* sg = reinterpret_cast<const double*>(f.sums) + (int64_t)n * (f.stat_groups > 1 ? tot : C) * 2;
        double s1 = 0.0, s2 = 0.0;
        for (int c = threadIdx.x; c < tot; c += kThreads) { s1 += sg[2 * c]; s2 += sg[2 * c + 1]; }
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            double v = pass == 0 ? s1 : s2;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) sh[wp] = v;
            __syncthreads();
            double t = 0.0;
            for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
            if (pass == 0) s1 = t; else s2 = t;
        }
        const double M = (double)c_valid * f.hw;
        mu_ln = s1 / M;
        double var = (s2 - M * mu_ln * mu_ln) / (M - 1.0);
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var);
        r_ln = 1.0 / (sd + (double)f.eps);
        if (store && threadIdx.x == 0) reinterpret_cast<float*>(f.sigma)[n] = (float)sd;
    }
    const double* sums = reinterpret_cast<const double*>(f.sums) + (int64_t)n * C * 2;
    V8 mean, inv;
    if (compute) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        float scv = 0.f, sfv = 0.f, mv = 0.f, iv = 0.f;
        if (c < c_valid) {
            if (f.mode == ACLGAN_NORM_LN) {
                const double s = (double)reinterpret_cast<const float*>(f.w)[c] * r_ln;
                scv = (float)s;
                sfv = (float)((double)reinterpret_cast<const float*>(f.b)[c] - mu_ln * s);
                mv = (float)mu_ln; iv = (float)r_ln;
            } else {
                const double mu = sums[2 * c] / f.hw;
                double var = sums[2 * c + 1] / f.hw - mu * mu;
                if (var < 0.0) var = 0.0;
                const double r = 1.0 / sqrt(var + (double)f.eps);
                double s = r, t = -mu * r;
                if (f.mode == ACLGAN_NORM_ADAIN) {
                    const int64_t ld = f.wb_stride > 0 ? f.wb_stride : c_valid;
                    const double w = reinterpret_cast<const float*>(f.w)[(int64_t)n * ld + c];
                    const double b = reinterpret_cast<const float*>(f.b)[(int64_t)n * ld + c];
                    s = r * w;
                    t = b - mu * s;
                }
                scv = (float)s; sfv = (float)t; mv = (float)mu; iv = (float)r;
            }
        }
        sc.v[i] = scv; sf.v[i] = sfv; mean.v[i] = mv; inv.v[i] = iv;
        sh_sc[c] = scv; sh_sf[c] = sfv;
    }
    }
    __syncthreads();
    if (!compute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc.v[i] = sh_sc[g * 8 + i]; sf.v[i] = sh_sf[g * 8 + i]; }
    }
    if (store && compute) {
        const int64_t o = (int64_t)n * C + g * 8;
        float4* d;
        d = reinterpret_cast<float4*>(reinterpret_cast<float*>(f.scale) + o);
        d[0] = make_float4(sc.v[0], sc.v[1], sc.v[2], sc.v[3]); d[1] = make_float4(sc.v[4], sc.v[5], sc.v[6], sc.v[7]);
        d = reinterpret_cast<float4*>(reinterpret_cast<float*>(f.shift) + o);
        d[0] = make_float4(sf.v[0], sf.v[1], sf.v[2], sf.v[3]); d[1] = make_float4(sf.v[4], sf.v[5], sf.v[6], sf.v[7]);
        d = reinterpret_cast<float4*>(reinterpret_cast<float*>(f.mean) + o);
        d[0] = make_float4(mean.v[0], mean.v[1], mean.v[2], mean.v[3]); d[1] = make_float4(mean.v[4], mean.v[5], mean.v[6], mean.v[7]);
        d = reinterpret_cast<float4*>(reinterpret_cast<float*>(f.inv) + o);
        d[0] = make_float4(inv.v[0], inv.v[1], inv.v[2], inv.v[3]); d[1] = make_float4(inv.v[4], inv.v[5], inv.v[6], inv.v[7]);
    }
}

// ACT: 0 none, 1 relu, 2 lrelu; FIN: the statistics -> scale / shift step fused (finalize8)
template <int KIND, int PLANES, int ACT, int RES, int UNR, int FIN>
__global__ void __launch_bounds__(kThreads) norm_apply_rows_kernel(aclgan_apply_args a, int rows_per_cta, aclgan_norm_finalize_args f) {
    const int p = a.dst.pad, h = a.y.h, w = a.y.w, hp = h + 2 * p, wp = w + 2 * p, C = a.y.c;
    const int cg = C >> 3, lanes = kThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    V8 sc, sf;
    bool affine = a.scale != 0;
    if (FIN) {
        // (before the early exit below: the LayerNorm totals are a block-wide reduction)
        finalize8(f, n, g < cg ? g : 0, blockIdx.x == 0 && lane == 0, sc, sf);
        affine = true;
    }
    if (lane >= lanes) return;
    if (!FIN && affine) {
        sc = coef8(a.scale, (int64_t)n * C + g * 8);
        sf = coef8(a.shift, (int64_t)n * C + g * 8);
    }
    const int rp = a.res.pad, rwp = a.res.w + 2 * rp, rhp = a.res.h + 2 * rp;
    const float slope = a.slope;
    const int row_end = min(hp, ((int)blockIdx.x + 1) * rows_per_cta);
    for (int Y = blockIdx.x * rows_per_cta; Y < row_end; ++Y) {
        const int y = reflect1(Y - p, h);
        const int64_t src_row = ((int64_t)n * h + y) * w * C + g * 8;
        const int64_t res_row = RES ? (((int64_t)n * rhp + y + rp) * rwp + rp) * C + g * 8 : 0;
        const int64_t dst_row = ((int64_t)n * hp + Y) * wp * C + g * 8;
        for (int X0 = lane; X0 < wp; X0 += lanes * UNR) {
            V8 yv[UNR], rv[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int X = min(X0 + u * lanes, wp - 1);           // clamped duplicate: loaded, never stored
                const int x = reflect1(X - p, w);
                yv[u] = ldg8(a.y.ptr, KIND, src_row + (int64_t)x * C);
                if (RES) rv[u] = ld8_planes<PLANES>(a.res.data, res_row + (int64_t)x * C);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int X = X0 + u * lanes;
                if (X >= wp) break;
                V8 v = yv[u];
                if (affine) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = v.v[i] * sc.v[i] + sf.v[i];
                }
                if (ACT == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = fmaxf(v.v[i], 0.f);
                } else if (ACT == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = v.v[i] > 0.f ? v.v[i] : v.v[i] * slope;
                }
                if (RES) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] += rv[u].v[i];
                }
                st8_planes<PLANES>(a.dst.data, dst_row + (int64_t)X * C, v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward: dz of one pixel
// MASK: 0 none, 1 from z = y*scale + shift (norm blocks), 2 from the forward output plane (no-norm blocks)
struct BwdRow {
    int64_t gp_row, gp_mrow;     // element offsets of gradient-plane row (y + p) and of its mirror row (or -1)
    int64_t dense_row;           // offset of row y in dense [n][h][w][C] tensors (gr, y)
    int64_t out_row;             // offset of row (y + po) of the forward output plane, at its column po
};

template <int KIND>
__device__ __forceinline__ void add8(V8& acc, uint64_t base, int64_t idx) {
    const V8 t = ldg8(base, KIND, idx);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
}

// 8 channels exactly as loaded (16 B of bf16 or 32 B of fp32): the loads of a whole batch of pixels stay in flight in this form
// (4 registers per bf16 vector) and are unpacked at the point of use
template <int KIND>
struct Raw8 {
    uint4 q[KIND == 0 ? 1 : 2];
};
template <int KIND>
__device__ __forceinline__ Raw8<KIND> ldg_raw(uint64_t base, int64_t idx) {
    Raw8<KIND> r;
    if (KIND == 0) {
        r.q[0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx);
        r.q[0] = p[0]; r.q[KIND == 0 ? 0 : 1] = p[1];
    }
    return r;
}
template <int KIND>
__device__ __forceinline__ V8 unpack8(const Raw8<KIND>& q) {
    V8 r;
    if (KIND == 0) {
        const uint32_t w[4] = {q.q[0].x, q.q[0].y, q.q[0].z, q.q[0].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r.v[2 * i] = __uint_as_float(w[i] << 16);
            r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    } else {
        const uint4 a = q.q[0], b = q.q[KIND == 0 ? 0 : 1];
        r.v[0] = __uint_as_float(a.x); r.v[1] = __uint_as_float(a.y); r.v[2] = __uint_as_float(a.z); r.v[3] = __uint_as_float(a.w);
        r.v[4] = __uint_as_float(b.x); r.v[5] = __uint_as_float(b.y); r.v[6] = __uint_as_float(b.z); r.v[7] = __uint_as_float(b.w);
    }
    return r;
}

// raw loads of pixel x of the current row (no dependent arithmetic): gradient-plane value, dense gradient, y / forward output
template <int KIND, int MASK, int NORM, int GP, int GR>
struct PixRaw {
    Raw8<KIND> g, gr;
    Raw8<(MASK == 2 && !NORM) ? 0 : KIND> yv;      // (MASK 2 reads the bf16 forward output plane)
};

template <int KIND, int MASK, int NORM, int GP, int GR>
__device__ __forceinline__ void pix_raw(const aclgan_block_bwd_args& a, const BwdRow& r, int x, int C, PixRaw<KIND, MASK, NORM, GP, GR>& o) {
    if (GP) o.g = ldg_raw<KIND>(a.gp, r.gp_row + (int64_t)(x + a.gp_pad) * C);
    if (GR) o.gr = ldg_raw<KIND>(a.gr, r.dense_row + (int64_t)x * C);
    if constexpr (MASK == 1 || NORM) o.yv = ldg_raw<KIND>(a.y.ptr, r.dense_row + (int64_t)x * C);
    else if constexpr (MASK == 2) o.yv = ldg_raw<0>(a.out.data[0], r.out_row + (int64_t)x * C);
}

// dz = (fold of the padded-plane gradient + dense gradient) * act'(z)
template <int KIND, int MASK, int NORM, int GP, int GR>
__device__ __forceinline__ V8 pix_dz(const aclgan_block_bwd_args& a, const BwdRow& r, int x, int C, const PixRaw<KIND, MASK, NORM, GP, GR>& in,
                                     const V8& yv, const V8& scale, const V8& shift, float slope) {
    V8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    if (GP) {
        acc = unpack8<KIND>(in.g);
        const int p = a.gp_pad;
        if (p > 0) {        // reflect images of this pixel in the padded plane (border-adjacent pixels only)
            const int mx = mirror1(x, a.w, p);
            if (mx > -1000) add8<KIND>(acc, a.gp, r.gp_row + (int64_t)(mx + p) * C);
            if (r.gp_mrow >= 0) {
                add8<KIND>(acc, a.gp, r.gp_mrow + (int64_t)(x + p) * C);
                if (mx > -1000) add8<KIND>(acc, a.gp, r.gp_mrow + (int64_t)(mx + p) * C);
            }
        }
    }
    if (GR) {
        const V8 gr = unpack8<KIND>(in.gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += gr.v[i];
    }
    if (MASK == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = yv.v[i] * scale.v[i] + shift.v[i];
            if (!(z > 0.f)) acc.v[i] *= slope;
        }
    } else if (MASK == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (!(yv.v[i] > 0.f)) acc.v[i] *= slope;
    }
    return acc;
}

__device__ __forceinline__ BwdRow make_row(const aclgan_block_bwd_args& a, int n, int y, int g) {
    BwdRow r;
    const int p = a.gp_pad, C = a.c;
    const int64_t wpp = a.w + 2 * p, hpp = a.h + 2 * p;
    r.gp_row = ((int64_t)n * hpp + y + p) * wpp * C + g * 8;
    const int my = p > 0 ? mirror1(y, a.h, p) : -1000;
    r.gp_mrow = my > -1000 ? ((int64_t)n * hpp + my + p) * wpp * C + g * 8 : -1;
    r.dense_row = ((int64_t)n * a.h + y) * a.w * C + g * 8;
    const int po = a.out.pad;
    r.out_row = (((int64_t)n * (a.out.h + 2 * po) + y + po) * (a.out.w + 2 * po) + po) * a.out.c + g * 8;
    return r;
}

// per-thread channel partials -> CTA -> atomics (fp64 statistics and / or fp32 bias gradient)
__device__ __forceinline__ void cta_reduce16(float* red, const V8& s, const V8& q, int C, int g, int lane, int lanes, bool active,
                                             double* stat, float* dbias, int dbias_n) {
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[((lane * C) + g * 8 + i) * 2] = s.v[i];
            red[((lane * C) + g * 8 + i) * 2 + 1] = q.v[i];
        }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float x = 0.f, y = 0.f;
        for (int l = 0; l < lanes; ++l) { x += red[(l * C + ch) * 2]; y += red[(l * C + ch) * 2 + 1]; }
        if (stat != nullptr) {
            atomicAdd(&stat[ch * 2], (double)x);
            atomicAdd(&stat[ch * 2 + 1], (double)y);
        }
        if (dbias != nullptr && ch < dbias_n) atomicAdd(dbias + ch, x);
    }
}

// ------------------------------------------------------------------------------------------ backward reduce
template <int KIND, int MASK, int GP, int GR, int UNR>
__global__ void __launch_bounds__(kThreads) bwd_reduce_rows_kernel(aclgan_block_bwd_args a, int rows_per_cta) {
    extern __shared__ float red[];          // [lanes][C][2]
    const int C = a.c, cg = C >> 3, lanes = kThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const bool active = lane < lanes;
    V8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = q.v[i] = 0.f;
    if (active) {
        const int64_t ci = (int64_t)n * C + g * 8;
        const V8 scale = MASK == 1 ? coef8(a.scale, ci) : V8(), shift = MASK == 1 ? coef8(a.shift, ci) : V8();
        V8 inv = coef8(a.inv, ci), nmi = coef8(a.mean, ci);
#pragma unroll
        for (int i = 0; i < 8; ++i) nmi.v[i] = -nmi.v[i] * inv.v[i];          // yhat = y * inv - mean * inv
        const float slope = a.slope;
        const int row_end = min(a.h, ((int)blockIdx.x + 1) * rows_per_cta);
        for (int y = blockIdx.x * rows_per_cta; y < row_end; ++y) {
            const BwdRow r = make_row(a, n, y, g);
            for (int x0 = lane; x0 < a.w; x0 += lanes * UNR) {
                PixRaw<KIND, MASK, 1, GP, GR> in[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) pix_raw<KIND, MASK, 1, GP, GR>(a, r, min(x0 + u * lanes, a.w - 1), C, in[u]);
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int x = x0 + u * lanes;
                    if (x >= a.w) break;
                    const V8 yv = unpack8<KIND>(in[u].yv);
                    const V8 dz = pix_dz<KIND, MASK, 1, GP, GR>(a, r, x, C, in[u], yv, scale, shift, slope);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        s.v[i] += dz.v[i];
                        q.v[i] += dz.v[i] * (yv.v[i] * inv.v[i] + nmi.v[i]);
                    }
                }
            }
        }
    }
    cta_reduce16(red, s, q, C, g, lane, lanes, active, reinterpret_cast<double*>(a.sums) + (int64_t)n * C * 2, nullptr, 0);
}

// T1 / T2 -> (ca, cb, cc) of this thread's 8 channels of image n: the arithmetic of norm_bwd_finalize_kernel (elementwise.cu),
// done by every CTA for itself (no separate launch between the reduce and the apply pass); `store` (CTA 0 of the image, one
// thread per channel group) also performs the kernel's side effects: the norm layer's parameter gradients and, for
// LayerNorm, the gradient of the conv bias in front of it
__device__ __forceinline__ void bwd_finalize8(const aclgan_norm_bwd_finalize_args& f, int n, int g, bool store, V8& ca, V8& cb, V8& cc) {
    __shared__ double sh[kThreads / 32];
    __shared__ float sh_a[kFinMaxC], sh_b[kFinMaxC], sh_c[kFinMaxC];
    const bool compute = threadIdx.x < (f.c >> 3);            // (see finalize8)
    const int C = f.c, c_valid = f.c_valid;
    const double* sums = reinterpret_cast<const double*>(f.sums) + (int64_t)n * C * 2;
    const float* inv = reinterpret_cast<const float*>(f.inv) + (int64_t)n * C;
    double s1 = 0.0, s2 = 0.0, sd = 0.0;
    if (f.mode == ACLGAN_NORM_LN) {
        const float* gam = reinterpret_cast<const float*>(f.w);
        for (int c = threadIdx.x; c < c_valid; c += kThreads) {
            s1 += (double)gam[c] * sums[2 * c];
            s2 += (double)gam[c] * sums[2 * c + 1];
        }
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            double v = pass == 0 ? s1 : s2;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) sh[wp] = v;
            __syncthreads();
            double t = 0.0;
            for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
            if (pass == 0) s1 = t; else s2 = t;
        }
        sd = reinterpret_cast<const float*>(f.sigma)[n];
    }
    const double M = (double)c_valid * f.hw;
    if (compute) {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        float av = 0.f, bv = 0.f, cv = 0.f;
        if (c < c_valid) {
            const double r = inv[c];
            if (f.mode == ACLGAN_NORM_LN) {
                const float* gam = reinterpret_cast<const float*>(f.w);
                av = (float)(r * gam[c]);
                bv = (float)(sd > 0.0 ? -s2 / ((M - 1.0) * sd) : 0.0);
                cv = (float)(-r * s1 / M);
                if (store) {
                    atomicAdd(reinterpret_cast<float*>(f.dw) + c, (float)sums[2 * c + 1]);
                    atomicAdd(reinterpret_cast<float*>(f.db) + c, (float)sums[2 * c]);
                    if (f.dbias != 0) {
                        const double mu = reinterpret_cast<const float*>(f.mean)[(int64_t)n * C + c];
                        double s1f;
                        if (f.fstat_groups > 1) {
                            s1f = 0.0;
                            for (int gph = 0; gph < f.fstat_groups; ++gph)
                                s1f += reinterpret_cast<const double*>(f.fsums)[(((int64_t)n * f.fstat_groups + gph) * c_valid + c) * 2];
                        } else {
                            s1f = reinterpret_cast<const double*>(f.fsums)[((int64_t)n * C + c) * 2];
                        }
                        const double syh = (s1f - (double)f.hw * mu) * r;
                        const double dbv = (double)av * sums[2 * c] + (double)bv * syh + (double)cv * (double)f.hw;
                        atomicAdd(reinterpret_cast<float*>(f.dbias) + c, (float)dbv);
                    }
                }
            } else {
                double gg = 1.0;
                if (f.mode == ACLGAN_NORM_ADAIN) {
                    const int64_t ld = f.wb_stride > 0 ? f.wb_stride : c_valid;
                    gg = reinterpret_cast<const float*>(f.w)[(int64_t)n * ld + c];
                    if (store) {
                        reinterpret_cast<float*>(f.dw)[(int64_t)n * ld + c] = (float)sums[2 * c + 1];
                        reinterpret_cast<float*>(f.db)[(int64_t)n * ld + c] = (float)sums[2 * c];
                    }
                }
                av = (float)(r * gg);
                bv = (float)(-r * gg * sums[2 * c + 1] / f.hw);
                cv = (float)(-r * gg * sums[2 * c] / f.hw);
            }
        }
        ca.v[i] = av; cb.v[i] = bv; cc.v[i] = cv;
        sh_a[c] = av; sh_b[c] = bv; sh_c[c] = cv;
    }
    }
    __syncthreads();
    if (!compute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { ca.v[i] = sh_a[g * 8 + i]; cb.v[i] = sh_b[g * 8 + i]; cc.v[i] = sh_c[g * 8 + i]; }
    }
}

// ------------------------------------------------------------------------------------------ backward apply
// dy = ca*dz + cb*yhat + cc  (= ca*dz + B*y + Cc with B = cb*inv, Cc = cc - B*mean), written as the zero-bordered plane(s)
template <int KIND, int PLANES, int MASK, int NORM, int GP, int GR, int DBIAS, int UNR>
__global__ void __launch_bounds__(kThreads) bwd_apply_rows_kernel(aclgan_block_bwd_args a, int rows_per_cta, int fin,
                                                                  aclgan_norm_bwd_finalize_args f) {
    extern __shared__ float red[];
    const int C = a.c, cg = C >> 3, lanes = kThreads / cg;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
    const int n = blockIdx.y;
    const bool active = lane < lanes;
    const int pz = a.dy.pad, hz = a.h + 2 * pz, wz = a.w + 2 * pz;
    V8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = q.v[i] = 0.f;
    V8 ca, cb, cc;
    if (NORM && fin) bwd_finalize8(f, n, g, blockIdx.x == 0 && lane == 0, ca, cb, cc);      // (block-wide for LayerNorm: all threads)
    if (active) {
        const int64_t ci = (int64_t)n * C + g * 8;
        const V8 scale = MASK == 1 ? coef8(a.scale, ci) : V8(), shift = MASK == 1 ? coef8(a.shift, ci) : V8();
        if (NORM) {
            if (!fin) {
                ca = coef8(a.ca, ci);
                cb = coef8(a.cb, ci);
                cc = coef8(a.cc, ci);
            }
            const V8 inv = coef8(a.inv, ci), mean = coef8(a.mean, ci);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cb.v[i] *= inv.v[i];                    // B
                cc.v[i] -= cb.v[i] * mean.v[i];         // Cc
            }
        }
        const float slope = a.slope;
        const int row_end = min(hz, ((int)blockIdx.x + 1) * rows_per_cta);
        V8 zero;
#pragma unroll
        for (int i = 0; i < 8; ++i) zero.v[i] = 0.f;
        for (int Y = blockIdx.x * rows_per_cta; Y < row_end; ++Y) {
            const int y = Y - pz;
            const int64_t dst_row = ((int64_t)n * hz + Y) * wz * C + g * 8;
            if (y < 0 || y >= a.h) {
                for (int X = lane; X < wz; X += lanes) st8_planes<PLANES>(a.dy.data, dst_row + (int64_t)X * C, zero);
                continue;
            }
            const BwdRow r = make_row(a, n, y, g);
            for (int X0 = lane; X0 < wz; X0 += lanes * UNR) {
                PixRaw<KIND, MASK, NORM, GP, GR> in[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) pix_raw<KIND, MASK, NORM, GP, GR>(a, r, min(max(X0 + u * lanes - pz, 0), a.w - 1), C, in[u]);
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int X = X0 + u * lanes;
                    if (X >= wz) break;
                    const int x = X - pz;
                    if (x < 0 || x >= a.w) {
                        st8_planes<PLANES>(a.dy.data, dst_row + (int64_t)X * C, zero);
                        continue;
                    }
                    V8 yv;
                    if (MASK != 0 || NORM) yv = unpack8<(MASK == 2 && !NORM) ? 0 : KIND>(in[u].yv);
                    const V8 dz = pix_dz<KIND, MASK, NORM, GP, GR>(a, r, x, C, in[u], yv, scale, shift, slope);
                    V8 out;
                    if (NORM) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) out.v[i] = ca.v[i] * dz.v[i] + (cb.v[i] * yv.v[i] + cc.v[i]);
                    } else {
                        out = dz;
                    }
                    if (DBIAS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) s.v[i] += dz.v[i];
                    }
                    st8_planes<PLANES>(a.dy.data, dst_row + (int64_t)X * C, out);
                }
            }
        }
    }
    if (DBIAS) cta_reduce16(red, s, q, C, g, lane, lanes, active, nullptr, reinterpret_cast<float*>(a.dbias), a.dbias_n);
}

static int rows_per_cta_for(int rows, int n_images) {
    // about two CTAs per SM over all images = ONE wave at the occupancy of the backward kernels (4 pixels per thread in flight,
    // ~120 registers): measured on the 256 x 256 batch-8 step-pair, env ACLGAN_ROWS_PER_SM = 12 / 6 / 3 / 2: 31.40 / 31.28 / 31.12 /
    // 30.99 ms - a second, partially filled wave of one-row CTAs costs more than the longer per-CTA row loop
    static int per_sm = 0;
    if (per_sm == 0) {
        const char* e = getenv("ACLGAN_ROWS_PER_SM");
        per_sm = e != nullptr && atoi(e) > 0 ? atoi(e) : 2;
    }
    const int want = (per_sm * num_sms() + n_images - 1) / n_images;
    int r = (rows + want - 1) / want;
    return r < 1 ? 1 : r;
}

static bool rows_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ACLGAN_ELEMENTWISE_ROWS");
        v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }
    return v == 1;
}

template <int KIND, int PLANES, int ACT>
static void launch_apply_res(const aclgan_apply_args* a, const aclgan_norm_finalize_args* f, dim3 grid, int rpc, cudaStream_t st) {
    if (f != nullptr) {
        if (a->has_res) norm_apply_rows_kernel<KIND, PLANES, ACT, 1, 4, 1><<<grid, kThreads, 0, st>>>(*a, rpc, *f);
        else norm_apply_rows_kernel<KIND, PLANES, ACT, 0, 4, 1><<<grid, kThreads, 0, st>>>(*a, rpc, *f);
        return;
    }
    aclgan_norm_finalize_args none;
    memset(&none, 0, sizeof(none));
    if (a->has_res) norm_apply_rows_kernel<KIND, PLANES, ACT, 1, 4, 0><<<grid, kThreads, 0, st>>>(*a, rpc, none);
    else norm_apply_rows_kernel<KIND, PLANES, ACT, 0, 4, 0><<<grid, kThreads, 0, st>>>(*a, rpc, none);
}
template <int KIND, int PLANES>
static void launch_apply_act(const aclgan_apply_args* a, const aclgan_norm_finalize_args* f, dim3 grid, int rpc, cudaStream_t st) {
    if (a->act == ACLGAN_ACT_RELU) launch_apply_res<KIND, PLANES, 1>(a, f, grid, rpc, st);
    else if (a->act == ACLGAN_ACT_LRELU) launch_apply_res<KIND, PLANES, 2>(a, f, grid, rpc, st);
    else launch_apply_res<KIND, PLANES, 0>(a, f, grid, rpc, st);
}

// pixels per thread and loop iteration of the backward kernels (env ACLGAN_BWD_UNR = 2 | 4): a CTA owns one or two plane rows, so
// the kernels are bound by the latency of their few dependent load -> compute -> store rounds, not by bandwidth
static int bwd_unr() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("ACLGAN_BWD_UNR");
        v = (e != nullptr && atoi(e) == 2) ? 2 : 4;
    }
    return v;
}

template <int KIND, int MASK>
static void launch_reduce_g(const aclgan_block_bwd_args* a, dim3 grid, int rpc, size_t smem, cudaStream_t st) {
    if constexpr (KIND == 0) if (bwd_unr() == 4) {
        if (a->gp != 0 && a->gr != 0) bwd_reduce_rows_kernel<KIND, MASK, 1, 1, 4><<<grid, kThreads, smem, st>>>(*a, rpc);
        else if (a->gp != 0) bwd_reduce_rows_kernel<KIND, MASK, 1, 0, 4><<<grid, kThreads, smem, st>>>(*a, rpc);
        else bwd_reduce_rows_kernel<KIND, MASK, 0, 1, 4><<<grid, kThreads, smem, st>>>(*a, rpc);
        return;
    }
    if (a->gp != 0 && a->gr != 0) bwd_reduce_rows_kernel<KIND, MASK, 1, 1, 2><<<grid, kThreads, smem, st>>>(*a, rpc);
    else if (a->gp != 0) bwd_reduce_rows_kernel<KIND, MASK, 1, 0, 2><<<grid, kThreads, smem, st>>>(*a, rpc);
    else bwd_reduce_rows_kernel<KIND, MASK, 0, 1, 2><<<grid, kThreads, smem, st>>>(*a, rpc);
}

template <int KIND, int PLANES, int MASK, int NORM, int GP, int GR>
static void launch_apply_bwd_db(const aclgan_block_bwd_args* a, const aclgan_norm_bwd_finalize_args* fp, dim3 grid, int rpc, size_t smem,
                                cudaStream_t st) {
    aclgan_norm_bwd_finalize_args f;
    if (fp != nullptr) f = *fp; else memset(&f, 0, sizeof(f));
    const int fin = fp != nullptr ? 1 : 0;
    if constexpr (KIND == 0 && PLANES == 1) if (bwd_unr() == 4) {
        if (a->dbias != 0) bwd_apply_rows_kernel<KIND, PLANES, MASK, NORM, GP, GR, 1, 4><<<grid, kThreads, smem, st>>>(*a, rpc, fin, f);
        else bwd_apply_rows_kernel<KIND, PLANES, MASK, NORM, GP, GR, 0, 4><<<grid, kThreads, 0, st>>>(*a, rpc, fin, f);
        return;
    }
    if (a->dbias != 0) bwd_apply_rows_kernel<KIND, PLANES, MASK, NORM, GP, GR, 1, 2><<<grid, kThreads, smem, st>>>(*a, rpc, fin, f);
    else bwd_apply_rows_kernel<KIND, PLANES, MASK, NORM, GP, GR, 0, 2><<<grid, kThreads, 0, st>>>(*a, rpc, fin, f);
}
template <int KIND, int PLANES, int MASK, int NORM>
static void launch_apply_bwd_g(const aclgan_block_bwd_args* a, const aclgan_norm_bwd_finalize_args* f, dim3 grid, int rpc, size_t smem,
                               cudaStream_t st) {
    if (a->gp != 0 && a->gr != 0) launch_apply_bwd_db<KIND, PLANES, MASK, NORM, 1, 1>(a, f, grid, rpc, smem, st);
    else if (a->gp != 0) launch_apply_bwd_db<KIND, PLANES, MASK, NORM, 1, 0>(a, f, grid, rpc, smem, st);
    else launch_apply_bwd_db<KIND, PLANES, MASK, NORM, 0, 1>(a, f, grid, rpc, smem, st);
}
template <int KIND, int PLANES>
static void launch_apply_bwd_mode(const aclgan_block_bwd_args* a, const aclgan_norm_bwd_finalize_args* f, dim3 grid, int rpc, size_t smem,
                                  cudaStream_t st) {
    if (a->norm) {
        if (a->mask_mode == ACLGAN_MASK_FROM_Z) launch_apply_bwd_g<KIND, PLANES, 1, 1>(a, f, grid, rpc, smem, st);
        else launch_apply_bwd_g<KIND, PLANES, 0, 1>(a, f, grid, rpc, smem, st);
    } else {
        if (a->mask_mode == ACLGAN_MASK_FROM_OUT) launch_apply_bwd_g<KIND, PLANES, 2, 0>(a, nullptr, grid, rpc, smem, st);
        else launch_apply_bwd_g<KIND, PLANES, 0, 0>(a, nullptr, grid, rpc, smem, st);
    }
}

}  // namespace

int rows_norm_apply(const aclgan_apply_args* a, cudaStream_t st) { return rows_norm_finalize_apply(nullptr, a, st); }

// f != nullptr: the statistics -> coefficient step of aclgan_norm_finalize fused into the apply launch
int rows_norm_finalize_apply(const aclgan_norm_finalize_args* f, const aclgan_apply_args* a, cudaStream_t st) {
    if (f != nullptr) {
        // (callers opt in by calling the fused entry point: engine.py, ACLGAN_FUSE_FINALIZE=1)
        if (f->n != a->y.n || f->c != a->y.c || f->c % 8 != 0 || f->c > kFinMaxC || f->mode < ACLGAN_NORM_IN || f->mode > ACLGAN_NORM_LN) return -100;
    }
    if (!rows_enabled() || a->upsample != 1 || a->act == ACLGAN_ACT_TANH) return -100;
    const int C = a->y.c, p = a->dst.pad;
    if (C % 8 != 0 || C / 8 > kThreads || a->y.h <= p || a->y.w <= p || (a->y.kind != 0 && a->y.kind != 1)) return -100;
    if (a->dst.planes != 1 && a->dst.planes != 2) return -100;
    if (a->has_res && a->res.planes != a->dst.planes) return -100;
    const int hp = a->y.h + 2 * p;
    const int rpc = rows_per_cta_for(hp, a->y.n);
    dim3 grid((hp + rpc - 1) / rpc, a->y.n);
    if (a->y.kind == 0) {
        if (a->dst.planes == 1) launch_apply_act<0, 1>(a, f, grid, rpc, st); else launch_apply_act<0, 2>(a, f, grid, rpc, st);
    } else {
        if (a->dst.planes == 1) launch_apply_act<1, 1>(a, f, grid, rpc, st); else launch_apply_act<1, 2>(a, f, grid, rpc, st);
    }
    return (int)cudaGetLastError();
}

static bool bwd_rows_ok(const aclgan_block_bwd_args* a) {
    if (!rows_enabled() || a->upsample != 1) return false;
    const int C = a->c, p = a->gp_pad;
    if (C % 8 != 0 || C / 8 > kThreads) return false;
    if (a->gp == 0 && a->gr == 0) return false;
    if ((a->mask_mode == ACLGAN_MASK_FROM_Z || a->norm) && a->y.kind != a->g_kind) return false;
    if (a->gp != 0 && (a->h <= 2 * p + 1 || a->w <= 2 * p + 1)) return false;
    if (a->mask_mode == ACLGAN_MASK_FROM_Z && !a->norm) return false;
    if (a->mask_mode == ACLGAN_MASK_FROM_OUT && a->norm) return false;
    return a->g_kind == 0 || a->g_kind == 1;
}

int rows_bwd_reduce(const aclgan_block_bwd_args* a, cudaStream_t st) {
    if (!bwd_rows_ok(a) || !a->norm) return -100;
    const int C = a->c, lanes = kThreads / (C / 8);
    const int rpc = rows_per_cta_for(a->h, a->n);
    dim3 grid((a->h + rpc - 1) / rpc, a->n);
    const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
    const bool z = a->mask_mode == ACLGAN_MASK_FROM_Z;
    if (a->g_kind == 0) {
        if (z) launch_reduce_g<0, 1>(a, grid, rpc, smem, st); else launch_reduce_g<0, 0>(a, grid, rpc, smem, st);
    } else {
        if (z) launch_reduce_g<1, 1>(a, grid, rpc, smem, st); else launch_reduce_g<1, 0>(a, grid, rpc, smem, st);
    }
    return (int)cudaGetLastError();
}

int rows_bwd_apply(const aclgan_block_bwd_args* a, cudaStream_t st) { return rows_bwd_finalize_apply(nullptr, a, st); }

// f != nullptr: the T1 / T2 -> coefficient step of aclgan_norm_bwd_finalize (and its parameter-gradient side effects) fused in
int rows_bwd_finalize_apply(const aclgan_norm_bwd_finalize_args* f, const aclgan_block_bwd_args* a, cudaStream_t st) {
    if (f != nullptr) {
        // (callers opt in by calling the fused entry point: engine.py, ACLGAN_FUSE_FINALIZE=1)
        if (!a->norm || f->n != a->n || f->c != a->c || f->c > kFinMaxC || f->mode < ACLGAN_NORM_IN || f->mode > ACLGAN_NORM_LN) return -100;
    }
    if (!bwd_rows_ok(a)) return -100;
    if (a->dy.planes != 1 && a->dy.planes != 2) return -100;
    const int C = a->c, lanes = kThreads / (C / 8);
    const int hz = a->h + 2 * a->dy.pad;
    const int rpc = rows_per_cta_for(hz, a->n);
    dim3 grid((hz + rpc - 1) / rpc, a->n);
    const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
    if (a->g_kind == 0) {
        if (a->dy.planes == 1) launch_apply_bwd_mode<0, 1>(a, f, grid, rpc, smem, st); else launch_apply_bwd_mode<0, 2>(a, f, grid, rpc, smem, st);
    } else {
        if (a->dy.planes == 1) launch_apply_bwd_mode<1, 1>(a, f, grid, rpc, smem, st); else launch_apply_bwd_mode<1, 2>(a, f, grid, rpc, smem, st);
    }
    return (int)cudaGetLastError();
}

}  // namespace aclgan
