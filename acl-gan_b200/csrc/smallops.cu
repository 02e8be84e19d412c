// Small operators of the ACL-GAN step that are neither convolutions nor plane-sized norm passes:
// AvgPool 3/2/1 with valid-count divisor (image pyramid of the multi-scale discriminator), the style head (global average
// pool + 1x1 conv), the style MLP, the discriminator head (1x1 conv to one logit) fused with the LSGAN terms and their
// gradient seed, the focus blend and its adjoint, the L1 identity loss, the focus size / digit losses, the final linear
// combination of all loss accumulators into the loss_* scalars, and fp32 / bf16 accumulate + scale helpers.
// All reductions: per-thread partials -> warp shuffles -> one atomic per warp / CTA; loss accumulators are fp64.
//
// Replaces (reference file:line): networks.py:33,53 (AvgPool2d), :222-223 (AdaptiveAvgPool2d + Conv2d 1x1), :280-292 (MLP),
// :45 (Conv2d 1x1 head), :60-106 (calc_dis_loss / calc_gen_loss / calc_gen_d2_loss), trainer.py:85-88 (focus_translation),
// :61-62 (recon_criterion), :146-161 (focus losses), :142-165 / :288-290 (weighted loss totals).
#include "common.cuh"

namespace aclgan {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sum of one double per thread; valid in thread 0
__device__ __forceinline__ double cta_sum_d(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t += sh[i];
    return t;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// value of channel c at padded-plane element index `idx` (hi + lo planes)
__device__ __forceinline__ float plane_val(const aclgan_act& a, int64_t idx) {
    float v = bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(a.data[0])[idx]);
    if (a.planes == 2) v += bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(a.data[1])[idx]);
    return v;
}

__device__ __forceinline__ void store_grad(uint64_t base, int kind, int64_t idx, float v) {
    if (kind == 0) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(base)[idx] = v;
}

// ------------------------------------------------------------------------------------------ K10 avg pool
// window geometry exactly as the pooling definition: rows [2*ph - 1, 2*ph + 2) clipped to the image, divisor = number of
// valid elements; the window is summed in row-major order in fp32 and divided once (bit-identical to ATen's kernel)
__global__ void avgpool_fwd_kernel(aclgan_avgpool_args a, int ho, int wo) {
    const int64_t total = (int64_t)a.planes * ho * wo;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int pw = (int)(t % wo), ph = (int)((t / wo) % ho);
    const int64_t pl = t / ((int64_t)wo * ho);
    int hs = ph * 2 - 1, ws = pw * 2 - 1;
    int he = min(hs + 3, a.h + 1), we = min(ws + 3, a.w + 1);
    hs = max(hs, 0); ws = max(ws, 0); he = min(he, a.h); we = min(we, a.w);
    const float* s = reinterpret_cast<const float*>(a.src) + pl * a.h * a.w;
    float acc = 0.f;
    for (int y = hs; y < he; ++y)
        for (int x = ws; x < we; ++x) acc += __ldg(s + (int64_t)y * a.w + x);
    const int cnt = (he - hs) * (we - ws);
    reinterpret_cast<float*>(a.dst)[t] = cnt > 0 ? __fdiv_rn(acc, (float)cnt) : 0.f;
}

__global__ void avgpool_bwd_kernel(aclgan_avgpool_args a, int ho, int wo) {
    const int64_t total = (int64_t)a.planes * a.h * a.w;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int x = (int)(t % a.w), y = (int)((t / a.w) % a.h);
    const int64_t pl = t / ((int64_t)a.w * a.h);
    const int hq = y + 1, wq = x + 1;                       // coordinates in the padded frame
    const int ph0 = hq < 3 ? 0 : (hq - 3) / 2 + 1, ph1 = min(hq / 2 + 1, ho);
    const int pw0 = wq < 3 ? 0 : (wq - 3) / 2 + 1, pw1 = min(wq / 2 + 1, wo);
    const float* g = reinterpret_cast<const float*>(a.src) + pl * ho * wo;
    float acc = 0.f;
    for (int ph = ph0; ph < ph1; ++ph)
        for (int pw = pw0; pw < pw1; ++pw) {
            int hs = ph * 2 - 1, ws = pw * 2 - 1;
            int he = min(hs + 3, a.h + 1), we = min(ws + 3, a.w + 1);
            hs = max(hs, 0); ws = max(ws, 0); he = min(he, a.h); we = min(we, a.w);
            acc += __fdiv_rn(__ldg(g + (int64_t)ph * wo + pw), (float)((he - hs) * (we - ws)));
        }
    float* d = reinterpret_cast<float*>(a.dst);
    d[t] = a.accumulate ? d[t] + acc : acc;
}

// ------------------------------------------------------------------------------------------ K11 style head
// one CTA per sample: threads stride over channels for the global average (coalesced over the NHWC plane), then one warp
// per style component reduces pooled . W[j]
__global__ void __launch_bounds__(256) style_head_fwd_kernel(aclgan_style_head_args a) {
    extern __shared__ float sh_pool[];           // [c_valid]
    const int n = blockIdx.x;
    const int hp = a.x.h + 2 * a.x.pad, wp = a.x.w + 2 * a.x.pad;
    const float inv_hw = 1.f / (float)(a.x.h * a.x.w);
    for (int c = threadIdx.x; c < a.c_valid; c += blockDim.x) {
        float s = 0.f;
        for (int y = 0; y < a.x.h; ++y)
            for (int x = 0; x < a.x.w; ++x)
                s += plane_val(a.x, (((int64_t)n * hp + y + a.x.pad) * wp + x + a.x.pad) * a.x.c + c);
        s *= inv_hw;
        sh_pool[c] = s;
        reinterpret_cast<float*>(a.pooled)[(int64_t)n * a.c_valid + c] = s;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* W = reinterpret_cast<const float*>(a.weight);
    for (int j = warp; j < a.style_dim; j += blockDim.x >> 5) {
        float s = 0.f;
        for (int c = lane; c < a.c_valid; c += 32) s += sh_pool[c] * __ldg(W + (int64_t)j * a.c_valid + c);
        s = warp_sum(s);
        if (lane == 0) reinterpret_cast<float*>(a.style)[(int64_t)n * a.style_dim + j] = s + __ldg(reinterpret_cast<const float*>(a.bias) + j);
    }
}

// grid (n + 1): CTAs 0..n-1 write the plane gradient of their sample, CTA n accumulates the parameter gradients
__global__ void __launch_bounds__(256) style_head_bwd_kernel(aclgan_style_head_args a) {
    const float* ds = reinterpret_cast<const float*>(a.dstyle);
    const float* W = reinterpret_cast<const float*>(a.weight);
    if ((int)blockIdx.x == a.x.n) {
        if (a.dweight == 0) return;
        const float* pooled = reinterpret_cast<const float*>(a.pooled);
        for (int i = threadIdx.x; i < a.style_dim * a.c_valid; i += blockDim.x) {
            const int j = i / a.c_valid, c = i - j * a.c_valid;
            float s = 0.f;
            for (int n = 0; n < a.x.n; ++n) s += ds[(int64_t)n * a.style_dim + j] * pooled[(int64_t)n * a.c_valid + c];
            reinterpret_cast<float*>(a.dweight)[i] += s;
        }
        for (int j = threadIdx.x; j < a.style_dim; j += blockDim.x) {
            float s = 0.f;
            for (int n = 0; n < a.x.n; ++n) s += ds[(int64_t)n * a.style_dim + j];
            reinterpret_cast<float*>(a.dbias)[j] += s;
        }
        return;
    }
    const int n = blockIdx.x;
    const float inv_hw = 1.f / (float)(a.x.h * a.x.w);
    const int hw = a.x.h * a.x.w;
    for (int c = threadIdx.x; c < a.x.c; c += blockDim.x) {
        float g = 0.f;
        if (c < a.c_valid) {
            for (int j = 0; j < a.style_dim; ++j) g += ds[(int64_t)n * a.style_dim + j] * __ldg(W + (int64_t)j * a.c_valid + c);
            g *= inv_hw;
        }
        for (int p = 0; p < hw; ++p) store_grad(a.gr, a.g_kind, ((int64_t)n * hw + p) * a.x.c + c, g);
    }
}

// ------------------------------------------------------------------------------------------ K12 MLP (linear layers)
// out[s][j] = b[j] + sum_k x[s][k] W[j][k] (+ ReLU): one warp per output feature j, lanes over k, samples in register chunks
constexpr int kLinChunk = 8;

__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ W,
                                                         const float* __restrict__ b, float* __restrict__ out, int n, int K, int J,
                                                         int relu) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= J) return;
    const float* wr = W + (int64_t)warp * K;
    for (int s0 = 0; s0 < n; s0 += kLinChunk) {
        float acc[kLinChunk];
#pragma unroll
        for (int i = 0; i < kLinChunk; ++i) acc[i] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float wv = __ldg(wr + k);
#pragma unroll
            for (int i = 0; i < kLinChunk; ++i)
                if (s0 + i < n) acc[i] += wv * __ldg(x + (int64_t)(s0 + i) * x_stride + k);
        }
#pragma unroll
        for (int i = 0; i < kLinChunk; ++i) {
            const float v = warp_sum(acc[i]) + __ldg(b + warp);
            if (lane == 0 && s0 + i < n) out[(int64_t)(s0 + i) * J + warp] = relu ? fmaxf(v, 0.f) : v;
        }
    }
}

// dz = dout * (out > 0 if relu); dW[j][k] += sum_s dz[s][j] x[s][k]; db[j] += sum_s dz[s][j]   (one warp per j)
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ out,
                                                           const float* __restrict__ dout, float* __restrict__ dW,
                                                           float* __restrict__ db, int n, int K, int J, int relu) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= J) return;
    float bsum = 0.f;
    for (int s = 0; s < n; ++s) {
        float dz = dout[(int64_t)s * J + warp];
        if (relu && !(out[(int64_t)s * J + warp] > 0.f)) dz = 0.f;
        bsum += dz;
    }
    if (lane == 0) db[warp] += bsum;
    for (int k = lane; k < K; k += 32) {
        float acc = 0.f;
        for (int s = 0; s < n; ++s) {
            float dz = dout[(int64_t)s * J + warp];
            if (relu && !(out[(int64_t)s * J + warp] > 0.f)) dz = 0.f;
            acc += dz * __ldg(x + (int64_t)s * x_stride + k);
        }
        dW[(int64_t)warp * K + k] += acc;
    }
}

// dx[s][k] = sum_j dz[s][j] W[j][k]: CTA = (sample s, chunk of 64 output features), threads over k (coalesced W rows);
// partial sums meet in dx through fp32 atomics (dx zeroed by the launcher)
constexpr int kLinJChunk = 64;
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ W, const float* __restrict__ out,
                                                           const float* __restrict__ dout, float* __restrict__ dx, int64_t dx_stride,
                                                           int n, int K, int J, int relu) {
    __shared__ float dz[kLinJChunk];
    const int s = blockIdx.y, j0 = blockIdx.x * kLinJChunk;
    for (int i = threadIdx.x; i < kLinJChunk; i += blockDim.x) {
        float v = 0.f;
        if (j0 + i < J) {
            v = dout[(int64_t)s * J + j0 + i];
            if (relu && !(out[(int64_t)s * J + j0 + i] > 0.f)) v = 0.f;
        }
        dz[i] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float acc = 0.f;
        const int jn = min(kLinJChunk, J - j0);
        for (int i = 0; i < jn; ++i) acc += dz[i] * __ldg(W + (int64_t)(j0 + i) * K + k);
        atomicAdd(dx + (int64_t)s * dx_stride + k, acc);
    }
}

// ------------------------------------------------------------------------------------------ K16 discriminator head + GAN terms
// NSGAN term of one logit, F.binary_cross_entropy(F.sigmoid(o), t) for t in {0, 1} as PyTorch evaluates it in fp32 (sigmoid
// first, logs clamped at -100), and d term / d o through BCE's backward (denominator floored at 1e-12) and sigmoid's
__device__ __forceinline__ void nsgan_term(float o, float t, float& term, float& dterm) {
    const float p = 1.f / (1.f + expf(-o));
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
    term = -(t * lp + (1.f - t) * lq);
    dterm = (p - t) / fmaxf((1.f - p) * p, 1e-12f) * ((1.f - p) * p);
}

// one warp per pixel: logit = b + x[pixel] . w (lanes over channels, 8-channel vectors), GAN term and gradient seed
__global__ void __launch_bounds__(256) dis_head_fwd_kernel(aclgan_dis_head_args a) {
    __shared__ double sh[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = a.x.h * a.x.w;
    const int64_t npix = (int64_t)a.x.n * hw;
    const int n_per = a.x.n / a.groups;
    const int hp = a.x.h + 2 * a.x.pad, wp = a.x.w + 2 * a.x.pad;
    const float* w = reinterpret_cast<const float*>(a.weight);
    const float bias = __ldg(reinterpret_cast<const float*>(a.bias));
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t pix = (int64_t)blockIdx.x * 8 + warp; pix < npix; pix += (int64_t)gridDim.x * 8) {
        const int n = (int)(pix / hw), r = (int)(pix % hw);
        const int y = r / a.x.w, x = r % a.x.w;
        const int64_t base = (((int64_t)n * hp + y + a.x.pad) * wp + x + a.x.pad) * a.x.c;
        float s = 0.f;
        for (int c = lane; c < a.c_valid; c += 32) s += plane_val(a.x, base + c) * __ldg(w + c);
        s = warp_sum(s) + bias;
        const int g = n / n_per;
        if (lane == 0) {
            reinterpret_cast<float*>(a.logits)[pix] = s;
            if (a.gan_kind == ACLGAN_GAN_NSGAN) {
                float term, dterm;
                nsgan_term(s, a.target[g], term, dterm);
                if (a.dlogits != 0) reinterpret_cast<float*>(a.dlogits)[pix] = a.gweight[g] * dterm / (float)(n_per * hw);
                part[g] += (double)term;
            } else {
                const float diff = s - a.target[g];
                if (a.dlogits != 0)
                    reinterpret_cast<float*>(a.dlogits)[pix] = a.gweight[g] * 2.f * diff / (float)(n_per * hw);
                part[g] += (double)diff * (double)diff;
            }
        }
    }
    if (a.loss == 0) return;
    for (int g = 0; g < a.groups; ++g) {
        const double t = cta_sum_d(part[g], sh);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(reinterpret_cast<double*>(a.loss) + a.loss_slot[g], t / (double)(n_per * hw));
    }
}

// gr[pixel][c] = dlogit[pixel] * w[c] (dense gradient of the feature plane); dW[c] += sum_pixels dlogit x[pixel][c]; db += sum
__global__ void __launch_bounds__(256) dis_head_bwd_kernel(aclgan_dis_head_bwd_args a) {
    extern __shared__ float sh_dw[];     // [c_valid]
    const int hw = a.x.h * a.x.w;
    const int64_t npix = (int64_t)a.x.n * hw;
    const int hp = a.x.h + 2 * a.x.pad, wp = a.x.w + 2 * a.x.pad;
    const float* w = reinterpret_cast<const float*>(a.weight);
    const float* dl = reinterpret_cast<const float*>(a.dlogits);
    const bool tw = a.dweight != 0;
    for (int c = threadIdx.x; c < a.c_valid; c += blockDim.x) sh_dw[c] = 0.f;
    __syncthreads();
    // CTA handles a contiguous range of pixels; thread = channel (strided), loop over the CTA's pixels
    const int64_t per = (npix + gridDim.x - 1) / gridDim.x;
    const int64_t p0 = (int64_t)blockIdx.x * per, p1 = min(npix, p0 + per);
    float dbs = 0.f;
    for (int c = threadIdx.x; c < a.x.c; c += blockDim.x) {
        const float wv = c < a.c_valid ? __ldg(w + c) : 0.f;
        float acc = 0.f;
        for (int64_t pix = p0; pix < p1; ++pix) {
            const float d = __ldg(dl + pix);
            if (a.gr != 0) store_grad(a.gr, a.g_kind, pix * a.x.c + c, d * wv);
            if (tw && c < a.c_valid) {
                const int n = (int)(pix / hw), r = (int)(pix % hw);
                const int y = r / a.x.w, x = r % a.x.w;
                acc += d * plane_val(a.x, (((int64_t)n * hp + y + a.x.pad) * wp + x + a.x.pad) * a.x.c + c);
            }
            if (c == 0) dbs += d;
        }
        if (tw && c < a.c_valid) atomicAdd(reinterpret_cast<float*>(a.dweight) + c, acc);
    }
    if (tw && threadIdx.x == 0 && p1 > p0) atomicAdd(reinterpret_cast<float*>(a.dbias), dbs);
}

// ------------------------------------------------------------------------------------------ K14 focus blend
// dst = fg * m + bg * (1 - m), m = (mask + 1) / 2 broadcast over the 3 colour channels; separately rounded products and
// sum (no fma contraction), i.e. the same fp32 operations as x_fg * x_map + x_bg * (1 - x_map)
__global__ void blend_fwd_kernel(aclgan_blend_args a) {
    const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * hw;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t n = t / hw, p = t % hw;
    const float* o = reinterpret_cast<const float*>(a.out4) + n * 4 * hw + p;
    const float* bg = reinterpret_cast<const float*>(a.bg) + n * 3 * hw + p;
    float* d = reinterpret_cast<float*>(a.dst) + n * 3 * hw + p;
    const float m = __fdiv_rn(__fadd_rn(o[3 * hw], 1.f), 2.f);
    const float om = __fsub_rn(1.f, m);
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c * hw] = __fadd_rn(__fmul_rn(o[c * hw], m), __fmul_rn(bg[c * hw], om));
}

// d out4[:, c] (+)= d * m (c < 3);  d out4[:, 3] (+)= 0.5 * sum_c d_c (fg_c - bg_c);  d bg (+)= d * (1 - m)
__global__ void blend_bwd_kernel(aclgan_blend_args a) {
    const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * hw;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t n = t / hw, p = t % hw;
    const float* o = reinterpret_cast<const float*>(a.out4) + n * 4 * hw + p;
    const float* bg = reinterpret_cast<const float*>(a.bg) + n * 3 * hw + p;
    const float* dd = reinterpret_cast<const float*>(a.ddst) + n * 3 * hw + p;
    const float m = (o[3 * hw] + 1.f) * 0.5f;
    float dm = 0.f;
    float* dout = reinterpret_cast<float*>(a.dout4) + n * 4 * hw + p;
    float* dbg = a.dbg != 0 ? reinterpret_cast<float*>(a.dbg) + n * 3 * hw + p : nullptr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float d = dd[c * hw];
        dm += d * (o[c * hw] - bg[c * hw]);
        if (a.dout4 != 0) dout[c * hw] = a.acc_out4 ? dout[c * hw] + d * m : d * m;
        if (dbg != nullptr) dbg[c * hw] = a.acc_bg ? dbg[c * hw] + d * (1.f - m) : d * (1.f - m);
    }
    if (a.dout4 != 0) dout[3 * hw] = a.acc_out4 ? dout[3 * hw] + 0.5f * dm : 0.5f * dm;
}

// ------------------------------------------------------------------------------------------ K17 / K18 loss reductions
// L1: acc[slot] += mean |a - b| over the first c channels; da (+)= sign(a - b) * gscale
__global__ void __launch_bounds__(256) loss_l1_kernel(aclgan_loss_reduce_args a) {
    __shared__ double sh[8];
    const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * a.c * hw;
    double part = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / (a.c * hw), r = t % (a.c * hw);
        const int64_t ia = n * a.ca * hw + r;
        const float d = reinterpret_cast<const float*>(a.a)[ia] - reinterpret_cast<const float*>(a.b)[t];
        part += (double)fabsf(d);
        if (a.da != 0) {
            const float s = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * a.gscale;
            float* g = reinterpret_cast<float*>(a.da) + ia;
            *g = a.acc_da ? *g + s : s;
        }
    }
    const double tsum = cta_sum_d(part, sh);
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<double*>(a.acc) + a.slot, tsum / (double)total);
}

// focus pass 1: acc[slot] += sum (m - upper), acc[slot+1] += sum (lower - m), acc[slot+2] += sum 1 / (|m - 0.5| + eps)
// over the WHOLE batch (trainer.py:146-151), m = (out4[:, 3] + 1) / 2
__global__ void __launch_bounds__(256) loss_focus_kernel(aclgan_loss_reduce_args a) {
    __shared__ double sh[8];
    const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * hw;
    double s1 = 0.0, s2 = 0.0, dg = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / hw, p = t % hw;
        const float m = (reinterpret_cast<const float*>(a.a)[(n * a.ca + 3) * hw + p] + 1.f) * 0.5f;
        s1 += (double)(m - a.upper);
        s2 += (double)(a.lower - m);
        dg += (double)(1.f / (fabsf(m - 0.5f) + a.eps));
    }
    double* acc = reinterpret_cast<double*>(a.acc) + a.slot;
    double t = cta_sum_d(s1, sh);
    if (threadIdx.x == 0) atomicAdd(acc, t);
    t = cta_sum_d(s2, sh);
    if (threadIdx.x == 0) atomicAdd(acc + 1, t);
    t = cta_sum_d(dg, sh);
    if (threadIdx.x == 0) atomicAdd(acc + 2, t);
}

// focus pass 2: size loss from the batch sums, and d(size + digit)/d mask written into channel 3 of d out4
__global__ void __launch_bounds__(256) focus_grad_kernel(aclgan_focus_grad_args a) {
    const double* sums = reinterpret_cast<const double*>(a.sums) + a.slot;
    const float s1 = fmaxf((float)sums[0], 0.f), s2 = fmaxf((float)sums[1], 0.f);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.size_slot >= 0)
        reinterpret_cast<double*>(a.sums)[a.size_slot] = (double)(s1 * s1 * a.delta + s2 * s2 * a.delta);
    const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * hw;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / hw, p = t % hw;
        const int64_t i3 = (n * 4 + 3) * hw + p;
        const float m = (reinterpret_cast<const float*>(a.out4)[i3] + 1.f) * 0.5f;
        const float dev = m - 0.5f, den = fabsf(dev) + a.eps;
        const float sg = dev > 0.f ? 1.f : (dev < 0.f ? -1.f : 0.f);
        const float dm = (2.f * a.delta) * (s1 - s2) - sg / (den * den);
        float* g = reinterpret_cast<float*>(a.dout4) + i3;
        *g = a.acc ? *g + dm * a.gscale : dm * a.gscale;
    }
}

// out[j] = sum_k M[j][k] * acc[k]  (the weighted totals of trainer.py:142-165 / 288-290 and the per-term loss_* scalars)
__global__ void loss_combine_kernel(const double* __restrict__ acc, const float* __restrict__ M, float* __restrict__ out, int J, int K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += (double)M[(int64_t)j * K + k] * acc[k];
    out[j] = (float)s;
}

// ------------------------------------------------------------------------------------------ helpers
template <typename T>
__global__ void axpby_kernel(T* dst, const T* a, const T* b, float alpha, float beta, int64_t n);
template <>
__global__ void axpby_kernel<float>(float* dst, const float* a, const float* b, float alpha, float beta, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = alpha * a[i] + (b != nullptr ? beta * b[i] : 0.f);
}
template <>
__global__ void axpby_kernel<__nv_bfloat16>(__nv_bfloat16* dst, const __nv_bfloat16* a, const __nv_bfloat16* b, float alpha,
                                            float beta, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(alpha * __bfloat162float(a[i]) + (b != nullptr ? beta * __bfloat162float(b[i]) : 0.f));
}

// ------------------------------------------------------------------------------------------ input pipeline
// uint8 NHWC batch (decoded / resized / cropped by the loader workers) -> fp32 NCHW in [-1, 1] with per-sample horizontal flip:
// transforms.RandomHorizontalFlip + ToTensor + Normalize(0.5, 0.5) (reference utils.py:83-100) with the same rounding sequence
// (u8 -> float, / 255, - 0.5, / 0.5), so the result is bit-identical to the torchvision path
__global__ void augment_u8_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ flip, float* __restrict__ dst, int n,
                                  int h, int w) {
    const int64_t total = (int64_t)n * h * w;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int x = (int)(t % w), y = (int)((t / w) % h), i = (int)(t / ((int64_t)w * h));
    const int xs = (flip != nullptr && flip[i]) ? w - 1 - x : x;
    const uint8_t* s = src + (((int64_t)i * h + y) * w + xs) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        dst[(((int64_t)i * 3 + c) * h + y) * w + x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[c], 255.f), 0.5f), 0.5f);
}

// dst[c] += sum_n sums[n][c][0] (fp64 per-(n, c) sums of the generic backward-reduce kernel -> fp32 conv-bias gradient)
__global__ void stats_to_bias_kernel(const double* __restrict__ sums, float* __restrict__ dst, int n, int c, int c_valid) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c_valid) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += sums[((int64_t)i * c + ch) * 2];
    dst[ch] += (float)s;
}

static inline int grid1d(int64_t total, int block, int cap = 148 * 8) {
    int64_t g = (total + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace aclgan

using namespace aclgan;

static inline int pool_out(int h) { return (h + 2 - 3) / 2 + 1; }

extern "C" int aclgan_avgpool3x3s2_fwd(const aclgan_avgpool_args* a, void* stream) {
    if (a->h < 1 || a->w < 1 || a->planes < 1) return ACLGAN_ERR_SHAPE;
    const int ho = pool_out(a->h), wo = pool_out(a->w);
    const int64_t total = (int64_t)a->planes * ho * wo;
    avgpool_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a, ho, wo);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_avgpool3x3s2_bwd(const aclgan_avgpool_args* a, void* stream) {
    if (a->h < 1 || a->w < 1 || a->planes < 1) return ACLGAN_ERR_SHAPE;
    const int ho = pool_out(a->h), wo = pool_out(a->w);
    const int64_t total = (int64_t)a->planes * a->h * a->w;
    avgpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a, ho, wo);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_style_head_fwd(const aclgan_style_head_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->x.c || a->style_dim < 1) return ACLGAN_ERR_SHAPE;
    style_head_fwd_kernel<<<a->x.n, 256, a->c_valid * sizeof(float), (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_style_head_bwd(const aclgan_style_head_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->x.c || a->gr == 0) return ACLGAN_ERR_SHAPE;
    style_head_bwd_kernel<<<a->x.n + 1, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_mlp_fwd(const aclgan_mlp_args* a, void* stream) {
    if (a->n_layers < 1 || a->n_layers > 4 || a->n < 1) return ACLGAN_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 0; l < a->n_layers; ++l) {
        const int K = a->dims[l], J = a->dims[l + 1];
        const int64_t xs = l == 0 && a->h0_stride > 0 ? a->h0_stride : K;
        linear_fwd_kernel<<<(J * 32 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(a->h[l]), xs,
                                                                 reinterpret_cast<const float*>(a->w[l]),
                                                                 reinterpret_cast<const float*>(a->b[l]),
                                                                 reinterpret_cast<float*>(a->h[l + 1]), a->n, K, J,
                                                                 l + 1 < a->n_layers ? 1 : 0);
    }
    return (int)cudaGetLastError();
}

extern "C" int aclgan_mlp_bwd(const aclgan_mlp_args* a, void* stream) {
    if (a->n_layers < 1 || a->n_layers > 4 || a->n < 1 || a->dh[a->n_layers] == 0) return ACLGAN_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = a->n_layers - 1; l >= 0; --l) {
        const int K = a->dims[l], J = a->dims[l + 1];
        const int relu = l + 1 < a->n_layers ? 1 : 0;
        const int64_t xs = l == 0 && a->h0_stride > 0 ? a->h0_stride : K;
        const float* out = reinterpret_cast<const float*>(a->h[l + 1]);
        const float* dout = reinterpret_cast<const float*>(a->dh[l + 1]);
        if (a->dw[l] != 0)
            linear_bwd_w_kernel<<<(J * 32 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(a->h[l]), xs, out, dout,
                                                                       reinterpret_cast<float*>(a->dw[l]),
                                                                       reinterpret_cast<float*>(a->db[l]), a->n, K, J, relu);
        if (a->dh[l] != 0) {
            cudaError_t e = cudaMemsetAsync(reinterpret_cast<void*>(a->dh[l]), 0, (size_t)a->n * K * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
            dim3 grid((J + kLinJChunk - 1) / kLinJChunk, a->n);
            linear_bwd_x_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(a->w[l]), out, dout,
                                                      reinterpret_cast<float*>(a->dh[l]), K, a->n, K, J, relu);
        }
    }
    return (int)cudaGetLastError();
}

extern "C" int aclgan_dis_head_fwd(const aclgan_dis_head_args* a, void* stream) {
    if (a->groups < 1 || a->groups > 4 || a->x.n % a->groups != 0 || a->c_valid < 1 || a->c_valid > a->x.c) return ACLGAN_ERR_SHAPE;
    if (a->gan_kind != ACLGAN_GAN_LSGAN && a->gan_kind != ACLGAN_GAN_NSGAN) return ACLGAN_ERR_UNSUPPORTED;
    const int64_t npix = (int64_t)a->x.n * a->x.h * a->x.w;
    dis_head_fwd_kernel<<<grid1d(npix, 8, 148 * 2), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_dis_head_bwd(const aclgan_dis_head_bwd_args* a, void* stream) {
    if (a->c_valid < 1 || a->c_valid > a->x.c || a->dlogits == 0) return ACLGAN_ERR_SHAPE;
    const int64_t npix = (int64_t)a->x.n * a->x.h * a->x.w;
    dis_head_bwd_kernel<<<grid1d(npix, 16, 148), 256, a->c_valid * sizeof(float), (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_focus_blend_fwd(const aclgan_blend_args* a, void* stream) {
    const int64_t total = (int64_t)a->n * a->h * a->w;
    if (total < 1 || a->dst == 0) return ACLGAN_ERR_SHAPE;
    blend_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_focus_blend_bwd(const aclgan_blend_args* a, void* stream) {
    const int64_t total = (int64_t)a->n * a->h * a->w;
    if (total < 1 || a->ddst == 0) return ACLGAN_ERR_SHAPE;
    blend_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_loss_reduce(const aclgan_loss_reduce_args* a, void* stream) {
    if (a->acc == 0 || a->n < 1) return ACLGAN_ERR_SHAPE;
    if (a->mode == ACLGAN_LOSS_L1) {
        const int64_t total = (int64_t)a->n * a->c * a->h * a->w;
        loss_l1_kernel<<<grid1d(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>(*a);
    } else if (a->mode == ACLGAN_LOSS_FOCUS) {
        if (a->ca < 4) return ACLGAN_ERR_SHAPE;
        const int64_t total = (int64_t)a->n * a->h * a->w;
        loss_focus_kernel<<<grid1d(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>(*a);
    } else {
        return ACLGAN_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

extern "C" int aclgan_focus_grad(const aclgan_focus_grad_args* a, void* stream) {
    const int64_t total = (int64_t)a->n * a->h * a->w;
    if (total < 1 || a->dout4 == 0 || a->sums == 0) return ACLGAN_ERR_SHAPE;
    focus_grad_kernel<<<grid1d(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_loss_combine(uint64_t acc, uint64_t M, uint64_t out, int32_t J, int32_t K, void* stream) {
    if (J < 1 || K < 1) return ACLGAN_ERR_SHAPE;
    loss_combine_kernel<<<(J + 63) / 64, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(acc),
                                                                       reinterpret_cast<const float*>(M),
                                                                       reinterpret_cast<float*>(out), J, K);
    return (int)cudaGetLastError();
}

// dst = alpha * a + beta * b (b may be 0); kind 0 bf16, 1 fp32; dst may alias a or b
extern "C" int aclgan_axpby(uint64_t dst, uint64_t a, uint64_t b, float alpha, float beta, int64_t n, int32_t kind, void* stream) {
    if (n < 1) return ACLGAN_OK;
    if (kind == 1)
        axpby_kernel<float><<<grid1d(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), alpha, beta, n);
    else
        axpby_kernel<__nv_bfloat16><<<grid1d(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<const __nv_bfloat16*>(a),
            reinterpret_cast<const __nv_bfloat16*>(b), alpha, beta, n);
    return (int)cudaGetLastError();
}

// cudaMemsetAsync / cudaMemcpyAsync on the caller's stream (memset / memcpy nodes under capture - no fill / copy kernels)
extern "C" int aclgan_zero(uint64_t ptr, int64_t bytes, void* stream) {
    if (bytes <= 0) return ACLGAN_OK;
    return (int)cudaMemsetAsync(reinterpret_cast<void*>(ptr), 0, (size_t)bytes, (cudaStream_t)stream);
}
extern "C" int aclgan_copy(uint64_t dst, uint64_t src, int64_t bytes, void* stream) {
    if (bytes <= 0) return ACLGAN_OK;
    return (int)cudaMemcpyAsync(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(src), (size_t)bytes,
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}

extern "C" int aclgan_stats_to_bias(uint64_t sums, uint64_t dst, int32_t n, int32_t c, int32_t c_valid, void* stream) {
    if (n < 1 || c_valid < 1 || c_valid > c) return ACLGAN_ERR_SHAPE;
    stats_to_bias_kernel<<<(c_valid + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(sums),
                                                                                reinterpret_cast<float*>(dst), n, c, c_valid);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_augment_u8(uint64_t src, uint64_t flip, uint64_t dst, int32_t n, int32_t h, int32_t w, void* stream) {
    if (n < 1 || h < 1 || w < 1) return ACLGAN_ERR_SHAPE;
    const int64_t total = (int64_t)n * h * w;
    augment_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t*>(src), reinterpret_cast<const uint8_t*>(flip), reinterpret_cast<float*>(dst), n, h, w);
    return (int)cudaGetLastError();
}
