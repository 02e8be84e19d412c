// Sub-pixel form of the decoder's up-blocks: nearest-2x-upsample -> ReflectionPad2d(2) -> Conv2d 5x5 (reference
// networks.py:256-257; 2 x 13.4 of the decoder's 47 GMAC per image) WITHOUT materialising the up-sampled plane and with 9 instead
// of 25 taps per output pixel:
//
//   MAIN  the four output phases (py, px) of an output 2x2 block read the SAME 3x3 source neighbourhood, so they are folded
//         into the N dimension of ONE 3x3 stride-1 convolution of the reflect-pad-1 source plane with 4*Cout output channels
//         (tensor-core segment kernels of igemm.cu; weights = sums of the 5x5 taps that hit the same source pixel, derived here);
//         the epilogue stores depth-to-space and skips the source pixels on the border ring.
//   RING  there the reference reflects in UP-SAMPLED coordinates (row -1 reads S[0], row -2 reads S[1]): the two outermost
//         output rows / columns are recomputed exactly by the plain 5x5 convolution on four thin strips of the exactly padded
//         up-sampled plane, gathered here (columns as transposed strips, convolved with the transposed filter).
//
// This file holds the layout kernels around those GEMMs (the GEMMs themselves are ordinary plans of plans.cu): weight derivation,
// strip gather, the backward re-packing of dY (space-to-depth with the ring zeroed + ring strips), the gather of the strip input
// gradients back onto the source pixels, and the fold of the phase-weight gradient onto the 5x5 taps.
// Geometry validated index-for-index on the CPU by tools/subpixel_pipeline.py / tests/test_subpixel_math.py.
#include "common.cuh"

namespace aclgan {

__device__ __forceinline__ int up_reflect(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}
// source index read by coordinate y of the reflect-padded up-sampled axis (L = source length)
__device__ __forceinline__ int up_src(int y, int L) { return up_reflect(y, 2 * L) >> 1; }

// the 5x5 taps that read source offset u - 1 in output phase p: p = 0: {0,1},{2,3},{4};  p = 1: {0},{1,2},{3,4}
__device__ __forceinline__ int tap_first(int p, int u) { return p == 0 ? 2 * u : (u == 0 ? 0 : 2 * u - 1); }
__device__ __forceinline__ int tap_count(int p, int u) { return p == 0 ? (u < 2 ? 2 : 1) : (u == 0 ? 1 : 2); }
__device__ __forceinline__ int tap_source(int p, int a) { return p == 0 ? a / 2 : (a + 1) / 2; }

__device__ __forceinline__ float up_bf16(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// ------------------------------------------------------------------------------------------ phase weights
__global__ void up_derive_kernel(aclgan_up_derive_args a) {
    const int64_t total = (int64_t)4 * a.co * a.ci * 9;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4 * a.co && a.bias4 != 0)
        reinterpret_cast<float*>(a.bias4)[t] = __ldg(reinterpret_cast<const float*>(a.bias) + (t % a.co));
    if (t >= total) return;
    const int v = (int)(t % 3), u = (int)((t / 3) % 3);
    const int ci = (int)((t / 9) % a.ci);
    const int cop = (int)(t / (9 * (int64_t)a.ci));
    const int ph = cop / a.co, co = cop - ph * a.co, py = ph >> 1, px = ph & 1;
    const float* w = reinterpret_cast<const float*>(a.w5) + ((int64_t)co * a.ci + ci) * 25;
    float s = 0.f;
    for (int i = 0; i < tap_count(py, u); ++i)
        for (int j = 0; j < tap_count(px, v); ++j) s += __ldg(w + (tap_first(py, u) + i) * 5 + tap_first(px, v) + j);
    const __nv_bfloat16 hi = __float2bfloat16_rn(s);
    const __nv_bfloat16 lo = __float2bfloat16_rn(s - __bfloat162float(hi));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (a.pk[k][0] == 0) continue;
        const int64_t o = a.aff[k][0] + cop * a.aff[k][1] + ci * a.aff[k][2] + u * a.aff[k][3] + v * a.aff[k][4];
        reinterpret_cast<__nv_bfloat16*>(a.pk[k][0])[o] = hi;
        if (a.planes == 2) reinterpret_cast<__nv_bfloat16*>(a.pk[k][1])[o] = lo;
    }
}

// ------------------------------------------------------------------------------------------ forward strips
// one thread per (strip pixel, 8-channel group): 16-byte copies of the raw bf16 plane(s)
__global__ void up_gather_strips_kernel(aclgan_up_strips_args a) {
    const int n = a.src.n, H = a.src.h, W = a.src.w, cg = a.src.c / 8;
    const int sp = a.src.pad, swp = W + 2 * sp, shp = H + 2 * sp;
    const int rw = 2 * W + 4, cw = 2 * H;              // stored strip widths
    const int64_t n_rows = (int64_t)2 * n * 6 * rw, n_cols = (int64_t)2 * n * 6 * cw;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_rows + n_cols) * cg) return;
    const int g = (int)(t % cg);
    int64_t pix = t / cg;
    const bool is_col = pix >= n_rows;
    if (is_col) pix -= n_rows;
    const int sw = is_col ? cw : rw;
    const int xx = (int)(pix % sw), yy = (int)((pix / sw) % 6), z = (int)(pix / ((int64_t)sw * 6));
    const int side = z / n, img = z - side * n;
    int i, j;
    if (!is_col) {
        i = up_src(side ? 2 * H - 4 + yy : yy - 2, H);
        j = up_src(xx - 2, W);
    } else {                                            // transposed: strip row = up-sampled column, strip column = up-sampled row
        j = up_src(side ? 2 * W - 4 + yy : yy - 2, W);
        i = xx >> 1;
    }
    const int64_t so = (((int64_t)img * shp + i + sp) * swp + j + sp) * a.src.c + g * 8;
    const aclgan_act& d = is_col ? a.cols : a.rows;
    const int64_t dof = pix * d.c + g * 8;
    for (int p = 0; p < a.src.planes; ++p)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.data[p]) + dof) =
            *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.src.data[p]) + so);
}

// ------------------------------------------------------------------------------------------ backward: dY re-packing
// which: 0 = space-to-depth plane (ring source pixels and the border zero), 1 = ring rows, 2 = ring columns (transposed)
__global__ void up_dy_pack_kernel(aclgan_up_dy_pack_args a, int which) {
    const int n = a.dy.n, H2 = a.dy.h, W2 = a.dy.w, H = H2 / 2, W = W2 / 2;
    const aclgan_act& d = which == 0 ? a.s2d : (which == 1 ? a.rows : a.cols);
    const int cg = d.c / 8, hp = d.h + 2 * d.pad, wp = d.w + 2 * d.pad;
    const int64_t total = (int64_t)d.n * hp * wp * cg;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int g = (int)(t % cg);
    const int64_t pix = t / cg;
    const int X = (int)(pix % wp) - d.pad, Y = (int)((pix / wp) % hp) - d.pad, z = (int)(pix / ((int64_t)wp * hp));
    int64_t so = -1;                                    // source element index in dy, or -1 = zero
    if (X >= 0 && X < d.w && Y >= 0 && Y < d.h) {
        if (which == 0) {
            const int c0 = g * 8, ph = c0 / a.cout, c = c0 - ph * a.cout;
            const bool ring = X == 0 || Y == 0 || X == W - 1 || Y == H - 1;
            if (!ring && ph < 4) so = (((int64_t)z * H2 + 2 * Y + (ph >> 1)) * W2 + 2 * X + (ph & 1)) * a.dy.c + c;
        } else {
            const int side = z / n, img = z - side * n;
            if (g * 8 < a.dy.c) {
                if (which == 1) so = (((int64_t)img * H2 + (side ? H2 - 2 + Y : Y)) * W2 + X) * a.dy.c + g * 8;
                else so = (((int64_t)img * H2 + X + 2) * W2 + (side ? W2 - 2 + Y : Y)) * a.dy.c + g * 8;
            }
        }
    }
    const int64_t dof = pix * d.c + g * 8;
    for (int p = 0; p < a.dy.planes; ++p) {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (so >= 0) q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dy.data[p]) + so);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.data[p]) + dof) = q;
    }
}

// ------------------------------------------------------------------------------------------ backward: strip gradients -> source
__device__ __forceinline__ void add8(float (&acc)[8], uint64_t base, int kind, int64_t idx) {
    if (kind == 0) {
        const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[2 * i] += __uint_as_float(w[i] << 16);
            acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
        const float4 x = p[0], y = p[1];
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
    }
}

// one thread per (source pixel of the 2-pixel border band, 8-channel group): gathers every strip element that was copied from
// this source pixel (deterministic, no atomics) and adds the sum to the interior of the padded-plane gradient
__global__ void up_scatter_kernel(aclgan_up_scatter_args a) {
    const int H = a.h, W = a.w, cg = a.c / 8;
    const int64_t total = (int64_t)a.n * H * W * cg;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int g = (int)(t % cg);
    const int64_t pix = t / cg;
    const int j = (int)(pix % W), i = (int)((pix / W) % H), img = (int)(pix / ((int64_t)W * H));
    if (!(i < 2 || i >= H - 2 || j < 2 || j >= W - 2)) return;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int rw = 2 * W + 4, cw = 2 * H;
    // up-sampled columns that read source column j: 2j, 2j+1 and their reflections inside the 2-wide padding
    int xs[4], nx = 0;
    xs[nx++] = 2 * j; xs[nx++] = 2 * j + 1;
    if (j == 0) xs[nx++] = -1;
    if (j == 1) xs[nx++] = -2;
    if (j == W - 1) xs[nx++] = 2 * W;
    if (j == W - 2) xs[nx++] = 2 * W + 1;
    for (int side = 0; side < 2; ++side) {
        const int z = side * a.n + img;
        for (int yy = 0; yy < 6; ++yy) {
            if (up_src(side ? 2 * H - 4 + yy : yy - 2, H) != i) continue;
            for (int k = 0; k < nx; ++k) add8(acc, a.grows, a.kind, (((int64_t)z * 6 + yy) * rw + xs[k] + 2) * a.c + g * 8);
        }
        for (int xx = 0; xx < 6; ++xx) {
            if (up_src(side ? 2 * W - 4 + xx : xx - 2, W) != j) continue;
            add8(acc, a.gcols, a.kind, (((int64_t)z * 6 + xx) * cw + 2 * i) * a.c + g * 8);
            add8(acc, a.gcols, a.kind, (((int64_t)z * 6 + xx) * cw + 2 * i + 1) * a.c + g * 8);
        }
    }
    const int64_t o = (((int64_t)img * (H + 2) + i + 1) * (W + 2) + j + 1) * a.c + g * 8;
    if (a.kind == 0) {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(a.g) + o;
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = __float2bfloat16_rn(__bfloat162float(d[k]) + acc[k]);
    } else {
        float* d = reinterpret_cast<float*>(a.g) + o;
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] += acc[k];
    }
}

// ------------------------------------------------------------------------------------------ backward: phase dW -> 5x5 dW
__global__ void up_fold_wgrad_kernel(aclgan_up_fold_wgrad_args a) {
    const int64_t total = (int64_t)a.co * a.ci * 25;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = (int)(t % 5), aa = (int)((t / 5) % 5);
    const int ci = (int)((t / 25) % a.ci), co = (int)(t / (25 * (int64_t)a.ci));
    const float* src = reinterpret_cast<const float*>(a.dwp);
    float s = 0.f;
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const int u = tap_source(ph >> 1, aa), v = tap_source(ph & 1, b);
        s += src[a.affp[0] + (int64_t)(ph * a.co + co) * a.affp[1] + ci * a.affp[2] + u * a.affp[3] + v * a.affp[4]];
    }
    atomicAdd(reinterpret_cast<float*>(a.dw5) + a.aff5[0] + co * a.aff5[1] + ci * a.aff5[2] + aa * a.aff5[3] + b * a.aff5[4], s);
}

static inline unsigned blocks_for(int64_t total) { return (unsigned)((total + 255) / 256); }

}  // namespace aclgan

using namespace aclgan;

extern "C" int aclgan_up_derive_weights(const aclgan_up_derive_args* a, void* stream) {
    if (a->co < 1 || a->ci < 1 || a->planes < 1 || a->planes > 2) return ACLGAN_ERR_SHAPE;
    up_derive_kernel<<<blocks_for((int64_t)4 * a->co * a->ci * 9), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_up_gather_strips(const aclgan_up_strips_args* a, void* stream) {
    const int n = a->src.n, H = a->src.h, W = a->src.w;
    if (H < 3 || W < 3 || a->src.c % 8 != 0 || a->src.pad < 0) return ACLGAN_ERR_SHAPE;
    if (a->rows.n != 2 * n || a->rows.h != 2 || a->rows.w != 2 * W || a->rows.pad != 2 || a->rows.c != a->src.c) return ACLGAN_ERR_SHAPE;
    if (a->cols.n != 2 * n || a->cols.h != 2 || a->cols.w != 2 * H - 4 || a->cols.pad != 2 || a->cols.c != a->src.c) return ACLGAN_ERR_SHAPE;
    const int64_t total = ((int64_t)2 * n * 6 * (2 * W + 4) + (int64_t)2 * n * 6 * (2 * H)) * (a->src.c / 8);
    up_gather_strips_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_up_dy_pack(const aclgan_up_dy_pack_args* a, void* stream) {
    const int n = a->dy.n, H2 = a->dy.h, W2 = a->dy.w;
    if (a->dy.pad != 0 || H2 % 2 || W2 % 2 || H2 < 6 || W2 < 6 || a->cout % 8 != 0 || a->cout > a->dy.c) return ACLGAN_ERR_SHAPE;
    if (a->s2d.n != n || a->s2d.h != H2 / 2 || a->s2d.w != W2 / 2 || a->s2d.c < 4 * a->cout || a->s2d.c % 8) return ACLGAN_ERR_SHAPE;
    if (a->rows.n != 2 * n || a->rows.h != 2 || a->rows.w != W2 || a->rows.c < a->cout) return ACLGAN_ERR_SHAPE;
    if (a->cols.n != 2 * n || a->cols.h != 2 || a->cols.w != H2 - 4 || a->cols.c < a->cout) return ACLGAN_ERR_SHAPE;
    for (int which = 0; which < 3; ++which) {
        const aclgan_act& d = which == 0 ? a->s2d : (which == 1 ? a->rows : a->cols);
        const int64_t total = (int64_t)d.n * (d.h + 2 * d.pad) * (d.w + 2 * d.pad) * (d.c / 8);
        up_dy_pack_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*a, which);
    }
    return (int)cudaGetLastError();
}

extern "C" int aclgan_up_scatter_strips(const aclgan_up_scatter_args* a, void* stream) {
    if (a->h < 3 || a->w < 3 || a->c % 8 != 0 || a->n < 1) return ACLGAN_ERR_SHAPE;
    up_scatter_kernel<<<blocks_for((int64_t)a->n * a->h * a->w * (a->c / 8)), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int aclgan_up_fold_wgrad(const aclgan_up_fold_wgrad_args* a, void* stream) {
    if (a->co < 1 || a->ci < 1) return ACLGAN_ERR_SHAPE;
    up_fold_wgrad_kernel<<<blocks_for((int64_t)a->co * a->ci * 25), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}
