// Implicit-GEMM convolution (forward and data-gradient) on the 5th-gen tensor cores.
//
//   D[pixel][n] = sum_{segment} sum_{tap} sum_{chunk}  A_tap[pixel][64 ch] . B[n][k(tap, chunk) .. +64]
//
// * A tiles (128 pixels x 64 channels, bf16) are fetched straight from the padded NHWC activation in HBM
//   by TMA (rank-4 tensor maps, one box per filter tap: the reflect padding / stride-2 parity / small-C
//   pixel windows are all expressed in the tensor map + per-tap box offsets, nothing is im2col-materialised).
// * B tiles (block_n x 64, bf16, K-major) come from the packed weights by TMA.
// * both land in 128B-swizzled shared memory and feed tcgen05.mma (M=128, N=block_n, K=16) issued by one
//   thread; the fp32 accumulator lives in TMEM (2 x 256 columns, double buffered across tiles).
// * epilogue warps read TMEM with tcgen05.ld, add bias, apply the activation, and store bf16 / split-bf16 /
//   fp32 with arbitrary (n, y, x, c) strides, optionally replicating border pixels into a reflect-pad halo.
// * persistent CTAs (one per SM), 4-stage TMA->MMA mbarrier pipeline, warp-specialised roles.
//
// Replaces: nn.Conv2d forward inside Conv2dBlock.forward (reference networks.py:363,366) incl. the
// ReflectionPad2d gather (networks.py:319) and, for no-norm blocks, bias + ReLU/LeakyReLU/tanh
// (networks.py:345-353,370); the same kernel computes conv data gradients (autograd of networks.py:366).
#include <cstdlib>

#include "common.cuh"

namespace aclgan {

constexpr int kMaxStages = 4;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;          // 128 rows x 64 bf16
constexpr int kBBytesMax = 256 * 128;          // up to 256 rows x 64 bf16
constexpr int kPipeBytes = kMaxStages * (kABytes + kBBytesMax);   // 192 KB ring: 4 stages (M=128) or 3 stages (M=256)
constexpr int kStageOutBytes = 4 * 2 * 4096;    // epilogue staging: 4 warps x 2 planes x (32 rows x 128 B)
constexpr int kSmemBytes = kPipeBytes + kStageOutBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 256;

struct alignas(64) IgemmKParams {
    CUtensorMap a[2][ACLGAN_MAX_AVARIANTS];
    CUtensorMap b[2];
    int planes, nseg, block_n, n_tiles;
    int m_sub;   // 128-pixel tiles per CTA work item (1 | 2): with 2, every weight (B) stage feeds two A tiles
    int debug;   // perf triage only (env ACLGAN_IGEMM_DEBUG): 1 = MMA without TMA traffic, 2 = TMA without MMA
    int box_x, box_y, box_z, tiles_x, tiles_y, tiles_z;
    int cchunks, num_taps;
    int flat, flat_w, flat_img;
    int n_groups, group_taps;
    long long group_off[4];
    int tap_dx[ACLGAN_MAX_TAPS];
    int tap_dy[ACLGAN_MAX_TAPS];
    int tap_var[ACLGAN_MAX_TAPS];
    int tap_bk[ACLGAN_MAX_TAPS];
    // segment mode (igemm_seg_kernel)
    CUtensorMap a_seg[2];
    int seg_rows, num_segs, seg_taps;
    int seg_a_bytes, seg_b_bytes, seg_na, seg_nb;   // shared-memory ring geometry chosen by the host
    int seg_bo;                                     // 1: descriptors carry the matrix base offset of the shifted start row
    int seg_dx[16], seg_dy[16];
    int tap_row[ACLGAN_MAX_TAPS];
    aclgan_out_spec out;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == ACLGAN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACLGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == ACLGAN_ACT_TANH) return tanhf(v);
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// reflect-pad images of coordinate c in [0, L) for halo width p:  -c (1 <= c <= p) and 2(L-1)-c (L-1-p <= c <= L-2)
__device__ __forceinline__ int mirror_coords(int c, int L, int p, int (&out)[3]) {
    int n = 0;
    out[n++] = c;
    if (p > 0) {
        if (c >= 1 && c <= p) out[n++] = -c;
        if (c >= L - 1 - p && c <= L - 2) out[n++] = 2 * (L - 1) - c;
    }
    return n;
}

// what one epilogue thread knows about its accumulator row (= output pixel)
struct RowCtx {
    int x, y, z;
    bool valid;
    int64_t pix0;          // element offset of the pixel (channel 0) in the output
    int ys[3], xs[3];      // the pixel's own coordinates + its reflect-halo replicas
    int ny, nx;
};

// ---- generic (cold) path: any output kind / stride / partial channel count; scalar, rolled loops (small code) ----
__device__ __noinline__ void store_generic(const aclgan_out_spec& o, int64_t pix, int ch0, int cnt, const float* v) {
    if (o.kind == ACLGAN_OUT_BF16 || o.kind == ACLGAN_OUT_SPLIT) {
        __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(o.ptr[0]) + pix + (int64_t)ch0 * o.sc;
        __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(o.ptr[1]) + pix + (int64_t)ch0 * o.sc;
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[i]);
            d0[(int64_t)i * o.sc] = hi;
            if (o.kind == ACLGAN_OUT_SPLIT) d1[(int64_t)i * o.sc] = __float2bfloat16_rn(v[i] - __bfloat162float(hi));
        }
    } else {
        float* d = reinterpret_cast<float*>(o.ptr[0]) + pix + (int64_t)ch0 * o.sc;
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            if (o.kind == ACLGAN_OUT_F32_ATOMIC) atomicAdd(d + (int64_t)i * o.sc, v[i]);
            else d[(int64_t)i * o.sc] = v[i];
        }
    }
}

__device__ __noinline__ void epilogue_tile_generic(const IgemmKParams& P, uint32_t t_row, int n0, const RowCtx& rc) {
    const aclgan_out_spec& o = P.out;
    const float* bias = reinterpret_cast<const float*>(o.bias);
    const int step = P.block_n >= 32 ? 32 : 16;
#pragma unroll 1
    for (int c = 0; c < P.block_n; c += step) {
        uint32_t raw[32];
        if (step == 32) {
            tmem_ld_32x32(t_row + c, raw);
        } else {
            uint32_t r16[16];
            tmem_ld_32x16(t_row + c, r16);
#pragma unroll
            for (int i = 0; i < 16; ++i) raw[i] = r16[i];
        }
        tmem_ld_wait();
        const int ch0 = n0 + c;
        int cnt = o.C - ch0;
        if (cnt > step) cnt = step;
        if (!rc.valid || cnt <= 0) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            float t = v[i];
            if (bias != nullptr && ch0 + i < o.bias_n) t += __ldg(bias + ch0 + i);
            v[i] = apply_act(t, o.act, o.slope);
        }
#pragma unroll 1
        for (int iy = 0; iy < rc.ny; ++iy)
#pragma unroll 1
            for (int ix = 0; ix < rc.nx; ++ix)
                store_generic(o, rc.pix0 + (int64_t)(rc.ys[iy] - rc.y) * o.sy + (int64_t)(rc.xs[ix] - rc.x) * o.sx, ch0,
                              cnt, v);
    }
}

// ---- fast (hot) path: channel-contiguous bf16 / split-bf16 / fp32 output, all block_n channels stored ----
// One warp's 32 rows x 128 B staging tile: 16-byte pieces XOR-swizzled by (row & 7) so both the row-wise writes
// (each lane its own row) and the line-wise reads (8 lanes per row) are bank-conflict free.
__device__ __forceinline__ void stage_piece(uint8_t* stg, int lane, int piece, const uint4& q) {
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((piece ^ (lane & 7)) << 4)) = q;
}

// staged 32 rows x 128 B -> global memory as full 128-byte lines (4 rows per store instruction), then the reflect-halo
// replicas of border pixels (each lane copies its own row again, read back from the staging tile)
__device__ __forceinline__ void flush_rows(const aclgan_out_spec& o, const uint8_t* stg, uint8_t* base, int64_t row_byte_off,
                                           const RowCtx& rc, int esz, int lane) {
    __syncwarp();
    const int piece = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3);
        const int64_t off = __shfl_sync(0xffffffffu, row_byte_off, r);
        const int ok = __shfl_sync(0xffffffffu, (int)rc.valid, r);
        if (ok) {
            const uint4 q = *reinterpret_cast<const uint4*>(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + off + (piece << 4)) = q;
        }
    }
    if (rc.valid && rc.ny * rc.nx > 1) {
#pragma unroll 1
        for (int iy = 0; iy < rc.ny; ++iy)
#pragma unroll 1
            for (int ix = (iy == 0 ? 1 : 0); ix < rc.nx; ++ix) {
                uint8_t* dst = base + row_byte_off +
                               ((int64_t)(rc.ys[iy] - rc.y) * o.sy + (int64_t)(rc.xs[ix] - rc.x) * o.sx) * esz;
#pragma unroll
                for (int pc = 0; pc < 8; ++pc)
                    *reinterpret_cast<uint4*>(dst + (pc << 4)) =
                        *reinterpret_cast<const uint4*>(stg + lane * 128 + ((pc ^ (lane & 7)) << 4));
            }
    }
    __syncwarp();
}

// per-channel sum / sum of squares over the 32 staged rows of this warp (the values exactly as stored: bf16-rounded
// or fp32), written to the warp's reduction slots red[set][channel - n0][2].  Rows [0, rb) belong to the tile's first
// image (set 0), rows [rb, 32) to the next one (set 1; only tiles of the flattened grid straddle two images).
__device__ __forceinline__ void stage_colsums(const uint8_t* stg, float* red, int ch_local0, bool f32, int lane, int rb) {
    const int piece = lane >> 2, word = (lane & 3) << 2;
#pragma unroll 1
    for (int set = 0; set < 2; ++set) {
        const int r0 = set == 0 ? 0 : rb, r1 = set == 0 ? rb : 32;
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(stg + r * 128 + ((piece ^ (r & 7)) << 4) + word);
            if (f32) {
                const float x = __uint_as_float(w);
                s0 += x; q0 += x * x;
            } else {
                const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xFFFF0000u);
                s0 += x0; q0 += x0 * x0; s1 += x1; q1 += x1 * x1;
            }
        }
        float* rs = red + set * 512;
        if (f32) {
            rs[(ch_local0 + lane) * 2] = s0; rs[(ch_local0 + lane) * 2 + 1] = q0;
        } else {
            *reinterpret_cast<float4*>(rs + (ch_local0 + 2 * lane) * 2) = make_float4(s0, q0, s1, q1);
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// end of a tile with statistics: the four epilogue warps' slots -> fp64 atomics into stats[n][C][2].
// combine != 0: the tile's rows belong to image n_first (set 0) and, when `straddle`, n_first + 1 (set 1); the four warps
// are summed first (4x fewer atomics).  combine == 0 (tiny planes, several images per tile): every warp's 32 rows
// belong to the single image n_first (the host guarantees box_x * box_y % 32 == 0) and the warp adds its own slots.
__device__ __forceinline__ void stats_tile_end(const IgemmKParams& P, uint8_t* stage_out, int q, int lane, int n0, bool combine,
                                               int n_first, bool straddle, bool tile_ok) {
    const aclgan_out_spec& o = P.out;
    double* stats = reinterpret_cast<double*>(o.stats);
    if (combine) {
        named_bar_sync(1, 128);
        if (tile_ok) {
            for (int set = 0; set < (straddle ? 2 : 1); ++set) {
                const int n = n_first + set;
                if (n >= o.N) break;
                for (int ch = q * 32 + lane; ch < P.block_n; ch += 128) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float2 v = *reinterpret_cast<const float2*>(stage_out + w * 8192 + 4096 + set * 2048 + ch * 8);
                        s += v.x; qq += v.y;
                    }
                    double* d = stats + ((int64_t)n * o.C + n0 + ch) * 2;
                    atomicAdd(d, (double)s);
                    atomicAdd(d + 1, (double)qq);
                }
            }
        }
        named_bar_sync(1, 128);
    } else {
        __syncwarp();
        if (tile_ok && n_first < o.N) {
            const float* red = reinterpret_cast<const float*>(stage_out + q * 8192 + 4096);
            for (int ch = lane; ch < P.block_n; ch += 32) {
                double* d = stats + ((int64_t)n_first * o.C + n0 + ch) * 2;
                atomicAdd(d, (double)red[ch * 2]);
                atomicAdd(d + 1, (double)red[ch * 2 + 1]);
            }
        }
        __syncwarp();
    }
}

// statistics bookkeeping of one tile for one epilogue warp
struct StatCtx {
    bool combine, straddle;
    int n_first;     // image of set 0
    int rb;          // first row of this warp that belongs to image n_first + 1 (32: none)
};

__device__ __forceinline__ StatCtx make_stat_ctx(const IgemmKParams& P, int tx, int tz, int q, int lane, const RowCtx& rc) {
    StatCtx sc;
    if (P.flat) {
        const int64_t q0 = (int64_t)tx * P.box_x;
        const int z0 = (int)(q0 / P.flat_img), z1 = (int)((q0 + 127) / P.flat_img);
        sc.combine = true;
        sc.n_first = z0;
        sc.straddle = z1 != z0;
        const uint32_t m = __ballot_sync(0xffffffffu, rc.z != z0);
        sc.rb = m == 0 ? 32 : __ffs(m) - 1;
    } else if (P.box_z == 1) {
        sc.combine = true; sc.straddle = false; sc.n_first = tz; sc.rb = 32;
    } else {
        sc.combine = false; sc.straddle = false; sc.rb = 32;
        sc.n_first = tz * P.box_z + (q * 32) / (P.box_x * P.box_y);
    }
    return sc;
}

__device__ __noinline__ void epilogue_tile_fast(const IgemmKParams& P, uint32_t t_row, int n0, const RowCtx& rc, uint8_t* stg,
                                                int lane, int stat_rb) {
    const aclgan_out_spec& o = P.out;
    const bool f32 = (o.kind == ACLGAN_OUT_F32);
    const bool st = o.stats != 0;
    const int esz = f32 ? 4 : 2;
    const float* bias = reinterpret_cast<const float*>(o.bias);
#pragma unroll 1
    for (int c = 0; c < P.block_n; c += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c, raw);
        const int ch0 = n0 + c;
        // one bias value per lane, broadcast with shuffles (instead of 32 loads per thread)
        float bl = 0.f;
        if (bias != nullptr && ch0 + lane < o.bias_n) bl = __ldg(bias + ch0 + lane);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + __shfl_sync(0xffffffffu, bl, i);
        if (o.act == ACLGAN_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (o.act == ACLGAN_ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * o.slope;
        } else if (o.act == ACLGAN_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
        }
        if (st && !rc.valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;       // rows outside the output contribute nothing to the statistics
        }
        if (rc.valid || st) {
            if (f32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 qv;
                    qv.x = __float_as_uint(v[4 * i]); qv.y = __float_as_uint(v[4 * i + 1]);
                    qv.z = __float_as_uint(v[4 * i + 2]); qv.w = __float_as_uint(v[4 * i + 3]);
                    stage_piece(stg, lane, i, qv);
                }
            } else {
                const int piece0 = (c & 32) ? 4 : 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 qv;
                    qv.x = pack_bf16x2(v[8 * i], v[8 * i + 1]); qv.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                    qv.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); qv.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                    stage_piece(stg, lane, piece0 + i, qv);
                }
                if (o.kind == ACLGAN_OUT_SPLIT) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 qv;
                        qv.x = pack_bf16x2(v[8 * i], v[8 * i + 1]); qv.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                        qv.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); qv.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                        stage_piece(stg + 4096, lane, piece0 + i, qv);
                    }
                }
            }
        }
        if (f32 || (c & 32)) {      // a full 128-byte row segment is staged: 32 fp32 or 64 bf16 channels
            const int g0 = f32 ? ch0 : ch0 - 32;
            const int64_t row_off = (rc.pix0 + g0) * esz;
            if (st) {
                __syncwarp();
                stage_colsums(stg, reinterpret_cast<float*>(stg + 4096), g0 - n0, f32, lane, stat_rb);
            }
            flush_rows(o, stg, reinterpret_cast<uint8_t*>(o.ptr[0]), row_off, rc, esz, lane);
            if (o.kind == ACLGAN_OUT_SPLIT)
                flush_rows(o, stg + 4096, reinterpret_cast<uint8_t*>(o.ptr[1]), row_off, rc, esz, lane);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_out = smem + kPipeBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPipeBytes + kStageOutBytes);
    uint64_t* full_bar = bars;                     // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;       // [kMaxStages]
    uint64_t* tfull_bar = bars + 2 * kMaxStages;   // [2]
    uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
    const int m_sub = P.m_sub;
    const int stage_bytes = m_sub * kABytes + kBBytesMax;
    const int num_stages = kPipeBytes / stage_bytes;           // 4 or 3
    // TMEM: one accumulator set = m_sub x col_stride columns; two sets (double buffering) when they fit in 512
    const int col_stride = P.block_n < 32 ? 32 : P.block_n;
    const int set_cols = m_sub * col_stride;
    const int acc_sets = (2 * set_cols <= 512) ? 2 : 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < P.planes; ++p) {
            for (int v = 0; v < ACLGAN_MAX_AVARIANTS; ++v) tma_prefetch_desc(&P.a[p][v]);
            tma_prefetch_desc(&P.b[p]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = P.tiles_x * P.tiles_y * P.tiles_z;
    const int m_items = (m_tiles + m_sub - 1) / m_sub;
    const int group_items = m_items * P.n_tiles;
    const int total_tiles = group_items * P.n_groups;          // CTA work items (x output-parity groups)
    const int k_iters = P.nseg * P.group_taps * P.cchunks;
    const uint32_t stage_tx = m_sub * kABytes + P.block_n * 128;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int grp = tile / group_items;
            const int gt = tile % group_items;
            const int nt = gt % P.n_tiles;
            const int mi = gt / P.n_tiles;
            const int tap0 = grp * P.group_taps;
            int x0[2], y0[2], z0[2];
            for (int s = 0; s < m_sub; ++s) {
                int mt = mi * m_sub + s;       // a tile index past the end decodes to z >= N: zero-filled, never stored
                x0[s] = (mt % P.tiles_x) * P.box_x;
                mt /= P.tiles_x;
                y0[s] = (mt % P.tiles_y) * P.box_y;
                z0[s] = (mt / P.tiles_y) * P.box_z;
            }
#pragma unroll 1
            for (int seg = 0; seg < P.nseg; ++seg) {
                const int pa = (seg == 2) ? 1 : 0;   // A plane: hi, hi, lo
                const int pb = (seg == 1) ? 1 : 0;   // B plane: hi, lo, hi
#pragma unroll 1
                for (int t = tap0; t < tap0 + P.group_taps; ++t) {
                    const CUtensorMap* am = &P.a[pa][P.tap_var[t]];
                    const int dx = P.tap_dx[t], dy = P.tap_dy[t];
                    const int bk = P.tap_bk[t];
#pragma unroll 1
                    for (int cc = 0; cc < P.cchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * stage_bytes;
                        uint8_t* sb = sa + m_sub * kABytes;
                        if (P.debug == 1) {
                            mbar_arrive(&full_bar[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
#pragma unroll 1
                            for (int s = 0; s < m_sub; ++s)
                                tma_load_4d(sa + s * kABytes, am, &full_bar[stage], cc * 64, x0[s] + dx, y0[s] + dy, z0[s]);
                            tma_load_2d(sb, &P.b[pb], &full_bar[stage], bk + cc * 64, nt * P.block_n);
                        }
                        if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)P.block_n, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it % acc_sets;
            const uint32_t acc_phase = (it / acc_sets) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * set_cols;
#pragma unroll 1
            for (int k = 0; k < k_iters; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                const uint32_t sb = sa + m_sub * kABytes;
                const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
                if (P.debug == 2) {
                    mbar_arrive(&empty_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    continue;
                }
#pragma unroll 1
                for (int s = 0; s < m_sub; ++s) {
                    const uint64_t da = make_smem_desc_sw128(sa + s * kABytes, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        // +32 bytes (= 16 bf16 of K) inside the 128B swizzle row -> +2 in the >>4 encoded address
                        umma_bf16(d_tmem + s * col_stride, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
            if (P.debug == 2) mbar_arrive(&tfull_bar[acc]);
            else umma_commit(&tfull_bar[acc]);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue ----------------
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it % acc_sets;
            const uint32_t acc_phase = (it / acc_sets) & 1;
            const int grp = tile / group_items;
            const int gt = tile % group_items;
            const int nt = gt % P.n_tiles;
            const int mi = gt / P.n_tiles;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < m_sub; ++sub) {
                int mt = mi * m_sub + sub;
                const bool sub_ok = mt < m_tiles;      // odd tile count: the padding tile of the last work item
                const int tx = mt % P.tiles_x;
                mt /= P.tiles_x;
                const int ty = mt % P.tiles_y;
                const int tz = mt / P.tiles_y;
                RowCtx rc;
                if (P.flat) {
                    const int64_t qq = (int64_t)tx * P.box_x + row;
                    rc.z = (int)(qq / P.flat_img);
                    const int rem = (int)(qq % P.flat_img);
                    rc.y = rem / P.flat_w;
                    rc.x = rem % P.flat_w;
                } else {
                    rc.x = tx * P.box_x + row % P.box_x;
                    rc.y = ty * P.box_y + (row / P.box_x) % P.box_y;
                    rc.z = tz * P.box_z + row / (P.box_x * P.box_y);
                }
                rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N) && P.debug != 3;
                if (P.debug == 4) continue;
                rc.pix0 = o.off + P.group_off[grp] + (int64_t)rc.z * o.sn + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
                rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
                rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
                const uint32_t t_row = tmem_base + acc * set_cols + sub * col_stride + ((uint32_t)(q * 32) << 16);
                const int n0 = nt * P.block_n;
                const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
                const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (P.block_n >= group) &&
                                  (n0 + P.block_n <= o.C) && (P.debug != 5);
                StatCtx sc;
                if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
                if (fast) epilogue_tile_fast(P, t_row, n0, rc, stage_out + q * 8192, lane, o.stats != 0 ? sc.rb : 32);
                else epilogue_tile_generic(P, t_row, n0, rc);
                if (o.stats != 0) stats_tile_end(P, stage_out, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster (one TPC) compute a 256-pixel x block_n tile.  Each CTA stages
// its own 128-pixel A tile and HALF of the weight tile; the leader's single thread issues UMMA M=256, which reads A / B
// from both shared memories and writes each CTA's 128 accumulator rows into that CTA's TMEM.  Per SM and MMA this
// halves the B bytes read from shared memory (64 B/clk instead of 96 B/clk for N=256) - the 1-CTA kernel above is
// limited by exactly that operand bandwidth.  Synchronisation: TMA completions of both CTAs are counted on the leader's
// full barrier; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; both epilogues report
// "accumulator drained" to the leader.
// =====================================================================================================================
constexpr int kPairStages = 6;
constexpr int kPairStageBytes = kABytes + kBBytesMax / 2;                  // 16 KB A + up to 16 KB B half
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kStageOutBytes + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
igemm_pair_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_out = smem + kPairStages * kPairStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kStageOutBytes);
    uint64_t* full_bar = bars;                         // [kPairStages]  (the leader's are used)
    uint64_t* empty_bar = bars + kPairStages;          // [kPairStages]  (local, multicast-arrived)
    uint64_t* tfull_bar = bars + 2 * kPairStages;      // [2]            (local, multicast-arrived)
    uint64_t* tempty_bar = bars + 2 * kPairStages + 2; // [2]            (the leader's are used, 8 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int half_n = P.block_n / 2;
    const int col_stride = P.block_n < 32 ? 32 : P.block_n;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < P.planes; ++p) {
            for (int v = 0; v < ACLGAN_MAX_AVARIANTS; ++v) tma_prefetch_desc(&P.a[p][v]);
            tma_prefetch_desc(&P.b[p]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(&full_bar[s], 2);       // leader's expect_tx arrive + the peer's remote arrive
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 8);     // 4 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = P.tiles_x * P.tiles_y * P.tiles_z;
    const int m_items = (m_tiles + 1) / 2;
    const int group_items = m_items * P.n_tiles;
    const int total_items = group_items * P.n_groups;
    const int n_clusters = gridDim.x / 2;
    const int cluster_id = blockIdx.x / 2;
    const int k_iters = P.nseg * P.group_taps * P.cchunks;
    const uint32_t cta_tx = kABytes + half_n * 128;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int item = cluster_id; item < total_items; item += n_clusters) {
            const int grp = item / group_items;
            const int gi = item % group_items;
            const int tap0 = grp * P.group_taps;
            const int nt = gi % P.n_tiles;
            int mt = (gi / P.n_tiles) * 2 + (int)rank;   // past-the-end tile: decodes to z >= N (zero fill, never stored)
            const int x0 = (mt % P.tiles_x) * P.box_x;
            mt /= P.tiles_x;
            const int y0 = (mt % P.tiles_y) * P.box_y;
            const int z0 = (mt / P.tiles_y) * P.box_z;
#pragma unroll 1
            for (int seg = 0; seg < P.nseg; ++seg) {
                const int pa = (seg == 2) ? 1 : 0;
                const int pb = (seg == 1) ? 1 : 0;
#pragma unroll 1
                for (int t = tap0; t < tap0 + P.group_taps; ++t) {
                    const CUtensorMap* am = &P.a[pa][P.tap_var[t]];
                    const int dx = P.tap_dx[t], dy = P.tap_dy[t];
                    const int bk = P.tap_bk[t];
#pragma unroll 1
                    for (int cc = 0; cc < P.cchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * kPairStageBytes;
                        uint8_t* sb = sa + kABytes;
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * cta_tx);
                        else mbar_arrive_leader(&full_bar[stage]);
                        tma_load_4d_pair(sa, am, &full_bar[stage], cc * 64, x0 + dx, y0 + dy, z0);
                        tma_load_2d_pair(sb, &P.b[pb], &full_bar[stage], bk + cc * 64, nt * P.block_n + (int)rank * half_n);
                        if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        const uint32_t idesc = make_idesc_bf16(256, (uint32_t)P.block_n, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int item = cluster_id; item < total_items; item += n_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * col_stride;
#pragma unroll 1
            for (int k = 0; k < k_iters; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
                const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_pair(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                umma_commit_pair(&empty_bar[stage], 3);
                if (++stage == kPairStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(&tfull_bar[acc], 3);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (both CTAs, own 128 rows) ----------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        int it = 0;
        for (int item = cluster_id; item < total_items; item += n_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int grp = item / group_items;
            const int gi = item % group_items;
            const int nt = gi % P.n_tiles;
            int mt = (gi / P.n_tiles) * 2 + (int)rank;
            const bool sub_ok = mt < m_tiles;
            const int tx = mt % P.tiles_x;
            mt /= P.tiles_x;
            const int ty = mt % P.tiles_y;
            const int tz = mt / P.tiles_y;
            RowCtx rc;
            if (P.flat) {
                const int64_t qq = (int64_t)tx * P.box_x + row;
                rc.z = (int)(qq / P.flat_img);
                const int rem = (int)(qq % P.flat_img);
                rc.y = rem / P.flat_w;
                rc.x = rem % P.flat_w;
            } else {
                rc.x = tx * P.box_x + row % P.box_x;
                rc.y = ty * P.box_y + (row / P.box_x) % P.box_y;
                rc.z = tz * P.box_z + row / (P.box_x * P.box_y);
            }
            rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N) && P.debug != 3;
            rc.pix0 = o.off + P.group_off[grp] + (int64_t)rc.z * o.sn + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
            rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
            rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (P.debug != 4) {
                const uint32_t t_row = tmem_base + acc * col_stride + ((uint32_t)(q * 32) << 16);
                const int n0 = nt * P.block_n;
                const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
                const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (P.block_n >= group) &&
                                  (n0 + P.block_n <= o.C);
                StatCtx sc;
                if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
                if (fast) epilogue_tile_fast(P, t_row, n0, rc, stage_out + q * 8192, lane, o.stats != 0 ? sc.rb : 32);
                else epilogue_tile_generic(P, t_row, n0, rc);
                if (o.stats != 0) stats_tile_end(P, stage_out, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    cluster_sync_all();          // the peer may still be arriving on / reading the leader's barriers and TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// =====================================================================================================================
// Segment kernel (stride-1 convolutions).  The plain kernels above fetch one 128-pixel A box per filter tap, i.e. every
// activation byte crosses L2 -> shared memory k*k times per 64-channel chunk; at ~42 B/clk/SM of L2 bandwidth that, not
// the tensor pipe, bounds them (131 FLOP per staged byte for a 256x256 CTA-pair tile against ~190 needed).  Here the
// A rows of the k taps of ONE filter row - consecutive pixels - are staged once as a segment of 128 + k - 1 pixels, and
// tap kw is issued as a UMMA whose A descriptor starts kw rows (kw * 128 B) into the segment: the 128B-swizzle pattern is
// a function of the shared-memory address bits, so a row-shifted window of a TMA-written block is a valid operand (the
// descriptor's matrix-base-offset field carries the phase of the unaligned start row).  Two TMA rings: A segments (one
// per filter row and chunk) and B weight tiles (one per tap).
// Templated on PAIR: cta_group::2 (two CTAs, 256-pixel tile, B split in halves) or one CTA.
// =====================================================================================================================
constexpr int kSegMaxA = 4;
constexpr int kSegMaxB = 12;
constexpr int kSegSmemBytes = 232448;   // the whole opt-in budget; the rings are sized from it by the host

__device__ __forceinline__ uint64_t with_base_offset(uint64_t desc, uint32_t rows) {
    return desc | (static_cast<uint64_t>(rows & 7u) << 49);
}

template <bool PAIR>
__device__ __forceinline__ void seg_kernel_body(const IgemmKParams& P, uint8_t* smem_raw) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + P.seg_na * P.seg_a_bytes;
    uint8_t* stage_out = b_ring + P.seg_nb * P.seg_b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kStageOutBytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kSegMaxA;
    uint64_t* b_full = a_empty + kSegMaxA;
    uint64_t* b_empty = b_full + kSegMaxB;
    uint64_t* tfull_bar = b_empty + kSegMaxB;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int b_rows = PAIR ? P.block_n / 2 : P.block_n;
    const int col_stride = P.block_n < 32 ? 32 : P.block_n;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < P.planes; ++p) {
            tma_prefetch_desc(&P.a_seg[p]);
            tma_prefetch_desc(&P.b[p]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSegMaxA; ++s) { mbar_init(&a_full[s], PAIR ? 2 : 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kSegMaxB; ++s) { mbar_init(&b_full[s], PAIR ? 2 : 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], PAIR ? 8 : 4); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = P.tiles_x * P.tiles_y * P.tiles_z;
    const int m_items = PAIR ? (m_tiles + 1) / 2 : m_tiles;
    const int total_items = m_items * P.n_tiles;
    const int n_workers = PAIR ? gridDim.x / 2 : gridDim.x;
    const int worker = PAIR ? blockIdx.x / 2 : blockIdx.x;
    const uint32_t a_tx = (uint32_t)P.seg_a_bytes * (PAIR ? 2u : 1u);
    const uint32_t b_tx = (uint32_t)b_rows * 128u * (PAIR ? 2u : 1u);

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        for (int item = worker; item < total_items; item += n_workers) {
            const int nt = item % P.n_tiles;
            int mt = (item / P.n_tiles) * (PAIR ? 2 : 1) + (int)rank;   // past-the-end tile: zero fill, never stored
            const int x0 = (mt % P.tiles_x) * P.box_x;
            mt /= P.tiles_x;
            const int y0 = (mt % P.tiles_y) * P.box_y;
            const int z0 = (mt / P.tiles_y) * P.box_z;
#pragma unroll 1
            for (int seg = 0; seg < P.nseg; ++seg) {
                const int pa = (seg == 2) ? 1 : 0;
                const int pb = (seg == 1) ? 1 : 0;
#pragma unroll 1
                for (int cc = 0; cc < P.cchunks; ++cc) {
#pragma unroll 1
                    for (int sg = 0; sg < P.num_segs; ++sg) {
                        mbar_wait(&a_empty[as], aph ^ 1);
                        if (!PAIR) mbar_arrive_expect_tx(&a_full[as], a_tx);
                        else if (leader) mbar_arrive_expect_tx(&a_full[as], a_tx);
                        else mbar_arrive_leader(&a_full[as]);
                        uint8_t* sa = a_ring + as * P.seg_a_bytes;
                        if (PAIR) tma_load_4d_pair(sa, &P.a_seg[pa], &a_full[as], cc * 64, x0 + P.seg_dx[sg], y0 + P.seg_dy[sg], z0);
                        else tma_load_4d(sa, &P.a_seg[pa], &a_full[as], cc * 64, x0 + P.seg_dx[sg], y0 + P.seg_dy[sg], z0);
                        if (++as == P.seg_na) { as = 0; aph ^= 1; }
#pragma unroll 1
                        for (int j = 0; j < P.seg_taps; ++j) {
                            const int t = sg * P.seg_taps + j;
                            mbar_wait(&b_empty[bs], bph ^ 1);
                            if (!PAIR) mbar_arrive_expect_tx(&b_full[bs], b_tx);
                            else if (leader) mbar_arrive_expect_tx(&b_full[bs], b_tx);
                            else mbar_arrive_leader(&b_full[bs]);
                            uint8_t* sb = b_ring + bs * P.seg_b_bytes;
                            const int brow = nt * P.block_n + (int)rank * b_rows;
                            if (PAIR) tma_load_2d_pair(sb, &P.b[pb], &b_full[bs], P.tap_bk[t] + cc * 64, brow);
                            else tma_load_2d(sb, &P.b[pb], &b_full[bs], P.tap_bk[t] + cc * 64, brow);
                            if (++bs == P.seg_nb) { bs = 0; bph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, (uint32_t)P.block_n, 0, 0);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        int it = 0;
        for (int item = worker; item < total_items; item += n_workers, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * col_stride;
            uint32_t first = 1;
            const int n_a = P.nseg * P.cchunks * P.num_segs;
#pragma unroll 1
            for (int ai = 0; ai < n_a; ++ai) {
                const int sg = ai % P.num_segs;
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                const uint32_t sa = smem_u32(a_ring + as * P.seg_a_bytes);
#pragma unroll 1
                for (int j = 0; j < P.seg_taps; ++j) {
                    const int row = P.tap_row[sg * P.seg_taps + j];
                    mbar_wait(&b_full[bs], bph);
                    tc_fence_after();
                    uint64_t da = make_smem_desc_sw128(sa + row * 128, 16, 1024);
                    if (P.seg_bo) da = with_base_offset(da, (uint32_t)row);
                    const uint64_t db = make_smem_desc_sw128(smem_u32(b_ring + bs * P.seg_b_bytes), 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (PAIR) umma_bf16_pair(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                        else umma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                    }
                    first = 0;
                    if (PAIR) umma_commit_pair(&b_empty[bs], 3); else umma_commit(&b_empty[bs]);
                    if (++bs == P.seg_nb) { bs = 0; bph ^= 1; }
                }
                if (PAIR) umma_commit_pair(&a_empty[as], 3); else umma_commit(&a_empty[as]);
                if (++as == P.seg_na) { as = 0; aph ^= 1; }
            }
            if (PAIR) umma_commit_pair(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (own 128 rows) ----------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        int it = 0;
        for (int item = worker; item < total_items; item += n_workers, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int nt = item % P.n_tiles;
            int mt = (item / P.n_tiles) * (PAIR ? 2 : 1) + (int)rank;
            const bool sub_ok = mt < m_tiles;
            const int tx = mt % P.tiles_x;
            mt /= P.tiles_x;
            const int ty = mt % P.tiles_y;
            const int tz = mt / P.tiles_y;
            RowCtx rc;
            if (P.flat) {
                const int64_t qq = (int64_t)tx * P.box_x + row;
                rc.z = (int)(qq / P.flat_img);
                const int rem = (int)(qq % P.flat_img);
                rc.y = rem / P.flat_w;
                rc.x = rem % P.flat_w;
            } else {
                rc.x = tx * P.box_x + row % P.box_x;
                rc.y = ty * P.box_y + (row / P.box_x) % P.box_y;
                rc.z = tz * P.box_z + row / (P.box_x * P.box_y);
            }
            rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N);
            rc.pix0 = o.off + (int64_t)rc.z * o.sn + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
            rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
            rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * col_stride + ((uint32_t)(q * 32) << 16);
            const int n0 = nt * P.block_n;
            const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
            const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (P.block_n >= group) && (n0 + P.block_n <= o.C);
            StatCtx sc;
            if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
            if (fast) epilogue_tile_fast(P, t_row, n0, rc, stage_out + q * 8192, lane, o.stats != 0 ? sc.rb : 32);
            else epilogue_tile_generic(P, t_row, n0, rc);
            if (o.stats != 0) stats_tile_end(P, stage_out, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(kThreads, 1) igemm_seg_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    seg_kernel_body<false>(P, smem_raw);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
igemm_seg_pair_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    seg_kernel_body<true>(P, smem_raw);
}

// statistics in the epilogue need the staged fast path for every tile and warp-uniform image indices
static bool stats_supported(const aclgan_igemm_plan* pl) {
    const aclgan_out_spec& o = pl->out;
    const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
    if (o.kind != ACLGAN_OUT_BF16 && o.kind != ACLGAN_OUT_F32) return false;
    if (o.sc != 1 || pl->block_n < group || pl->block_n > 256 || o.C % pl->block_n != 0) return false;
    if (pl->n_groups > 1) return false;
    if (pl->flat) return pl->flat_img >= 128;           // a 128-row tile then touches at most two images
    if (pl->box_z != 1 && (pl->box_x * pl->box_y) % 32 != 0) return false;
    return true;
}

static int fill_kparams(const aclgan_igemm_plan* pl, IgemmKParams* kp) {
    if (pl->out.stats != 0 && !stats_supported(pl)) return ACLGAN_ERR_UNSUPPORTED;
    if (pl->planes < 1 || pl->planes > 2 || (pl->nseg != 1 && pl->nseg != 3)) return ACLGAN_ERR_SHAPE;
    if (pl->nseg == 3 && pl->planes != 2) return ACLGAN_ERR_SHAPE;
    if (pl->block_n != 16 && pl->block_n != 32 && pl->block_n != 64 && pl->block_n != 128 && pl->block_n != 256)
        return ACLGAN_ERR_SHAPE;
    if (pl->box_x * pl->box_y * pl->box_z != kTileM) return ACLGAN_ERR_SHAPE;
    if (pl->num_taps < 1 || pl->num_taps > ACLGAN_MAX_TAPS || pl->cchunks < 1) return ACLGAN_ERR_SHAPE;
    if (pl->n_avariants < 1 || pl->n_avariants > ACLGAN_MAX_AVARIANTS) return ACLGAN_ERR_SHAPE;
    for (int p = 0; p < pl->planes; ++p) {
        for (int v = 0; v < ACLGAN_MAX_AVARIANTS; ++v) {
            const aclgan_tmap_spec* s = &pl->a[p][v < pl->n_avariants ? v : 0];
            int rc = encode_tmap(s, &kp->a[p][v]);
            if (rc) return rc;
        }
        int rc = encode_tmap(&pl->b[p], &kp->b[p]);
        if (rc) return rc;
    }
    if (pl->planes == 1) {
        for (int v = 0; v < ACLGAN_MAX_AVARIANTS; ++v) kp->a[1][v] = kp->a[0][v];
        kp->b[1] = kp->b[0];
    }
    kp->planes = pl->planes; kp->nseg = pl->nseg; kp->block_n = pl->block_n; kp->n_tiles = pl->n_tiles;
    {
        // two 128-pixel tiles per work item halve the weight-tile (B) traffic per FLOP; only worth it when
        // enough work items remain to keep the SMs busy
        const int m_tiles = pl->tiles_x * pl->tiles_y * pl->tiles_z;
        const char* env = getenv("ACLGAN_IGEMM_MSUB");
        int m_sub = (m_tiles * pl->n_tiles * (pl->n_groups > 1 ? pl->n_groups : 1) >= 2 * num_sms()) ? 2 : 1;
        if (env != nullptr) m_sub = atoi(env) == 2 ? 2 : 1;
        kp->m_sub = m_sub;
        const char* dbg = getenv("ACLGAN_IGEMM_DEBUG");
        kp->debug = dbg != nullptr ? atoi(dbg) : 0;
    }
    kp->box_x = pl->box_x; kp->box_y = pl->box_y; kp->box_z = pl->box_z;
    kp->tiles_x = pl->tiles_x; kp->tiles_y = pl->tiles_y; kp->tiles_z = pl->tiles_z;
    kp->cchunks = pl->cchunks; kp->num_taps = pl->num_taps;
    kp->flat = pl->flat; kp->flat_w = pl->flat_w > 0 ? pl->flat_w : 1; kp->flat_img = pl->flat_img > 0 ? pl->flat_img : 1;
    kp->n_groups = pl->n_groups > 0 ? pl->n_groups : 1;
    kp->group_taps = pl->n_groups > 1 ? pl->group_taps : pl->num_taps;
    if (kp->n_groups > 4 || kp->n_groups * kp->group_taps != pl->num_taps) return ACLGAN_ERR_SHAPE;
    for (int g = 0; g < 4; ++g) kp->group_off[g] = pl->n_groups > 1 ? pl->group_off[g] : 0;
    for (int t = 0; t < ACLGAN_MAX_TAPS; ++t) {
        kp->tap_dx[t] = pl->tap_dx[t]; kp->tap_dy[t] = pl->tap_dy[t];
        kp->tap_var[t] = pl->tap_var[t]; kp->tap_bk[t] = pl->tap_bk[t];
        if (t < pl->num_taps && (pl->tap_var[t] < 0 || pl->tap_var[t] >= pl->n_avariants)) return ACLGAN_ERR_SHAPE;
    }
    kp->out = pl->out;
    kp->seg_rows = pl->seg_rows; kp->num_segs = pl->num_segs; kp->seg_taps = pl->seg_taps;
    for (int i = 0; i < 16; ++i) { kp->seg_dx[i] = pl->seg_dx[i]; kp->seg_dy[i] = pl->seg_dy[i]; }
    for (int t = 0; t < ACLGAN_MAX_TAPS; ++t) kp->tap_row[t] = pl->tap_row[t];
    kp->seg_a_bytes = kp->seg_b_bytes = kp->seg_na = kp->seg_nb = 0;
    {
        const char* bo = getenv("ACLGAN_SEG_BO");
        kp->seg_bo = bo != nullptr ? atoi(bo) : 1;
    }
    return ACLGAN_OK;
}

// segment-mode launch (returns -100 when the plan is not eligible and the plain kernels must run it)
static int launch_seg(const aclgan_igemm_plan* plan, IgemmKParams& kp, int repeat, cudaStream_t stream) {
    const char* env = getenv("ACLGAN_SEGK");
    if (!plan->seg_mode || (env != nullptr && atoi(env) == 0)) return -100;
    if (plan->n_groups > 1 || plan->seg_rows % 8 != 0 || plan->seg_rows > 256 || plan->num_segs < 1 || plan->num_segs > 16 ||
        plan->num_segs * plan->seg_taps != plan->num_taps)
        return -100;
    const int m_tiles = plan->tiles_x * plan->tiles_y * plan->tiles_z;
    const char* penv = getenv("ACLGAN_IGEMM_PAIR");
    bool pair = (plan->block_n >= 32) && (((m_tiles + 1) / 2) * plan->n_tiles >= num_sms() / 2);
    if (penv != nullptr) pair = atoi(penv) != 0 && plan->block_n >= 32;
    const int b_rows = pair ? plan->block_n / 2 : plan->block_n;
    for (int p = 0; p < plan->planes; ++p) {
        int rc = encode_tmap(&plan->a_seg[p], &kp.a_seg[p]);
        if (rc) return rc;
        aclgan_tmap_spec bs = plan->b[p];
        bs.box[1] = b_rows;
        rc = encode_tmap(&bs, &kp.b[p]);
        if (rc) return rc;
    }
    if (plan->planes == 1) { kp.a_seg[1] = kp.a_seg[0]; kp.b[1] = kp.b[0]; }
    kp.seg_a_bytes = plan->seg_rows * 128;
    kp.seg_b_bytes = ((b_rows * 128 + 1023) / 1024) * 1024;
    const int budget = kSegSmemBytes - 1024 - kStageOutBytes - 512;
    kp.seg_na = 3;
    int nb = (budget - kp.seg_na * kp.seg_a_bytes) / kp.seg_b_bytes;
    if (nb > kSegMaxB) nb = kSegMaxB;
    if (nb < 2) return -100;
    kp.seg_nb = nb;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(igemm_seg_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (pair) {
        const int items = ((m_tiles + 1) / 2) * plan->n_tiles;
        int clusters = num_sms() / 2;
        if (items < clusters) clusters = items;
        for (int i = 0; i < repeat; ++i) igemm_seg_pair_kernel<<<2 * clusters, kThreads, kSegSmemBytes, stream>>>(kp);
    } else {
        const int items = m_tiles * plan->n_tiles;
        if (items <= 0) return ACLGAN_OK;
        const int grid = items < num_sms() ? items : num_sms();
        for (int i = 0; i < repeat; ++i) igemm_seg_kernel<<<grid, kThreads, kSegSmemBytes, stream>>>(kp);
    }
    return (int)cudaGetLastError();
}

}  // namespace aclgan

extern "C" int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream);

extern "C" int aclgan_igemm_stats_supported(const aclgan_igemm_plan* plan) { return aclgan::stats_supported(plan) ? 1 : 0; }

extern "C" int aclgan_igemm_launch(const aclgan_igemm_plan* plan, void* stream) {
    return aclgan_igemm_launch_repeat(plan, 1, stream);
}

// launches the same plan `repeat` times back to back (tensor maps encoded once): device-side timing of the kernel
// without per-launch host work (bench.py roofline leg, tools/triage_igemm.py)
extern "C" int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream) {
    using namespace aclgan;
    static bool attr_set = false;
    IgemmKParams kp;
    int rc = fill_kparams(plan, &kp);
    if (rc) return rc;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    {
        const int src = launch_seg(plan, kp, repeat, (cudaStream_t)stream);
        if (src != -100) return src;
    }
    const int m_tiles_all = plan->tiles_x * plan->tiles_y * plan->tiles_z;
    {
        // CTA pairs when there is enough work to fill the SMs pairwise (env ACLGAN_IGEMM_PAIR=0|1 overrides)
        const char* env = getenv("ACLGAN_IGEMM_PAIR");
        const int n_groups = plan->n_groups > 1 ? plan->n_groups : 1;
        bool pair = (plan->block_n >= 32) && (((m_tiles_all + 1) / 2) * plan->n_tiles * n_groups >= num_sms() / 2);
        if (env != nullptr) pair = atoi(env) != 0 && plan->block_n >= 32;
        if (pair) {
            static bool pair_attr = false;
            // the pair kernel loads half of the weight tile per CTA: B boxes of block_n / 2 rows
            for (int p = 0; p < plan->planes; ++p) {
                aclgan_tmap_spec bs = plan->b[p];
                bs.box[1] = plan->block_n / 2;
                int rc2 = encode_tmap(&bs, &kp.b[p]);
                if (rc2) return rc2;
            }
            if (plan->planes == 1) kp.b[1] = kp.b[0];
            if (!pair_attr) {
                cudaError_t e = cudaFuncSetAttribute(igemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kPairSmemBytes);
                if (e != cudaSuccess) return (int)e;
                pair_attr = true;
            }
            const int items = ((m_tiles_all + 1) / 2) * plan->n_tiles * n_groups;
            int clusters = num_sms() / 2;
            if (items < clusters) clusters = items;
            for (int i = 0; i < repeat; ++i)
                igemm_pair_kernel<<<2 * clusters, kThreads, kPairSmemBytes, (cudaStream_t)stream>>>(kp);
            return (int)cudaGetLastError();
        }
    }
    const int total = ((m_tiles_all + kp.m_sub - 1) / kp.m_sub) * plan->n_tiles * kp.n_groups;
    if (total <= 0) return ACLGAN_OK;
    const int grid = total < num_sms() ? total : num_sms();
    for (int i = 0; i < repeat; ++i) igemm_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(kp);
    return (int)cudaGetLastError();
}
