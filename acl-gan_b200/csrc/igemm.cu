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

__device__ __noinline__ void epilogue_tile_fast(const IgemmKParams& P, uint32_t t_row, int n0, const RowCtx& rc, uint8_t* stg,
                                                int lane) {
    const aclgan_out_spec& o = P.out;
    const bool f32 = (o.kind == ACLGAN_OUT_F32);
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
        if (rc.valid) {
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
                                  (n0 + P.block_n <= o.C) && (o.stats == 0) && (P.debug != 5);
                if (fast) epilogue_tile_fast(P, t_row, n0, rc, stage_out + q * 8192, lane);
                else epilogue_tile_generic(P, t_row, n0, rc);
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
                                  (n0 + P.block_n <= o.C) && (o.stats == 0);
                if (fast) epilogue_tile_fast(P, t_row, n0, rc, stage_out + q * 8192, lane);
                else epilogue_tile_generic(P, t_row, n0, rc);
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

static int fill_kparams(const aclgan_igemm_plan* pl, IgemmKParams* kp) {
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
    return ACLGAN_OK;
}

}  // namespace aclgan

extern "C" int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream);

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
