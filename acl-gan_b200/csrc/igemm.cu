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
constexpr int kEpiThreads = 384;       // warps 0-3: TMA / MMA / TMEM roles, 4-7 and 8-11: two epilogue warp groups

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
    CUtensorMap b_seg[2];                           // rank 3 (k element, weight row, tap): box = 64 x rows x seg_taps
    int seg_rows, num_segs, seg_taps;
    int seg_a_bytes, seg_btile_bytes, seg_stage_bytes, seg_ns;   // shared-memory ring geometry chosen by the host
    int seg_msub;   // 128-pixel tiles per CTA and work item (1 | 2): with 2 every staged weight tile feeds two accumulators
    int seg_egroups; // epilogue warp groups (1 | 2): with 2, warps 4-7 drain the first half of a tile's accumulator columns and warps
                     // 8-11 the second half - the epilogue of a 2-tiles-per-CTA launch is latency-bound and exposed
    int seg_esplit;  // two groups: 1 = each drains half of the columns of every tile (N >= 128), 0 = the groups alternate tiles (group g
                     // owns accumulator set g: N < 128, where a tile's epilogue is longer than its few MMAs)
    int seg_bres;   // 1: weights resident - the seg_taps (x planes) weight tiles are staged ONCE per CTA (one N tile, one chunk,
                    // one segment per tile: the small-channel window layers), the ring then carries A segments only
    int epi_direct;
    long long* prof;                                // perf triage: per-CTA role timers [cta][8] (clock cycles) or null
    int seg_dx[16], seg_dy[16];
    int tap_row[ACLGAN_MAX_TAPS];
    aclgan_out_spec out;
};

// What the epilogue needs, BY VALUE: the non-inlined epilogue functions must not dereference the kernel parameter
// struct through a pointer (that turns every field access into a generic load from the constant window - hundreds of
// cycles each, in the middle of the unrolled conversion loops)
struct EpiArgs {
    uint64_t ptr0, ptr1, bias, stats;
    int64_t sy, sx, sc;
    int kind, act, mirror, N, C, bias_n, block_n;
    float slope;
    int direct;      // 1: registers -> global without shared-memory staging / shuffles (perf triage, env ACLGAN_EPI_DIRECT)
    int d2s_c, z_mod, stats_c;      // depth-to-space / strip addressing of the sub-pixel up-convolution (aclgan_out_spec)
    int64_t d2s_sy, d2s_sx;
};

// element offset of GEMM column `ch` relative to the pixel: plain channel stride, or depth-to-space (phase -> pixel offset)
__device__ __forceinline__ int64_t chan_off(const EpiArgs& o, int ch) {
    if (o.d2s_c == 0) return (int64_t)ch * o.sc;
    const int ph = ch / o.d2s_c, c = ch - ph * o.d2s_c;
    return (int64_t)(ph >> 1) * o.d2s_sy + (int64_t)(ph & 1) * o.d2s_sx + c;
}

__device__ __forceinline__ EpiArgs make_epi_args(const IgemmKParams& P) {
    EpiArgs e;
    e.ptr0 = P.out.ptr[0]; e.ptr1 = P.out.ptr[1]; e.bias = P.out.bias; e.stats = P.out.stats;
    e.sy = P.out.sy; e.sx = P.out.sx; e.sc = P.out.sc;
    e.kind = P.out.kind; e.act = P.out.act; e.mirror = P.out.mirror; e.N = P.out.N; e.C = P.out.C;
    e.bias_n = P.out.bias_n; e.block_n = P.block_n; e.slope = P.out.slope;
    e.direct = P.epi_direct;
    e.d2s_c = P.out.d2s_c; e.z_mod = P.out.z_mod; e.stats_c = P.out.stats_c > 0 ? P.out.stats_c : P.out.C;
    e.d2s_sy = P.out.d2s_sy; e.d2s_sx = P.out.d2s_sx;
    return e;
}

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

// sub-pixel up-convolution helpers (aclgan_out_spec.ring / z_mod): border-ring rows are left to the strip convolutions; a strip
// launch holds the two sides of the ring as batch indices [0, z_mod) and [z_mod, 2 z_mod) of the same images
__device__ __forceinline__ bool ring_pixel(const aclgan_out_spec& o, int x, int y) {
    return o.ring != 0 && (x == 0 || y == 0 || x == o.W - 1 || y == o.H - 1);
}
__device__ __forceinline__ int64_t image_off(const aclgan_out_spec& o, int z) {
    if (o.z_mod <= 0) return (int64_t)z * o.sn;
    return (int64_t)(z % o.z_mod) * o.sn + (int64_t)(z / o.z_mod) * o.z_off;
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
__device__ __noinline__ void store_generic(const EpiArgs o, int64_t pix, int ch0, int cnt, const float* v) {
    if (o.kind == ACLGAN_OUT_BF16 || o.kind == ACLGAN_OUT_SPLIT) {
        __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(o.ptr0) + pix;
        __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(o.ptr1) + pix;
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const int64_t co = chan_off(o, ch0 + i);
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[i]);
            d0[co] = hi;
            if (o.kind == ACLGAN_OUT_SPLIT) d1[co] = __float2bfloat16_rn(v[i] - __bfloat162float(hi));
        }
    } else {
        float* d = reinterpret_cast<float*>(o.ptr0) + pix;
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const int64_t co = chan_off(o, ch0 + i);
            if (o.kind == ACLGAN_OUT_F32_ATOMIC) atomicAdd(d + co, v[i]);
            else d[co] = v[i];
        }
    }
}

__device__ __noinline__ void epilogue_tile_generic(const EpiArgs o, uint32_t t_row, int n0, const RowCtx& rc) {
    const float* bias = reinterpret_cast<const float*>(o.bias);
    const int step = o.block_n >= 32 ? 32 : 16;
#pragma unroll 1
    for (int c = 0; c < o.block_n; c += step) {
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
        if (step == 16 && cnt == 16 && o.sc == 1 && o.d2s_c == 0 && o.kind == ACLGAN_OUT_F32 && o.mirror == 0 &&
            ((rc.pix0 + ch0) & 3) == 0) {
            // hot small-N case (16-column fp32 tiles: the image gradients of the first-layer data gradients, ~1 M rows of
            // 64 B per launch): everything in registers, four 16-byte stores per row
            float w16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float t = __uint_as_float(raw[i]);
                if (bias != nullptr && ch0 + i < o.bias_n) t += __ldg(bias + ch0 + i);
                w16[i] = apply_act(t, o.act, o.slope);
            }
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(o.ptr0) + rc.pix0 + ch0);
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = make_float4(w16[4 * i], w16[4 * i + 1], w16[4 * i + 2], w16[4 * i + 3]);
            continue;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            float t = v[i];
            if (bias != nullptr && ch0 + i < o.bias_n) t += __ldg(bias + ch0 + i);
            v[i] = apply_act(t, o.act, o.slope);
        }
        // full, channel-contiguous 16-wide chunks (e.g. the 16-column image-gradient tiles of the first-layer data gradients:
        // 1 M rows x 64 B) are stored as 16-byte vectors - the scalar loop below made those launches store-bound
        const bool vec = o.sc == 1 && o.d2s_c == 0 && cnt == 16 && step == 16 && o.kind != ACLGAN_OUT_F32_ATOMIC;
#pragma unroll 1
        for (int iy = 0; iy < rc.ny; ++iy)
#pragma unroll 1
            for (int ix = 0; ix < rc.nx; ++ix) {
                const int64_t pix = rc.pix0 + (int64_t)(rc.ys[iy] - rc.y) * o.sy + (int64_t)(rc.xs[ix] - rc.x) * o.sx;
                if (vec && o.kind == ACLGAN_OUT_F32 && ((pix + ch0) & 3) == 0) {
                    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(o.ptr0) + pix + ch0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else if (vec && o.kind != ACLGAN_OUT_F32 && ((pix + ch0) & 7) == 0) {
                    uint4* d0 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(o.ptr0) + pix + ch0);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        d0[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
                    if (o.kind == ACLGAN_OUT_SPLIT) {
                        uint4* d1 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(o.ptr1) + pix + ch0);
                        float r[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            d1[i] = make_uint4(pack_bf16x2(r[8 * i], r[8 * i + 1]), pack_bf16x2(r[8 * i + 2], r[8 * i + 3]),
                                               pack_bf16x2(r[8 * i + 4], r[8 * i + 5]), pack_bf16x2(r[8 * i + 6], r[8 * i + 7]));
                    }
                } else {
                    store_generic(o, pix, ch0, cnt, v);
                }
            }
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
__device__ __noinline__ void flush_rows(const int64_t o_sy, const int64_t o_sx, const uint8_t* stg, uint8_t* base, int64_t row_byte_off,
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
                               ((int64_t)(rc.ys[iy] - rc.y) * o_sy + (int64_t)(rc.xs[ix] - rc.x) * o_sx) * esz;
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

// named barrier of one 128-thread epilogue group (immediate ids, so that ptxas reserves 3 barriers, not all 16)
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    (void)threads;
    if (id == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// end of a tile with statistics: the four epilogue warps' slots -> fp64 atomics into stats[n][C][2].
// combine != 0: the tile's rows belong to image n_first (set 0) and, when `straddle`, n_first + 1 (set 1); the four warps
// are summed first (4x fewer atomics).  combine == 0 (tiny planes, several images per tile): every warp's 32 rows
// belong to the single image n_first (the host guarantees box_x * box_y % 32 == 0) and the warp adds its own slots.
// red0: reduction slots of the group's warp 0 (warp w: red0 + w * 8192); bar_id: named barrier of the (128-thread) group
__device__ __forceinline__ void stats_tile_end(const EpiArgs& o, const uint8_t* red0, int bar_id, int q, int lane, int n0, bool combine,
                                               int n_first, bool straddle, bool tile_ok) {
    double* stats = reinterpret_cast<double*>(o.stats);
    if (combine) {
        named_bar_sync(bar_id, 128);
        if (tile_ok) {
            for (int set = 0; set < (straddle ? 2 : 1); ++set) {
                const int n = n_first + set;
                if (n >= o.N) break;
                for (int ch = q * 32 + lane; ch < o.block_n; ch += 128) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float2 v = *reinterpret_cast<const float2*>(red0 + w * 8192 + set * 2048 + ch * 8);
                        s += v.x; qq += v.y;
                    }
                    double* d = stats + ((int64_t)(o.z_mod > 0 ? n % o.z_mod : n) * o.stats_c + n0 + ch) * 2;
                    atomicAdd(d, (double)s);
                    atomicAdd(d + 1, (double)qq);
                }
            }
        }
        named_bar_sync(bar_id, 128);
    } else {
        __syncwarp();
        if (tile_ok && n_first < o.N) {
            const float* red = reinterpret_cast<const float*>(red0 + q * 8192);
            for (int ch = lane; ch < o.block_n; ch += 32) {
                double* d = stats + ((int64_t)(o.z_mod > 0 ? n_first % o.z_mod : n_first) * o.stats_c + n0 + ch) * 2;
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

// Column sums over the warp's 32 rows of a 32-column register chunk, without shared memory: transpose-reduce butterfly.  In
// step k a lane keeps the half of its remaining columns that its lane bit selects and receives the partner's partial sums of
// that half (16 + 8 + 4 + 2 + 1 = 31 shuffles); lane l ends with the sum of column l.
__device__ __forceinline__ float warp_colsum32(float (&a)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? a[i] : a[i + off];
            const float keep = up ? a[i + off] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return a[0];
}

// statistics of one register chunk (bf16 output, values exactly as stored): per-column sum / sum of squares over the rows of
// image set 0 (lanes < rb) and set 1 (lanes >= rb) -> the warp's reduction slots red[set][column][2]
__device__ __forceinline__ void chunk_stats_regs(const float (&v)[32], bool row_ok, int lane, int rb, float* red, int col0) {
    float xr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) xr[i] = row_ok ? __bfloat162float(__float2bfloat16_rn(v[i])) : 0.f;
#pragma unroll 1
    for (int set = 0; set < 2; ++set) {
        const bool any = set == 0 ? rb > 0 : rb < 32;            // warp-uniform
        float s = 0.f, q = 0.f;
        if (any) {
            const bool mine = (rb == 32 || rb == 0) ? true : (set == 0 ? lane < rb : lane >= rb);
            float a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = mine ? xr[i] : 0.f;
            s = warp_colsum32(a, lane);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = mine ? xr[i] * xr[i] : 0.f;
            q = warp_colsum32(a, lane);
        }
        *reinterpret_cast<float2*>(red + set * 512 + (col0 + lane) * 2) = make_float2(s, q);
    }
}

// one 32-column chunk of the accumulator row: bias, activation, conversion, staging, and - once a full 128-byte row
// segment is staged - statistics and the global stores
__device__ __forceinline__ void epilogue_chunk(const EpiArgs& o, int c, const uint32_t (&raw)[32], float bl, int n0,
                                               const RowCtx& rc, uint8_t* stg, int lane, int stat_rb) {
    const bool f32 = (o.kind == ACLGAN_OUT_F32);
    const bool st = o.stats != 0;
    const int esz = f32 ? 4 : 2;
    const int ch0 = n0 + c;
    float v[32];
    if (o.direct && o.kind == ACLGAN_OUT_BF16) {
        // no shared-memory staging: broadcast bias loads, each lane stores its own row's 64 bytes; the fused statistics are
        // reduced across the warp's rows with a shuffle butterfly straight from the registers
        const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(o.bias) + ch0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o.bias != 0 && ch0 + 4 * i + 3 < o.bias_n) b4 = __ldg(bp + i);
            v[4 * i] = __uint_as_float(raw[4 * i]) + b4.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b4.y;
            v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b4.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b4.w;
        }
        if (o.act == ACLGAN_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (o.act == ACLGAN_ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * o.slope;
        }
        if (rc.valid) {
            uint4 qv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                qv[i].x = pack_bf16x2(v[8 * i], v[8 * i + 1]); qv[i].y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                qv[i].z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); qv[i].w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            }
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(o.ptr0) + rc.pix0 + chan_off(o, ch0);
            uint4* dst = reinterpret_cast<uint4*>(base);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = qv[i];
            if (rc.ny * rc.nx > 1) {
                // reflect-halo replicas of border pixels (few lanes): the same 64 bytes again at the mirrored coordinates
#pragma unroll 1
                for (int iy = 0; iy < rc.ny; ++iy)
#pragma unroll 1
                    for (int ix = (iy == 0 ? 1 : 0); ix < rc.nx; ++ix) {
                        uint4* d2 = reinterpret_cast<uint4*>(base + (int64_t)(rc.ys[iy] - rc.y) * o.sy + (int64_t)(rc.xs[ix] - rc.x) * o.sx);
#pragma unroll
                        for (int i = 0; i < 4; ++i) d2[i] = qv[i];
                    }
            }
        }
        if (st) chunk_stats_regs(v, rc.valid, lane, stat_rb, reinterpret_cast<float*>(stg + 4096), c);
        return;
    }
    // bias: every lane needs the same 32 values -> broadcast loads (one L1 transaction each; shuffles would go through the
    // shared-memory crossbar, which the tensor core's operand reads keep busy)
    if (o.bias != 0 && ch0 + 31 < o.bias_n) {
        const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(o.bias) + ch0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + i);
            v[4 * i] = __uint_as_float(raw[4 * i]) + b4.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b4.y;
            v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b4.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b4.w;
        }
    } else if (o.bias != 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + __shfl_sync(0xffffffffu, bl, i);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    }
    if (o.act == ACLGAN_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (o.act == ACLGAN_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * o.slope;
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
        const int64_t row_off = (rc.pix0 + chan_off(o, g0)) * esz;      // (fast path: sc == 1; a segment never straddles phases)
        if (st) {
            __syncwarp();
            stage_colsums(stg, reinterpret_cast<float*>(stg + 4096), g0 - n0, f32, lane, stat_rb);
        }
        flush_rows(o.sy, o.sx, stg, reinterpret_cast<uint8_t*>(o.ptr0), row_off, rc, esz, lane);
        if (o.kind == ACLGAN_OUT_SPLIT)
            flush_rows(o.sy, o.sx, stg + 4096, reinterpret_cast<uint8_t*>(o.ptr1), row_off, rc, esz, lane);
    }
}

// Software-pipelined over 32-column chunks: the TMEM load (and the bias load) of the next chunk is in flight while the
// current one is converted and stored - a single warp has no other source of latency hiding here.
__device__ __noinline__ void epilogue_tile_fast(const EpiArgs o, uint32_t t_row, int n0, const RowCtx& rc, uint8_t* stg,
                                                int lane, int stat_rb, long long* pt = nullptr) {
    const float* bias = reinterpret_cast<const float*>(o.bias);
    const long long pt0 = pt ? clock64() : 0;
    long long pt_ld = 0, pt_chunk = 0;
    // one copy of the chunk code (instruction-cache footprint: the epilogue is otherwise fetch-bound): the next chunk's
    // TMEM / bias loads are issued into `nxt`, moved to `cur` (32 register moves) when they have landed
    uint32_t cur[32], nxt[32];
    float bl_cur, bl_nxt = 0.f;
    tmem_ld_32x32(t_row, nxt);
    if (bias != nullptr && n0 + lane < o.bias_n) bl_nxt = __ldg(bias + n0 + lane);
#pragma unroll 1
    for (int c = 0; c < o.block_n; c += 32) {
        long long t1 = pt ? clock64() : 0;
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
        bl_cur = bl_nxt;
        if (c + 32 < o.block_n) {
            tmem_ld_32x32(t_row + c + 32, nxt);
            bl_nxt = (bias != nullptr && n0 + c + 32 + lane < o.bias_n) ? __ldg(bias + n0 + c + 32 + lane) : 0.f;
        }
        if (pt) { const long long t2 = clock64(); pt_ld += t2 - t1; t1 = t2; }
        epilogue_chunk(o, c, cur, bl_cur, n0, rc, stg, lane, stat_rb);
        if (pt) pt_chunk += clock64() - t1;
    }
    if (pt && lane == 0) { pt[0] += clock64() - pt0; pt[1] += pt_ld; pt[2] += pt_chunk; }
}

__global__ void __launch_bounds__(kEpiThreads, 1) igemm_kernel(const __grid_constant__ IgemmKParams P) {
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
            mbar_init(&tempty_bar[s], 4 * (P.seg_esplit ? P.seg_egroups : 1));
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
    } else if (warp >= 4 && ((warp - 4) >> 2) < P.seg_egroups) {
        // ---------------- epilogue (one or two warp groups, see IgemmKParams::seg_egroups) ----------------
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int eg = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        EpiArgs ea = make_epi_args(P);
        const bool alternate = P.seg_egroups > 1 && !P.seg_esplit;
        const int ecols = P.seg_esplit ? P.block_n / P.seg_egroups : P.block_n;
        const int ecol0 = P.seg_esplit ? eg * ecols : 0;
        ea.block_n = ecols;
        uint8_t* const stg = stage_out + q * 8192 - (eg != 0 ? 4096 : 0);
        const uint8_t* const red0 = stage_out + (eg != 0 ? 0 : 4096);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            if (alternate && (it & 1) != eg) continue;
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
                rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N) && P.debug != 3 && !ring_pixel(o, rc.x, rc.y);
                if (P.debug == 4) continue;
                rc.pix0 = o.off + P.group_off[grp] + image_off(o, rc.z) + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
                rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
                rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
                const uint32_t t_row = tmem_base + acc * set_cols + sub * col_stride + ecol0 + ((uint32_t)(q * 32) << 16);
                const int n0 = nt * P.block_n + ecol0;
                const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
                const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (ecols >= group) && (o.d2s_c % group == 0) &&
                                  (n0 + ecols <= o.C) && (o.act != ACLGAN_ACT_TANH) && (P.debug != 5);
                StatCtx sc;
                if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
                if (fast) epilogue_tile_fast(ea, t_row, n0, rc, stg, lane, o.stats != 0 ? sc.rb : 32);
                else epilogue_tile_generic(ea, t_row, n0, rc);
                if (o.stats != 0) stats_tile_end(ea, red0, 1 + eg, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEpiThreads, 1)
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
    if (warp == 2) {
        // barrier initialisation and TMEM allocation by the same warp, one after the other.  (compute-sanitizer's racecheck
        // reports the cta_group::2 allocation itself: the instruction of EITHER CTA of the pair deposits the allocated address in
        // both CTAs' shared memory, which the tool sees as an unordered remote write against the local one - both write the same
        // value, and every reader sits behind the cluster barrier below: profiles/r2_sanitize_racecheck_dim32.log)
        if (lane == 0) {
            for (int s = 0; s < kPairStages; ++s) {
                mbar_init(&full_bar[s], 2);       // leader's expect_tx arrive + the peer's remote arrive
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tfull_bar[s], 1);
                mbar_init(&tempty_bar[s], 8 * (P.seg_esplit ? P.seg_egroups : 1));     // 4 epilogue warps per group of each CTA
            }
            fence_barrier_init();
        }
        __syncwarp();
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
    } else if (warp >= 4 && ((warp - 4) >> 2) < P.seg_egroups) {
        // ---------------- epilogue (both CTAs, own 128 rows; one or two warp groups, see IgemmKParams::seg_egroups) ----------------
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        EpiArgs ea = make_epi_args(P);
        const bool alternate = P.seg_egroups > 1 && !P.seg_esplit;
        const int ecols = P.seg_esplit ? P.block_n / P.seg_egroups : P.block_n;
        const int ecol0 = P.seg_esplit ? eg * ecols : 0;
        ea.block_n = ecols;
        uint8_t* const stg = stage_out + q * 8192 - (eg != 0 ? 4096 : 0);
        const uint8_t* const red0 = stage_out + (eg != 0 ? 0 : 4096);
        int it = 0;
        for (int item = cluster_id; item < total_items; item += n_clusters, ++it) {
            if (alternate && (it & 1) != eg) continue;
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
            rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N) && P.debug != 3 && !ring_pixel(o, rc.x, rc.y);
            rc.pix0 = o.off + P.group_off[grp] + image_off(o, rc.z) + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
            rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
            rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (P.debug != 4) {
                const uint32_t t_row = tmem_base + acc * col_stride + ecol0 + ((uint32_t)(q * 32) << 16);
                const int n0 = nt * P.block_n + ecol0;
                const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
                const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (ecols >= group) && (o.d2s_c % group == 0) &&
                                  (n0 + ecols <= o.C);
                StatCtx sc;
                if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
                if (fast) epilogue_tile_fast(ea, t_row, n0, rc, stg, lane, o.stats != 0 ? sc.rb : 32);
                else epilogue_tile_generic(ea, t_row, n0, rc);
                if (o.stats != 0) stats_tile_end(ea, red0, 1 + eg, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
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
#ifndef ACLGAN_SEG_MAXREG
#define ACLGAN_SEG_MAXREG 168     // leaves registers for a co-resident element-wise CTA of another chain (see seg_smem_bytes)
#endif
constexpr int kSegMaxStages = 6;
constexpr int kSegThreads = 384;       // warps 0-3: TMA / MMA / TMEM roles, 4-7 and 8-11: two epilogue groups
constexpr int kSegSmemBytes = 232448;   // the whole opt-in budget; the ring is sized from it by the host

// One pipeline stage = one filter row of one 64-channel chunk: the A segment (seg_rows pixels) and the weight tiles of
// the row's k taps ([tap][b_rows][64], ONE rank-3 TMA box), i.e. two TMA instructions, one barrier round trip and 4*k
// MMAs per stage - the single-thread producer / issuer loops cost a few hundred cycles per round trip, which a stage of
// only 4 MMAs does not hide once N < 256.
// the 4 * K MMAs of one filter row, straight-line (descriptor arithmetic on compile-time tap / k-slice indices)
template <int K, bool PAIR>
__device__ __forceinline__ void issue_row(uint32_t d_tmem, uint64_t da0, uint64_t db0, int64_t rstep8, uint32_t bt16, uint32_t idesc,
                                          uint32_t accum) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const uint64_t da = da0 + (uint64_t)(j * rstep8);
        const uint64_t db = db0 + (uint64_t)j * bt16;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (j == 0 && kk == 0) ? accum : 1u;
            if (PAIR) umma_bf16_pair(d_tmem, da + 2 * kk, db + 2 * kk, idesc, acc);
            else umma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, acc);
        }
    }
}

template <bool PAIR>
__device__ __forceinline__ void seg_kernel_body(const IgemmKParams& P, uint8_t* smem_raw) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int bres_bytes = P.seg_bres ? P.planes * P.seg_taps * P.seg_btile_bytes : 0;
    uint8_t* bres = smem + P.seg_ns * P.seg_stage_bytes;        // resident weight tiles [plane][tap][b_rows][64] (seg_bres)
    uint8_t* stage_out = bres + bres_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kStageOutBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = full_bar + kSegMaxStages;
    uint64_t* tfull_bar = empty_bar + kSegMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint64_t* bres_bar = tempty_bar + 3;                        // resident weights landed (seg_bres)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int b_rows = PAIR ? P.block_n / 2 : P.block_n;
    const int col_stride = P.block_n < 32 ? 32 : P.block_n;
    const bool prof = P.prof != nullptr;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < P.planes; ++p) {
            tma_prefetch_desc(&P.a_seg[p]);
            tma_prefetch_desc(&P.b_seg[p]);
        }
    }
    if (warp == 2) {
        // (barrier initialisation and TMEM allocation by the same warp, one after the other: see igemm_pair_kernel)
        if (lane == 0) {
            for (int s = 0; s < kSegMaxStages; ++s) { mbar_init(&full_bar[s], PAIR ? 2 : 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], (PAIR ? 8 : 4) * (P.seg_esplit ? P.seg_egroups : 1)); }
            mbar_init(bres_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        if (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Work item = m_sub (1 | 2) M tiles per CTA x one N tile.  With m_sub = 2 the weight tiles of a stage (3/4 of the staged
    // bytes of the 3x3 256->256 layer, whose two-deep ring runs at the ~42 B/clk/SM L2->SM limit) feed two accumulators, and
    // the 137 tile pairs of a batch-8 res-block conv become 69 items = ONE wave on the 74 clusters instead of 1.85.
    const int m_sub = P.seg_msub;
    const int tile_mul = PAIR ? 2 : 1;
    const int m_tiles = P.tiles_x * P.tiles_y * P.tiles_z;
    const int m_items = (m_tiles + tile_mul * m_sub - 1) / (tile_mul * m_sub);
    const int total_items = m_items * P.n_tiles;
    const int n_workers = PAIR ? gridDim.x / 2 : gridDim.x;
    const int worker = PAIR ? blockIdx.x / 2 : blockIdx.x;
    const uint32_t stage_tx = (uint32_t)(m_sub * P.seg_a_bytes + (P.seg_bres ? 0 : P.seg_taps * P.seg_btile_bytes)) * (PAIR ? 2u : 1u);
    const int n_rounds = P.nseg * P.cchunks * P.num_segs;       // stages per work item
    const int set_cols = m_sub * col_stride;                    // TMEM columns of one accumulator set
    const int acc_sets = (2 * set_cols <= 512) ? 2 : 1;         // double-buffered across items when they fit

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int st = 0;
        uint32_t ph = 0;
        long long t_wait = 0;
        const long long t_begin = prof ? clock64() : 0;
        if (!PAIR && P.seg_bres && worker < total_items) {
            mbar_arrive_expect_tx(bres_bar, (uint32_t)bres_bytes);
            for (int p = 0; p < P.planes; ++p)
                tma_load_3d(bres + p * P.seg_taps * P.seg_btile_bytes, &P.b_seg[p], bres_bar, 0, 0, 0);
        }
        for (int item = worker; item < total_items; item += n_workers) {
            const int nt = item % P.n_tiles;
            int x0[2], y0[2], z0[2];
            for (int sub = 0; sub < m_sub; ++sub) {
                int mt = ((item / P.n_tiles) * m_sub + sub) * tile_mul + (int)rank;   // past-the-end tile: zero fill, never stored
                x0[sub] = (mt % P.tiles_x) * P.box_x;
                mt /= P.tiles_x;
                y0[sub] = (mt % P.tiles_y) * P.box_y;
                z0[sub] = (mt / P.tiles_y) * P.box_z;
            }
            const int brow = nt * P.block_n + (int)rank * b_rows;
#pragma unroll 1
            for (int seg = 0; seg < P.nseg; ++seg) {
                const CUtensorMap* am = &P.a_seg[(seg == 2) ? 1 : 0];
                const CUtensorMap* bm = &P.b_seg[(seg == 1) ? 1 : 0];
#pragma unroll 1
                for (int cc = 0; cc < P.cchunks; ++cc) {
#pragma unroll 1
                    for (int sg = 0; sg < P.num_segs; ++sg) {
                        const long long tw = prof ? clock64() : 0;
                        mbar_wait(&empty_bar[st], ph ^ 1);
                        if (prof) t_wait += clock64() - tw;
                        uint8_t* sa = smem + st * P.seg_stage_bytes;
                        uint8_t* sb = sa + m_sub * P.seg_a_bytes;
                        if (P.debug == 1) {          // MMA without TMA traffic
                            if (!PAIR || leader) mbar_arrive(&full_bar[st]); else mbar_arrive_leader(&full_bar[st]);
                        } else if (PAIR) {
                            if (leader) mbar_arrive_expect_tx(&full_bar[st], stage_tx); else mbar_arrive_leader(&full_bar[st]);
                            for (int sub = 0; sub < m_sub; ++sub)
                                tma_load_4d_pair(sa + sub * P.seg_a_bytes, am, &full_bar[st], cc * 64, x0[sub] + P.seg_dx[sg],
                                                 y0[sub] + P.seg_dy[sg], z0[sub]);
                            tma_load_3d_pair(sb, bm, &full_bar[st], cc * 64, brow, sg * P.seg_taps);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[st], stage_tx);
                            for (int sub = 0; sub < m_sub; ++sub)
                                tma_load_4d(sa + sub * P.seg_a_bytes, am, &full_bar[st], cc * 64, x0[sub] + P.seg_dx[sg],
                                            y0[sub] + P.seg_dy[sg], z0[sub]);
                            if (!P.seg_bres) tma_load_3d(sb, bm, &full_bar[st], cc * 64, brow, sg * P.seg_taps);
                        }
                        if (++st == P.seg_ns) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
        if (prof) { P.prof[blockIdx.x * 16 + 0] = clock64() - t_begin; P.prof[blockIdx.x * 16 + 1] = t_wait; }
    } else if (warp == 1 && lane == 0 && leader) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, (uint32_t)P.block_n, 0, 0);
        int st = 0;
        uint32_t ph = 0;
        int it = 0;
        long long t_wfull = 0, t_wacc = 0;
        const long long t_begin = prof ? clock64() : 0;
        // (the host checked that tap_row is the same arithmetic progression in every filter row)
        const int row0 = P.tap_row[0];
        const int64_t rstep8 = P.seg_taps > 1 ? (int64_t)(P.tap_row[1] - P.tap_row[0]) * 8 : 0;   // rows -> descriptor units
        const uint32_t bt16 = (uint32_t)P.seg_btile_bytes >> 4;
        if (!PAIR && P.seg_bres && worker < total_items) mbar_wait(bres_bar, 0);
        for (int item = worker; item < total_items; item += n_workers, ++it) {
            const int acc = it % acc_sets;
            const uint32_t acc_phase = (it / acc_sets) & 1;
            long long tw = prof ? clock64() : 0;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            if (prof) t_wacc += clock64() - tw;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * set_cols;
            uint32_t accum = 0;
#pragma unroll 1
            for (int r = 0; r < n_rounds; ++r) {
                const int sg = r % P.num_segs;
                tw = prof ? clock64() : 0;
                mbar_wait(&full_bar[st], ph);
                if (prof) t_wfull += clock64() - tw;
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + st * P.seg_stage_bytes);
                // (resident weights: plane of B follows the hi/lo segment: (a_hi, b_hi), (a_hi, b_lo), (a_lo, b_hi))
                const uint32_t sb = P.seg_bres ? smem_u32(bres) + (uint32_t)(((r / (P.cchunks * P.num_segs)) == 1 ? 1 : 0) * P.seg_taps * P.seg_btile_bytes)
                                               : sa + m_sub * P.seg_a_bytes;
                if (P.debug != 2) {          // (2 = TMA traffic without MMA)
                    // A descriptor of tap j = stage base + (row0 + j * rstep) rows; B descriptor = base + j weight tiles
                    const uint64_t db0 = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll 1
                    for (int sub = 0; sub < m_sub; ++sub) {
                        const uint64_t da0 = make_smem_desc_sw128(sa + sub * P.seg_a_bytes, 16, 1024) + (uint64_t)(row0 * 8);
                        const uint32_t dt = d_tmem + sub * col_stride;
                        switch (P.seg_taps) {
                            case 3: issue_row<3, PAIR>(dt, da0, db0, rstep8, bt16, idesc, accum); break;
                            case 5: issue_row<5, PAIR>(dt, da0, db0, rstep8, bt16, idesc, accum); break;
                            case 7: issue_row<7, PAIR>(dt, da0, db0, rstep8, bt16, idesc, accum); break;
                            default:
#pragma unroll 1
                                for (int j = 0; j < P.seg_taps; ++j) issue_row<1, PAIR>(dt, da0 + (int64_t)j * rstep8, db0 + (uint64_t)j * bt16, 0, 0, idesc, accum);
                        }
                    }
                    accum = 1;
                }
                if (PAIR) umma_commit_pair(&empty_bar[st], 3); else umma_commit(&empty_bar[st]);
                if (++st == P.seg_ns) { st = 0; ph ^= 1; }
            }
            if (PAIR) umma_commit_pair(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]);
        }
        if (prof) {
            P.prof[blockIdx.x * 16 + 2] = clock64() - t_begin; P.prof[blockIdx.x * 16 + 3] = t_wfull; P.prof[blockIdx.x * 16 + 4] = t_wacc;
        }
    } else if (warp >= 4 && ((warp - 4) >> 2) < P.seg_egroups) {
        // ---------------- epilogue (own 128 rows; with two groups: own half of the accumulator columns) ----------------
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const aclgan_out_spec& o = P.out;
        EpiArgs ea = make_epi_args(P);
        const bool alternate = P.seg_egroups > 1 && !P.seg_esplit;      // group g drains the items with it % 2 == g (accumulator set g)
        const int ecols = P.seg_esplit ? P.block_n / P.seg_egroups : P.block_n;      // columns drained by this group
        const int ecol0 = P.seg_esplit ? eg * ecols : 0;
        ea.block_n = ecols;
        // reduction slots: group 0 uses the second half of each warp's 8 KB staging area, group 1 (direct epilogue only: the
        // staging half is unused) the first half; epilogue_chunk finds them at stg + 4096
        uint8_t* const stg = stage_out + q * 8192 - (eg != 0 ? 4096 : 0);
        const uint8_t* const red0 = stage_out + (eg != 0 ? 0 : 4096);
        int it = 0;
        long long t_wt = 0;
        const long long t_begin = prof ? clock64() : 0;
        for (int item = worker; item < total_items; item += n_workers, ++it) {
            if (alternate && (it & 1) != eg) continue;
            const int acc = it % acc_sets;
            const uint32_t acc_phase = (it / acc_sets) & 1;
            const int nt = item % P.n_tiles;
            const long long tw = prof ? clock64() : 0;
            mbar_wait(&tfull_bar[acc], acc_phase);
            if (prof) t_wt += clock64() - tw;
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < m_sub; ++sub) {
                int mt = ((item / P.n_tiles) * m_sub + sub) * tile_mul + (int)rank;
                const bool sub_ok = mt < m_tiles;
                const int tx = mt % P.tiles_x;
                mt /= P.tiles_x;
                const int ty = mt % P.tiles_y;
                const int tz = mt / P.tiles_y;
                RowCtx rc;
                if (P.flat) {
                    const int64_t qq = (int64_t)tx * P.box_x + row;
                    rc.z = (int)(qq / P.flat_img);
                    const int rem = (int)(qq - (int64_t)rc.z * P.flat_img);
                    rc.y = rem / P.flat_w;
                    rc.x = rem - rc.y * P.flat_w;
                } else {
                    rc.x = tx * P.box_x + row % P.box_x;
                    rc.y = ty * P.box_y + (row / P.box_x) % P.box_y;
                    rc.z = tz * P.box_z + row / (P.box_x * P.box_y);
                }
                rc.valid = sub_ok && (rc.x < o.W) && (rc.y < o.H) && (rc.z < o.N) && !ring_pixel(o, rc.x, rc.y);
                rc.pix0 = o.off + image_off(o, rc.z) + (int64_t)rc.y * o.sy + (int64_t)rc.x * o.sx;
                rc.ny = mirror_coords(rc.y, o.H, o.mirror, rc.ys);
                rc.nx = mirror_coords(rc.x, o.W, o.mirror, rc.xs);
                const uint32_t t_row = tmem_base + acc * set_cols + sub * col_stride + ecol0 + ((uint32_t)(q * 32) << 16);
                const int n0 = nt * P.block_n + ecol0;
                const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
                const bool fast = (o.sc == 1) && (o.kind != ACLGAN_OUT_F32_ATOMIC) && (ecols >= group) && (o.d2s_c % group == 0) &&
                                  (n0 + ecols <= o.C) && (o.act != ACLGAN_ACT_TANH);
                StatCtx sc;
                if (o.stats != 0) sc = make_stat_ctx(P, tx, tz, q, lane, rc);
                const long long te = prof ? clock64() : 0;
                if (fast) epilogue_tile_fast(ea, t_row, n0, rc, stg, lane, o.stats != 0 ? sc.rb : 32,
                                             (prof && warp == 4) ? P.prof + blockIdx.x * 16 + 8 : nullptr);
                else epilogue_tile_generic(ea, t_row, n0, rc);
                if (o.stats != 0) stats_tile_end(ea, red0, 1 + eg, q, lane, n0, sc.combine, sc.n_first, sc.straddle, sub_ok);
                if (prof && warp == 4 && lane == 0) P.prof[blockIdx.x * 16 + 7] += clock64() - te;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
            }
        }
        if (prof && warp == 4 && lane == 0) {
            P.prof[blockIdx.x * 16 + 5] = clock64() - t_begin; P.prof[blockIdx.x * 16 + 6] = t_wt;
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __maxnreg__(ACLGAN_SEG_MAXREG) igemm_seg_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    seg_kernel_body<false>(P, smem_raw);
}

__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(ACLGAN_SEG_MAXREG)
igemm_seg_pair_kernel(const __grid_constant__ IgemmKParams P) {
    extern __shared__ uint8_t smem_raw[];
    seg_kernel_body<true>(P, smem_raw);
}

// statistics in the epilogue need the staged fast path for every tile and warp-uniform image indices
static bool stats_supported(const aclgan_igemm_plan* pl) {
    const aclgan_out_spec& o = pl->out;
    const int group = (o.kind == ACLGAN_OUT_F32) ? 32 : 64;
    if (o.kind != ACLGAN_OUT_BF16 && o.kind != ACLGAN_OUT_F32) return false;
    if (o.sc != 1 || pl->block_n < group || pl->block_n > 256 || o.C % pl->block_n != 0 || o.act == ACLGAN_ACT_TANH) return false;
    if (o.d2s_c % group != 0) return false;             // depth-to-space phases narrower than a staged row segment: generic path
    if (pl->n_groups > 1) return false;
    if (pl->flat) return pl->flat_img >= 128;           // a 128-row tile then touches at most two images
    if (pl->box_z != 1 && (pl->box_x * pl->box_y) % 32 != 0) return false;
    return true;
}

static long long* g_prof = nullptr;      // set by aclgan_igemm_set_prof (perf triage)

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
    kp->seg_a_bytes = kp->seg_btile_bytes = kp->seg_stage_bytes = kp->seg_ns = 0;
    kp->seg_msub = 1;
    kp->seg_bres = 0;
    kp->seg_egroups = 1;
    kp->seg_esplit = 0;
    kp->prof = g_prof;
    {
        const char* ed = getenv("ACLGAN_EPI_DIRECT");
        kp->epi_direct = ed != nullptr ? atoi(ed) : 1;
    }
    {
        // (measured on B200: the 128B swizzle of a UMMA operand is a function of the absolute shared-memory address
        // bits, so the row-shifted start addresses of the segment kernel need NO matrix base offset in the descriptor;
        // with the offset set to (addr >> 7) & 7 the results are wrong)
        const char* sd = getenv("ACLGAN_SEG_DEBUG");     // perf triage: 1 = every tap reads rows [0, 128) (aligned start)
        if (sd != nullptr && atoi(sd) == 1)
            for (int t = 0; t < ACLGAN_MAX_TAPS; ++t) kp->tap_row[t] = 0;
    }
    return ACLGAN_OK;
}

// Two epilogue warp groups whenever every tile of the launch takes the direct bf16 epilogue (no shared-memory staging: group 1
// has none): N >= 128: the groups split the accumulator columns in halves; N = 64: they alternate tiles (needs two accumulator
// sets).  acc_cols = TMEM columns of one accumulator set.  env ACLGAN_EPI_GROUPS=1 switches back to one group.
static void decide_egroups(const aclgan_igemm_plan* plan, IgemmKParams& kp, int acc_cols) {
    const char* e = getenv("ACLGAN_EPI_GROUPS");
    const aclgan_out_spec& o = plan->out;
    const bool ok = plan->block_n >= 64 && o.kind == ACLGAN_OUT_BF16 && kp.epi_direct && o.sc == 1 && o.C % plan->block_n == 0 &&
                    o.act != ACLGAN_ACT_TANH && o.d2s_c % 64 == 0 && kp.debug == 0;
    // narrow tiles (N = 16 | 32: image-gradient planes of the first layers) take the generic epilogue for every tile, which uses
    // no staging either: alternating groups there too
    const bool generic_all = plan->block_n < (o.kind == ACLGAN_OUT_F32 ? 32 : 64) && o.stats == 0 && kp.debug == 0;
    kp.seg_egroups = ((ok || generic_all) && (e == nullptr || atoi(e) != 1)) ? 2 : 1;
    kp.seg_esplit = (ok && plan->block_n >= 128) ? 1 : 0;
    if (kp.seg_egroups == 2 && !kp.seg_esplit && 2 * acc_cols > 512) kp.seg_egroups = 1;
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
    // CTA pairs halve the weight bytes staged per SM; below N = 128 the stage is small anyway and one CTA per tile
    // issues its (N-independent, ~70-cycle) MMAs at twice the rate per SM
    bool pair = (plan->block_n >= 128) && (((m_tiles + 1) / 2) * plan->n_tiles >= num_sms() / 2);
    if (penv != nullptr) pair = atoi(penv) != 0 && plan->block_n >= 32;
    const int b_rows = pair ? plan->block_n / 2 : plan->block_n;
    if (b_rows % 8 != 0) return -100;
    for (int p = 0; p < plan->planes; ++p) {
        int rc = encode_tmap(&plan->a_seg[p], &kp.a_seg[p]);
        if (rc) return rc;
        // weights of the seg_taps taps of one filter row as ONE box: (64 k-elements, b_rows rows, seg_taps taps)
        const aclgan_tmap_spec& b2 = plan->b[p];
        aclgan_tmap_spec b3;
        memset(&b3, 0, sizeof(b3));
        const uint64_t tap_elems = (uint64_t)plan->cchunks * 64;          // k elements per tap (tap_bk[t] = t * tap_elems)
        b3.base = b2.base; b3.rank = 3; b3.elem_bytes = 2;
        b3.dims[0] = tap_elems;          b3.strides[0] = 2;             b3.box[0] = 64;
        b3.dims[1] = b2.dims[1];         b3.strides[1] = b2.strides[1]; b3.box[1] = b_rows;
        b3.dims[2] = plan->num_taps;     b3.strides[2] = tap_elems * 2; b3.box[2] = plan->seg_taps;
        rc = encode_tmap(&b3, &kp.b_seg[p]);
        if (rc) return rc;
    }
    for (int t = 0; t < plan->num_taps; ++t) {
        if (plan->tap_bk[t] != t * plan->cchunks * 64) return -100;
        const int j = t % plan->seg_taps;
        const int step = plan->seg_taps > 1 ? plan->tap_row[1] - plan->tap_row[0] : 0;
        if (plan->tap_row[t] != plan->tap_row[0] + j * step || plan->tap_row[t] < 0 ||
            plan->tap_row[t] + 128 > plan->seg_rows)
            return -100;
    }
    if (plan->planes == 1) { kp.a_seg[1] = kp.a_seg[0]; kp.b_seg[1] = kp.b_seg[0]; }
    kp.seg_a_bytes = plan->seg_rows * 128;
    kp.seg_btile_bytes = b_rows * 128;
    // weights resident in shared memory when every tile of the launch uses the same ones (window layers: one N tile, one
    // 64-wide K chunk per tap, one segment per tile): the per-tile weight re-load was 2/3 of the staged bytes
    {
        const char* e = getenv("ACLGAN_SEG_BRES");
        kp.seg_bres = (!pair && plan->n_tiles == 1 && plan->cchunks == 1 && plan->num_segs == 1 && (e == nullptr || atoi(e) != 0)) ? 1 : 0;
    }
    const int bres_bytes = kp.seg_bres ? plan->planes * plan->seg_taps * kp.seg_btile_bytes : 0;
    // Two M tiles per CTA and work item (m_sub = 2, experimental): every staged weight tile feeds two accumulators (half the
    // weight bytes per FLOP through the ~42 B/clk/SM L2->SM path) and the tile count per wave doubles (batch-8 res-block conv:
    // 137 tile pairs = 69 items = ONE wave on 74 clusters instead of 1.85).  Needs 2 * block_n <= 512 TMEM columns and the
    // whole shared memory (two stages of 2 A segments + the row's weight tiles).
    const int pair_items1 = ((m_tiles + 1) / 2) * plan->n_tiles;
    // MEASURED (tools/bench_layers.py, batch 8): forward 52.9 -> 66.5 us, dgrad 42.4 -> 45.4 us on the 3x3 256->256 layer: with
    // N = 256 the two accumulators fill TMEM, the epilogue (as long as the MMAs of a tile) is no longer overlapped, and that
    // costs more than the halved weight traffic and the single wave save.  Kept behind ACLGAN_SEG_MSUB=2, default 1.
    int m_sub = 1;
    (void)pair_items1;
    {
        const char* e = getenv("ACLGAN_SEG_MSUB");
        if (e != nullptr) m_sub = (atoi(e) == 2 && 2 * plan->block_n <= 512) ? 2 : 1;
    }
    decide_egroups(plan, kp, m_sub * (plan->block_n < 32 ? 32 : plan->block_n));
    if (m_sub != 1) kp.seg_egroups = 1;
    // shared memory requested per CTA: not all of it when it is not needed, so that an element-wise CTA of another chain
    // (<= 24 KB) can be co-resident on the SM and run under the tensor pipe's shadow (env ACLGAN_SEG_SMEM_KB overrides)
    static int seg_smem_default = 0;
    if (seg_smem_default == 0) {
        const char* e = getenv("ACLGAN_SEG_SMEM_KB");
        seg_smem_default = (e != nullptr ? atoi(e) : 176) * 1024;
        if (seg_smem_default > kSegSmemBytes || seg_smem_default < 96 * 1024) seg_smem_default = kSegSmemBytes;
    }
    int seg_smem = seg_smem_default, ns = 0;
    for (;;) {
        kp.seg_msub = m_sub;
        kp.seg_stage_bytes = m_sub * kp.seg_a_bytes + (kp.seg_bres ? 0 : plan->seg_taps * kp.seg_btile_bytes);
        ns = (seg_smem - 1024 - kStageOutBytes - 512 - bres_bytes) / kp.seg_stage_bytes;
        if (ns >= 2) break;
        if (seg_smem < kSegSmemBytes) { seg_smem = kSegSmemBytes; continue; }      // m_sub = 2 stages need the whole budget
        if (m_sub == 2) { m_sub = 1; seg_smem = seg_smem_default; continue; }
        return -100;
    }
    if (ns > kSegMaxStages) ns = kSegMaxStages;
    kp.seg_ns = ns;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(igemm_seg_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (pair) {
        const int items = ((m_tiles + 2 * m_sub - 1) / (2 * m_sub)) * plan->n_tiles;
        int clusters = num_sms() / 2;
        if (items < clusters) clusters = items;
        for (int i = 0; i < repeat; ++i) igemm_seg_pair_kernel<<<2 * clusters, kSegThreads, seg_smem, stream>>>(kp);
    } else {
        const int items = ((m_tiles + m_sub - 1) / m_sub) * plan->n_tiles;
        if (items <= 0) return ACLGAN_OK;
        const int grid = items < num_sms() ? items : num_sms();
        for (int i = 0; i < repeat; ++i) igemm_seg_kernel<<<grid, kSegThreads, seg_smem, stream>>>(kp);
    }
    return (int)cudaGetLastError();
}

}  // namespace aclgan

extern "C" int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream);

// perf triage: device buffer of [grid][8] int64 role timers filled by the segment kernel (0 switches it off):
// 0 producer total, 1 producer waiting for free stages, 2 MMA thread total, 3 MMA waiting for operands, 4 MMA waiting for
// a drained accumulator, 5 epilogue warp total, 6 epilogue waiting for a finished accumulator
extern "C" int aclgan_igemm_set_prof(uint64_t buf) { aclgan::g_prof = reinterpret_cast<long long*>(buf); return 0; }

extern "C" int aclgan_igemm_stats_supported(const aclgan_igemm_plan* plan) { return aclgan::stats_supported(plan) ? 1 : 0; }

extern "C" int aclgan_igemm_launch(const aclgan_igemm_plan* plan, void* stream) {
    return aclgan_igemm_launch_repeat(plan, 1, stream);
}

// launches the same plan `repeat` times back to back (tensor maps encoded once): device-side timing of the kernel
// without per-launch host work (bench.py roofline leg, tools/triage_igemm.py)
extern "C" int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream) {
    using namespace aclgan;
    if (plan->fold) return aclgan_fold_launch(plan, repeat, stream);      // (experimental fold-mode plans, csrc/fold.cu)
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
            decide_egroups(plan, kp, plan->block_n < 32 ? 32 : plan->block_n);
            const int items = ((m_tiles_all + 1) / 2) * plan->n_tiles * n_groups;
            int clusters = num_sms() / 2;
            if (items < clusters) clusters = items;
            for (int i = 0; i < repeat; ++i)
                igemm_pair_kernel<<<2 * clusters, kEpiThreads, kPairSmemBytes, (cudaStream_t)stream>>>(kp);
            return (int)cudaGetLastError();
        }
    }
    decide_egroups(plan, kp, kp.m_sub * (plan->block_n < 32 ? 32 : plan->block_n));
    const int total = ((m_tiles_all + kp.m_sub - 1) / kp.m_sub) * plan->n_tiles * kp.n_groups;
    if (total <= 0) return ACLGAN_OK;
    const int grid = total < num_sms() ? total : num_sms();
    for (int i = 0; i < repeat; ++i) igemm_kernel<<<grid, kEpiThreads, kSmemBytes, (cudaStream_t)stream>>>(kp);
    return (int)cudaGetLastError();
}
