// Convolution weight gradient on the tensor cores:
//
//   dW[m][tap][n] += sum_{pixels of this CTA's K range}  Mop[pixel][m] * Nop_tap[pixel][n]
//
// The reduction dimension is the pixel index, which is the OUTER (row) dimension of the NHWC planes, so both
// operands are fed to tcgen05.mma in MN-major form: a TMA box of 64 pixels x 64 channels lands as 64 rows of
// 128 B (128B swizzle) and is consumed as a [K=64 pixels][64 channels] block (descriptor LBO = distance between
// 64-channel chunks, SBO = 8 pixel rows).  M = 128 channels of one operand (dY normally), N = up to 256 channels
// of the other, taken at the filter-tap offset; fp32 accumulator in TMEM; split-K over pixel blocks with
// fp32 vector atomics into the gradient buffer, which has the packed-weight layout.
//
// Replaces: the weight-gradient half of autograd's convolution_backward for nn.Conv2d (reference
// networks.py:363,366 under loss.backward(), trainer.py:169,292).
#include <cstdlib>

#include "common.cuh"

namespace aclgan {

constexpr int kWStages = 4;
constexpr int kWChunkBytes = 64 * 128;               // 64 pixels x 64 channels bf16
constexpr int kWMBytes = 2 * kWChunkBytes;           // M operand: up to 2 chunks
constexpr int kWNBytes = 4 * kWChunkBytes;           // N operand: up to 4 chunks
constexpr int kWStageBytes = kWMBytes + kWNBytes;    // 48 KB
constexpr int kWStageOut = 4 * 4096;                 // epilogue staging: 4 warps x (32 rows x 128 B)
constexpr int kWSmemBytes = kWStages * kWStageBytes + kWStageOut + 1024 + 256;
constexpr int kWThreads = 256;

struct alignas(64) WgradKParams {
    CUtensorMap mop[2][ACLGAN_MAX_AVARIANTS];
    CUtensorMap nop[2][ACLGAN_MAX_AVARIANTS];
    int planes, nseg, m_chunks, n_chunks, m_tiles, n_tiles;
    int pix;     // pixels per stage (64 | 128): small-channel layers use 128 to halve the per-stage barrier overhead
    int box_x, box_y, box_z, blocks_x, blocks_y, blocks_z, ksplit, num_taps;
    int m_dx[ACLGAN_MAX_TAPS], m_dy[ACLGAN_MAX_TAPS], m_var[ACLGAN_MAX_TAPS];
    int n_dx[ACLGAN_MAX_TAPS], n_dy[ACLGAN_MAX_TAPS], n_var[ACLGAN_MAX_TAPS];
    int tap_out[ACLGAN_MAX_TAPS];
    float* dw;
    long long dw_sm, dw_st;
    int M, Nn;
    int debug;   // perf triage (env ACLGAN_WGRAD_DEBUG): 1 = skip the atomics, 2 = direct (unstaged) atomics
    // segment mode (wgrad_seg_kernel)
    CUtensorMap seg_map[2];
    int seg_rows, seg_taps, seg_on_m;
    int seg_m_bytes, seg_n_bytes, seg_stages;    // per-stage operand regions / ring depth chosen by the host
    int seg_step;      // rows of the staged segment between consecutive taps (1: horizontal segments; box_x: vertical window segments)
    int seg_merge;     // > 1: that many taps are ONE MMA (N = 64 * taps): their 64-column operand blocks are seg_step rows apart, which
                       // is the descriptor's leading-dimension offset (N-shifted, one 64-channel chunk, seg_step * 128 B % 1024 == 0)
    int seg_kw0[ACLGAN_MAX_TAPS], seg_cnt[ACLGAN_MAX_TAPS];
};

__global__ void __launch_bounds__(kWThreads, 1) wgrad_kernel(const __grid_constant__ WgradKParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_out = smem + kWStages * kWStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWStages * kWStageBytes + kWStageOut);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWStages;
    uint64_t* done_bar = bars + 2 * kWStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile decode: blockIdx.x = ((tap * m_tiles + mt) * n_tiles + nt) * ksplit + split
    int id = blockIdx.x;
    const int split = id % P.ksplit;  id /= P.ksplit;
    const int nt = id % P.n_tiles;    id /= P.n_tiles;
    const int mt = id % P.m_tiles;    id /= P.m_tiles;
    const int tap = id;

    const int blocks_total = P.blocks_x * P.blocks_y * P.blocks_z;
    const int per = (blocks_total + P.ksplit - 1) / P.ksplit;
    const int blk_begin = split * per;
    const int blk_end = min(blocks_total, blk_begin + per);
    const int n_iters = (blk_end > blk_begin ? blk_end - blk_begin : 0) * P.nseg;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int chunk_bytes = P.pix * 128;                       // one 64-channel chunk of P.pix pixels
    const int n_off = P.pix == 64 ? kWMBytes : P.m_chunks * chunk_bytes;   // N operand region inside a stage
    const uint32_t stage_tx = (uint32_t)(P.m_chunks + P.n_chunks) * chunk_bytes;

    if (n_iters > 0) {
        if (warp == 0 && lane == 0) {
            // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            for (int b = blk_begin; b < blk_end; ++b) {
                const int bx = b % P.blocks_x;
                const int by = (b / P.blocks_x) % P.blocks_y;
                const int bz = b / (P.blocks_x * P.blocks_y);
                const int x0 = bx * P.box_x, y0 = by * P.box_y, z0 = bz * P.box_z;
                for (int seg = 0; seg < P.nseg; ++seg) {
                    const int pm = (seg == 2) ? 1 : 0;
                    const int pn = (seg == 1) ? 1 : 0;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sm = smem + stage * kWStageBytes;
                    uint8_t* sn = sm + n_off;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                    const CUtensorMap* mm = &P.mop[pm][P.m_var[tap]];
                    const CUtensorMap* nm = &P.nop[pn][P.n_var[tap]];
                    for (int c = 0; c < P.m_chunks; ++c)
                        tma_load_4d(sm + c * chunk_bytes, mm, &full_bar[stage], (mt * 2 + c) * 64, x0 + P.m_dx[tap],
                                    y0 + P.m_dy[tap], z0);
                    for (int c = 0; c < P.n_chunks; ++c)
                        tma_load_4d(sn + c * chunk_bytes, nm, &full_bar[stage], (nt * P.n_chunks + c) * 64,
                                    x0 + P.n_dx[tap], y0 + P.n_dy[tap], z0);
                    if (++stage == kWStages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)(64 * P.n_chunks), 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < n_iters; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sm = smem_u32(smem + stage * kWStageBytes);
                const uint32_t sn = sm + n_off;
                // MN-major, 128B swizzle: LBO = next 64-channel chunk, SBO = next 8 pixel rows (1 KB)
                const uint64_t dm = make_smem_desc_sw128(sm, chunk_bytes, 1024);
                const uint64_t dn = make_smem_desc_sw128(sn, chunk_bytes, 1024);
                const int n_mma = P.pix / 16;
#pragma unroll 4
                for (int kk = 0; kk < n_mma; ++kk) {
                    // 16 pixels (= one UMMA K) further down: 16 rows x 128 B = 2048 B -> +128 encoded
                    umma_bf16(tmem_base, dm + 128 * kk, dn + 128 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == kWStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        } else if (warp >= 4) {
            // ---------------- epilogue: fp32 vector atomics into dW ----------------
            const int q = warp & 3;
            const int m = mt * 128 + q * 32 + lane;
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
            const int ncols = 64 * P.n_chunks;
            const int n_base = nt * ncols;
            float* row = P.dw + (long long)m * P.dw_sm + (long long)P.tap_out[tap] * P.dw_st;
            uint8_t* stg = stage_out + q * 4096;
            const bool row_ok = m < P.M;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
                if (P.debug == 1) continue;
                const int n0 = n_base + c;
                float* dst = row + n0;
                const bool full = (n0 + 32 <= P.Nn) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && P.debug != 2;
                if (__all_sync(0xffffffffu, full || !row_ok)) {
                    // stage 32 rows x 128 B, then 8 lanes add one full 128-byte line each (vector reductions that the
                    // memory system can merge per line instead of 32 scattered 16-byte reductions per instruction)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint4 qv;
                        qv.x = raw[4 * i]; qv.y = raw[4 * i + 1]; qv.z = raw[4 * i + 2]; qv.w = raw[4 * i + 3];
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((i ^ (lane & 7)) << 4)) = qv;
                    }
                    __syncwarp();
                    const int piece = lane & 7;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3);
                        const long long roff = __shfl_sync(0xffffffffu, (long long)(dst - P.dw), r);
                        const int ok = __shfl_sync(0xffffffffu, (int)row_ok, r);
                        if (ok) {
                            const float4 fv = *reinterpret_cast<const float4*>(stg + r * 128 + ((piece ^ (r & 7)) << 4));
                            atomicAdd(reinterpret_cast<float4*>(P.dw + roff + piece * 4), fv);
                        }
                    }
                    __syncwarp();
                } else if (row_ok) {
#pragma unroll 1
                    for (int i = 0; i < 32; ++i)
                        if (n0 + i < P.Nn) atomicAdd(dst + i, __uint_as_float(raw[i]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// =====================================================================================================================
// Segment variant (stride-1 convolutions): a CTA reduces all k taps of ONE filter row.  Per 64-pixel block of an output
// row the dY tile is staged once and the conv-input pixels once, as a segment of 64 + k - 1 pixels; tap kw multiplies
// the dY tile with the segment rows [kw, kw + 64) (descriptor start shifted by kw * 128 B: the swizzle is a function of
// the shared-memory address, see igemm.cu) into its own TMEM accumulator.  k times fewer dY bytes and ~k times fewer
// input bytes cross L2 -> shared memory than with one CTA per tap, which is what bounds the plain kernel.
// =====================================================================================================================
constexpr int kWSegMaxStages = 6;
constexpr int kWSegSmemBytes = 232448;

__global__ void __launch_bounds__(kWThreads, 1) wgrad_seg_kernel(const __grid_constant__ WgradKParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = P.seg_m_bytes + P.seg_n_bytes;
    uint8_t* stage_out = smem + P.seg_stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kWStageOut);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWSegMaxStages;
    uint64_t* done_bar = bars + 2 * kWSegMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWSegMaxStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile decode: blockIdx.x = ((row tap * m_tiles + mt) * n_tiles + nt) * ksplit + split
    int id = blockIdx.x;
    const int split = id % P.ksplit;  id /= P.ksplit;
    const int nt = id % P.n_tiles;    id /= P.n_tiles;
    const int mt = id % P.m_tiles;    id /= P.m_tiles;
    const int tap = id;

    const int blocks_total = P.blocks_x * P.blocks_y * P.blocks_z;
    const int per = (blocks_total + P.ksplit - 1) / P.ksplit;
    const int blk_begin = split * per;
    const int blk_end = min(blocks_total, blk_begin + per);
    const int n_iters = (blk_end > blk_begin ? blk_end - blk_begin : 0) * P.nseg;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWSegMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
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
    const int m_chunk_bytes = P.seg_m_bytes / P.m_chunks;       // 64- or seg_rows-pixel chunk
    const int n_chunk_bytes = P.seg_n_bytes / P.n_chunks;
    const int ncols = 64 * P.n_chunks;                          // accumulator columns per tap
    const int kw0 = P.seg_kw0[tap], cnt = P.seg_cnt[tap];       // this CTA's taps of the filter row

    if (n_iters > 0) {
        if (warp == 0 && lane == 0) {
            // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            for (int b = blk_begin; b < blk_end; ++b) {
                const int bx = b % P.blocks_x;
                const int by = (b / P.blocks_x) % P.blocks_y;
                const int bz = b / (P.blocks_x * P.blocks_y);
                const int x0 = bx * P.box_x, y0 = by * P.box_y, z0 = bz * P.box_z;
                for (int seg = 0; seg < P.nseg; ++seg) {
                    const int pm = (seg == 2) ? 1 : 0;
                    const int pn = (seg == 1) ? 1 : 0;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sm = smem + stage * stage_bytes;
                    uint8_t* sn = sm + P.seg_m_bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                    const CUtensorMap* mm = P.seg_on_m ? &P.seg_map[pm] : &P.mop[pm][0];
                    const CUtensorMap* nm = P.seg_on_m ? &P.nop[pn][0] : &P.seg_map[pn];
                    for (int c = 0; c < P.m_chunks; ++c)
                        tma_load_4d(sm + c * m_chunk_bytes, mm, &full_bar[stage], (mt * 2 + c) * 64, x0 + P.m_dx[tap],
                                    y0 + P.m_dy[tap], z0);
                    for (int c = 0; c < P.n_chunks; ++c)
                        tma_load_4d(sn + c * n_chunk_bytes, nm, &full_bar[stage], (nt * P.n_chunks + c) * 64,
                                    x0 + P.n_dx[tap], y0 + P.n_dy[tap], z0);
                    if (++stage == P.seg_stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)ncols, 1, 1);
            const uint32_t step8 = 8u * (uint32_t)P.seg_step;                                  // one pixel row = 128 B = 8 units
            const uint32_t shift_m = P.seg_on_m ? step8 : 0u, shift_n = P.seg_on_m ? 0u : step8;
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < n_iters; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sm = smem_u32(smem + stage * stage_bytes);
                const uint32_t sn = sm + P.seg_m_bytes;
                const uint64_t dm0 = make_smem_desc_sw128(sm, m_chunk_bytes, 1024);
                if (P.seg_merge > 1) {
                    // groups of seg_merge taps as one MMA of N = 64 * taps: column block c of the N operand = tap c of the group
                    const uint64_t dng = make_smem_desc_sw128(sn, P.seg_step * 128, 1024);
#pragma unroll 1
                    for (int j0 = 0; j0 < cnt; j0 += P.seg_merge) {
                        const int g = min(P.seg_merge, cnt - j0);
                        const uint32_t idg = make_idesc_bf16(128, (uint32_t)(64 * g), 1, 1);
                        const uint64_t dn = dng + (uint64_t)((kw0 + j0) * shift_n);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + j0 * 64, dm0 + 128 * kk, dn + 128 * kk, idg, (k | kk) != 0 ? 1u : 0u);
                    }
                } else {
                    const uint64_t dn0 = make_smem_desc_sw128(sn, n_chunk_bytes, 1024);
#pragma unroll 1
                    for (int j = 0; j < cnt; ++j) {
                        const uint64_t dm = dm0 + (uint64_t)((kw0 + j) * shift_m), dn = dn0 + (uint64_t)((kw0 + j) * shift_n);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)        // 16 pixels (one UMMA K) = 2048 B -> +128 encoded
                            umma_bf16(tmem_base + j * ncols, dm + 128 * kk, dn + 128 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == P.seg_stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        } else if (warp >= 4) {
            // ---------------- epilogue: fp32 vector atomics into dW, one accumulator per tap ----------------
            const int q = warp & 3;
            const int m = mt * 128 + q * 32 + lane;
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const int n_base = nt * ncols;
            uint8_t* stg = stage_out + q * 4096;
            const bool row_ok = m < P.M;
#pragma unroll 1
            for (int kw = 0; kw < cnt; ++kw) {
                const uint32_t t_row = tmem_base + kw * ncols + ((uint32_t)(q * 32) << 16);
                float* row = P.dw + (long long)m * P.dw_sm + (long long)(P.tap_out[tap] + kw) * P.dw_st;
#pragma unroll 1
                for (int c = 0; c < ncols; c += 32) {
                    uint32_t raw[32];
                    tmem_ld_32x32(t_row + c, raw);
                    tmem_ld_wait();
                    const int n0 = n_base + c;
                    float* dst = row + n0;
                    const bool full = (n0 + 32 <= P.Nn) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
                    if (__all_sync(0xffffffffu, full || !row_ok)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4 qv;
                            qv.x = raw[4 * i]; qv.y = raw[4 * i + 1]; qv.z = raw[4 * i + 2]; qv.w = raw[4 * i + 3];
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((i ^ (lane & 7)) << 4)) = qv;
                        }
                        __syncwarp();
                        const int piece = lane & 7;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = 4 * i + (lane >> 3);
                            const long long roff = __shfl_sync(0xffffffffu, (long long)(dst - P.dw), r);
                            const int ok = __shfl_sync(0xffffffffu, (int)row_ok, r);
                            if (ok) {
                                const float4 fv = *reinterpret_cast<const float4*>(stg + r * 128 + ((piece ^ (r & 7)) << 4));
                                atomicAdd(reinterpret_cast<float4*>(P.dw + roff + piece * 4), fv);
                            }
                        }
                        __syncwarp();
                    } else if (row_ok) {
#pragma unroll 1
                        for (int i = 0; i < 32; ++i)
                            if (n0 + i < P.Nn) atomicAdd(dst + i, __uint_as_float(raw[i]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace aclgan

using namespace aclgan;

extern "C" int aclgan_wgrad_launch_repeat(const aclgan_wgrad_plan* pl, int repeat, void* stream);

extern "C" int aclgan_wgrad_launch(const aclgan_wgrad_plan* pl, void* stream) {
    return aclgan_wgrad_launch_repeat(pl, 1, stream);
}

extern "C" int aclgan_wgrad_launch_repeat(const aclgan_wgrad_plan* pl, int repeat, void* stream) {
    static bool attr_set = false;
    if (pl->planes < 1 || pl->planes > 2 || (pl->nseg != 1 && pl->nseg != 3)) return ACLGAN_ERR_SHAPE;
    if (pl->m_chunks < 1 || pl->m_chunks > 2 || pl->n_chunks < 1 || pl->n_chunks > 4) return ACLGAN_ERR_SHAPE;
    const int pix = pl->box_x * pl->box_y * pl->box_z;
    if (pix != 64 && pix != 128) return ACLGAN_ERR_SHAPE;
    if (pix == 128 && (pl->m_chunks + pl->n_chunks) * 128 * 128 > kWStageBytes) return ACLGAN_ERR_SHAPE;
    if (pl->num_taps < 1 || pl->num_taps > ACLGAN_MAX_TAPS || pl->ksplit < 1) return ACLGAN_ERR_SHAPE;
    if (pl->n_mvariants < 1 || pl->n_mvariants > ACLGAN_MAX_AVARIANTS || pl->n_nvariants < 1 ||
        pl->n_nvariants > ACLGAN_MAX_AVARIANTS)
        return ACLGAN_ERR_SHAPE;
    WgradKParams kp;
    for (int p = 0; p < 2; ++p) {
        const int sp = p < pl->planes ? p : 0;
        for (int v = 0; v < ACLGAN_MAX_AVARIANTS; ++v) {
            int rc = encode_tmap(&pl->mop[sp][v < pl->n_mvariants ? v : 0], &kp.mop[p][v]);
            if (rc) return rc;
            rc = encode_tmap(&pl->nop[sp][v < pl->n_nvariants ? v : 0], &kp.nop[p][v]);
            if (rc) return rc;
        }
    }
    kp.planes = pl->planes; kp.nseg = pl->nseg; kp.m_chunks = pl->m_chunks; kp.n_chunks = pl->n_chunks;
    kp.pix = pix;
    kp.m_tiles = pl->m_tiles; kp.n_tiles = pl->n_tiles;
    kp.box_x = pl->box_x; kp.box_y = pl->box_y; kp.box_z = pl->box_z;
    kp.blocks_x = pl->blocks_x; kp.blocks_y = pl->blocks_y; kp.blocks_z = pl->blocks_z;
    kp.ksplit = pl->ksplit; kp.num_taps = pl->num_taps;
    for (int t = 0; t < ACLGAN_MAX_TAPS; ++t) {
        kp.m_dx[t] = pl->m_dx[t]; kp.m_dy[t] = pl->m_dy[t]; kp.m_var[t] = pl->m_var[t];
        kp.n_dx[t] = pl->n_dx[t]; kp.n_dy[t] = pl->n_dy[t]; kp.n_var[t] = pl->n_var[t];
        kp.tap_out[t] = pl->tap_out[t];
    }
    kp.dw = reinterpret_cast<float*>(pl->dw); kp.dw_sm = pl->dw_sm; kp.dw_st = pl->dw_st;
    kp.M = pl->M; kp.Nn = pl->Nn;
    {
        const char* dbg = getenv("ACLGAN_WGRAD_DEBUG");
        kp.debug = dbg != nullptr ? atoi(dbg) : 0;
    }
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wgrad_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSegSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = pl->num_taps * pl->m_tiles * pl->n_tiles * pl->ksplit;
    if (pl->seg_mode) {
        const int step = pl->seg_step > 0 ? pl->seg_step : 1;
        if (pix != 64 || pl->seg_rows % 8 != 0 || pl->seg_rows < 64 + (pl->seg_taps - 1) * step || pl->seg_rows > 256 ||
            pl->seg_taps * 64 * pl->n_chunks > 512)
            return ACLGAN_ERR_SHAPE;
        for (int p = 0; p < 2; ++p) {
            int rc = encode_tmap(&pl->seg_map[p < pl->planes ? p : 0], &kp.seg_map[p]);
            if (rc) return rc;
        }
        kp.seg_rows = pl->seg_rows; kp.seg_taps = pl->seg_taps; kp.seg_on_m = pl->seg_on_m;
        for (int t = 0; t < ACLGAN_MAX_TAPS; ++t) {
            kp.seg_kw0[t] = pl->seg_kw0[t]; kp.seg_cnt[t] = pl->seg_cnt[t];
            if (t < pl->num_taps && (pl->seg_cnt[t] < 1 || pl->seg_cnt[t] > pl->seg_taps ||
                                     (pl->seg_kw0[t] + pl->seg_cnt[t] - 1) * step + 64 > pl->seg_rows))
                return ACLGAN_ERR_SHAPE;
        }
        kp.seg_step = step;
        {
            const char* e = getenv("ACLGAN_WGRAD_MERGE");
            const bool on = e == nullptr || atoi(e) != 0;
            kp.seg_merge = (on && !pl->seg_on_m && pl->n_chunks == 1 && (step * 128) % 1024 == 0 && pl->seg_taps > 1) ? 4 : 1;
        }
        kp.seg_m_bytes = pl->m_chunks * (pl->seg_on_m ? pl->seg_rows : 64) * 128;
        kp.seg_n_bytes = pl->n_chunks * (pl->seg_on_m ? 64 : pl->seg_rows) * 128;
        // (not the whole shared memory: room for a co-resident element-wise CTA of another chain, see igemm.cu)
        static int seg_smem = 0;
        if (seg_smem == 0) {
            const char* e = getenv("ACLGAN_SEG_SMEM_KB");
            seg_smem = (e != nullptr ? atoi(e) : 176) * 1024;
            if (seg_smem > kWSegSmemBytes || seg_smem < 96 * 1024) seg_smem = kWSegSmemBytes;
        }
        int stages = (seg_smem - 1024 - kWStageOut - 256) / (kp.seg_m_bytes + kp.seg_n_bytes);
        if (stages > kWSegMaxStages) stages = kWSegMaxStages;
        if (stages < 2) return ACLGAN_ERR_SHAPE;
        kp.seg_stages = stages;
        for (int i = 0; i < repeat; ++i) wgrad_seg_kernel<<<grid, kWThreads, seg_smem, (cudaStream_t)stream>>>(kp);
        return (int)cudaGetLastError();
    }
    for (int i = 0; i < repeat; ++i) wgrad_kernel<<<grid, kWThreads, kWSmemBytes, (cudaStream_t)stream>>>(kp);
    return (int)cudaGetLastError();
}
