// EXPERIMENTAL (off by default, env ACLGAN_FOLD=1): forward of the few-output-channel final convolution (7x7, 64 -> 3|4,
// reference networks.py:260) with the filter COLUMNS folded into the GEMM N dimension.
//
// The plain / segment kernels issue one 128 x 16 x 16 MMA per (tap, k-slice): 196 MMAs per 128-pixel tile, each costing
// the ~69-cycle tcgen05 floor whatever N - the layer runs at 35 TFLOP/s.  Here the weight matrix has rows n = kw*8 + co
// (N = 64) and one "tap" per filter ROW: P[q][kw*8 + co] = sum_{kh, ci} X[q + kh*Wp][ci] * W[co][ci][kh][kw] over the
// flattened padded grid, 7 x 4 MMAs per tile, and the epilogue finishes the convolution with a diagonal sum
// out[q][co] = sum_kw P[q + kw][kw*8 + co] through a shared-memory copy of the accumulator tile (tiles are 128 positions
// stepping by 120, so the 7 rows a position needs are always in its own tile).
#include "common.cuh"

namespace aclgan {

constexpr int kFStages = 6;
constexpr int kFABytes = 128 * 128;
constexpr int kFBBytes = 64 * 128;
constexpr int kFStageBytes = kFABytes + kFBBytes;
constexpr int kFPitch = 65;                                   // floats per staged accumulator row (bank-conflict free)
constexpr int kFStageOut = 128 * kFPitch * 4;
constexpr int kFSmemBytes = kFStages * kFStageBytes + kFStageOut + 1024 + 256;
constexpr int kFThreads = 256;

struct alignas(64) FoldKParams {
    CUtensorMap a[2], b[2];
    int planes, nseg, cchunks, k, tiles, tile_step, flat_w, flat_img;
    int tap_dx[8], tap_bk[8];
    uint64_t ptr, bias;
    int64_t off, sn, sy, sx, sc;
    int N, H, W, C, bias_n, act, kind;
    float slope;
};

__global__ void __launch_bounds__(kFThreads, 1) igemm_fold_kernel(const __grid_constant__ FoldKParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* ptile = reinterpret_cast<float*>(smem + kFStages * kFStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFStages * kFStageBytes + kFStageOut);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kFStages;
    uint64_t* tfull_bar = bars + 2 * kFStages;
    uint64_t* tempty_bar = bars + 2 * kFStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int p = 0; p < P.planes; ++p) { tma_prefetch_desc(&P.a[p]); tma_prefetch_desc(&P.b[p]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kFStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int k_iters = P.nseg * P.k * P.cchunks;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
            const int x0 = tile * P.tile_step;
#pragma unroll 1
            for (int seg = 0; seg < P.nseg; ++seg) {
                const int pa = (seg == 2) ? 1 : 0, pb = (seg == 1) ? 1 : 0;
#pragma unroll 1
                for (int t = 0; t < P.k; ++t) {
#pragma unroll 1
                    for (int cc = 0; cc < P.cchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * kFStageBytes;
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kFStageBytes);
                        tma_load_4d(sa, &P.a[pa], &full_bar[stage], cc * 64, x0 + P.tap_dx[t], 0, 0);
                        tma_load_2d(sa + kFABytes, &P.b[pb], &full_bar[stage], P.tap_bk[t] + cc * 64, 0);
                        if (++stage == kFStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll 1
            for (int k = 0; k < k_iters; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * kFStageBytes);
                const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(sa + kFABytes, 16, 1024);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (++stage == kFStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[acc]);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: TMEM -> shared memory, diagonal sum, bias, activation, store ----------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const float* bias = reinterpret_cast<const float*>(P.bias);
        int it = 0;
        for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 64 + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) ptile[row * kFPitch + c + i] = __uint_as_float(raw[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);           // the accumulator is free for the tile after next
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (row < P.tile_step) {
                const int64_t qq = (int64_t)tile * P.tile_step + row;
                const int z = (int)(qq / P.flat_img);
                const int rem = (int)(qq - (int64_t)z * P.flat_img);
                const int y = rem / P.flat_w, x = rem - y * P.flat_w;
                if (x < P.W && y < P.H && z < P.N) {
                    const int64_t pix = P.off + (int64_t)z * P.sn + (int64_t)y * P.sy + (int64_t)x * P.sx;
#pragma unroll 1
                    for (int co = 0; co < P.C; ++co) {
                        float v = 0.f;
                        for (int kw = 0; kw < P.k; ++kw) v += ptile[(row + kw) * kFPitch + kw * 8 + co];
                        if (bias != nullptr && co < P.bias_n) v += __ldg(bias + co);
                        if (P.act == ACLGAN_ACT_TANH) v = tanhf(v);
                        else if (P.act == ACLGAN_ACT_RELU) v = fmaxf(v, 0.f);
                        else if (P.act == ACLGAN_ACT_LRELU) v = v > 0.f ? v : v * P.slope;
                        if (P.kind == ACLGAN_OUT_F32) reinterpret_cast<float*>(P.ptr)[pix + (int64_t)co * P.sc] = v;
                        else reinterpret_cast<__nv_bfloat16*>(P.ptr)[pix + (int64_t)co * P.sc] = __float2bfloat16_rn(v);
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the staged tile may be overwritten
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

}  // namespace aclgan

extern "C" int aclgan_fold_launch(const aclgan_igemm_plan* pl, int repeat, void* stream) {
    using namespace aclgan;
    if (!pl->fold || pl->fold > 8 || pl->block_n != 64 || !pl->flat || pl->tile_step + pl->fold - 1 > 128 || pl->tile_step < 1)
        return ACLGAN_ERR_SHAPE;
    if (pl->planes < 1 || pl->planes > 2 || (pl->nseg != 1 && pl->nseg != 3) || pl->num_taps != pl->fold) return ACLGAN_ERR_SHAPE;
    const aclgan_out_spec& o = pl->out;
    if ((o.kind != ACLGAN_OUT_F32 && o.kind != ACLGAN_OUT_BF16) || o.C > 8 || o.mirror != 0 || o.stats != 0) return ACLGAN_ERR_UNSUPPORTED;
    FoldKParams kp;
    for (int p = 0; p < 2; ++p) {
        const int sp = p < pl->planes ? p : 0;
        int rc = encode_tmap(&pl->a[sp][0], &kp.a[p]);
        if (rc) return rc;
        rc = encode_tmap(&pl->b[sp], &kp.b[p]);
        if (rc) return rc;
    }
    kp.planes = pl->planes; kp.nseg = pl->nseg; kp.cchunks = pl->cchunks; kp.k = pl->fold;
    kp.tiles = pl->tiles_x; kp.tile_step = pl->tile_step; kp.flat_w = pl->flat_w; kp.flat_img = pl->flat_img;
    for (int t = 0; t < 8; ++t) { kp.tap_dx[t] = pl->tap_dx[t]; kp.tap_bk[t] = pl->tap_bk[t]; }
    kp.ptr = o.ptr[0]; kp.bias = o.bias; kp.off = o.off; kp.sn = o.sn; kp.sy = o.sy; kp.sx = o.sx; kp.sc = o.sc;
    kp.N = o.N; kp.H = o.H; kp.W = o.W; kp.C = o.C; kp.bias_n = o.bias_n; kp.act = o.act; kp.kind = o.kind; kp.slope = o.slope;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (kp.tiles <= 0) return ACLGAN_OK;
    const int grid = kp.tiles < num_sms() ? kp.tiles : num_sms();
    for (int i = 0; i < repeat; ++i) igemm_fold_kernel<<<grid, kFThreads, kFSmemBytes, (cudaStream_t)stream>>>(kp);
    return (int)cudaGetLastError();
}
