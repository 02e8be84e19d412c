/* aclgan_b200 - C ABI of the B200-native ACL-GAN convolutional training step.
 *
 * The reference (hyperplane-lab/ACL-GAN) has NO native/FFI layer: its hot path is eager PyTorch
 * (networks.py / trainer.py -> ATen -> cuDNN).  This header is therefore the boundary SURVEY.md
 * section 8(b) "Level 2" defines: plain `extern "C"` entry points, raw device pointers + sizes, a
 * `cudaStream_t` passed as `void*`, int status return.  Each entry point names the reference
 * call site (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch tensors); never retained/freed;
 *   - launches go to the given stream, no hidden synchronisation, capturable in a CUDA graph;
 *   - return 0 on success, < 0 = aclgan error (bad shape / alignment / unsupported option),
 *     > 0 = cudaError_t / CUresult of the failing runtime call;
 *   - activations are "padded NHWC planes": [N][H+2p][W+2p][C] bf16, C a multiple of 8, one plane
 *     (bf16 mode) or two planes hi/lo (bf16x3 mode: x ~= hi + lo, products hi*hi + hi*lo + lo*hi
 *     accumulate in fp32 on the tensor cores, giving ~fp32 products for the parity mode).
 */
#ifndef ACLGAN_B200_H
#define ACLGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACLGAN_ABI_VERSION 1
#define ACLGAN_MAX_TAPS 64
#define ACLGAN_MAX_AVARIANTS 4

enum { ACLGAN_OK = 0, ACLGAN_ERR_SHAPE = -1, ACLGAN_ERR_ALIGN = -2, ACLGAN_ERR_UNSUPPORTED = -3, ACLGAN_ERR_DRIVER = -4 };
enum { ACLGAN_ACT_NONE = 0, ACLGAN_ACT_RELU = 1, ACLGAN_ACT_LRELU = 2, ACLGAN_ACT_TANH = 3 };
enum { ACLGAN_WINDOW_NONE = 0, ACLGAN_WINDOW_IN = 1 /* cin <= 8|16: first convs */, ACLGAN_WINDOW_OUT = 2 /* cout <= 8: final conv */ };
enum { ACLGAN_OUT_BF16 = 0, ACLGAN_OUT_F32 = 1, ACLGAN_OUT_SPLIT = 2, ACLGAN_OUT_F32_ATOMIC = 3 };

/* ---- generic TMA tensor-map description (bf16, 128B swizzle, zero fill out of bounds) ---- */
typedef struct aclgan_tmap_spec {
    uint64_t base;       /* device address of element (0,...,0) */
    uint32_t rank;       /* 2..5 */
    uint32_t elem_bytes; /* 2 */
    uint64_t dims[5];    /* extent per dimension, innermost first */
    uint64_t strides[5]; /* byte stride per dimension (strides[0] == elem_bytes) */
    uint32_t box[5];     /* box extent per dimension */
} aclgan_tmap_spec;

/* ---- epilogue / output description shared by the implicit-GEMM kernels ---- */
typedef struct aclgan_out_spec {
    uint64_t ptr[2];        /* plane pointers (bf16 hi[/lo]) or fp32 buffer in ptr[0] */
    int32_t kind;           /* ACLGAN_OUT_* */
    int32_t act;            /* ACLGAN_ACT_* applied after the bias */
    float slope;            /* LeakyReLU negative slope */
    int32_t mirror;         /* reflect-pad width to replicate border pixels into (0 = none) */
    int64_t off;            /* element offset of logical pixel (n=0, y=0, x=0), channel 0 */
    int64_t sn, sy, sx, sc; /* element strides */
    int32_t N, H, W, C;     /* valid logical extent: rows outside are not stored; channels >= C dropped */
    uint64_t bias;          /* fp32 [bias_n] or 0 */
    int32_t bias_n;         /* channels >= bias_n get no bias (stored padding channels) */
    uint64_t stats;         /* double [N][C][2] (sum, sum of squares of the stored values) accumulated with atomics, or 0:
                               the statistics pass of InstanceNorm / AdaIN / LayerNorm fused into the conv epilogue */
    /* ---- sub-pixel up-convolution (csrc/upconv.cu): the four output phases of nearest-2x-upsample + 5x5 are folded into N ---- */
    int32_t d2s_c;          /* 0, or channels per phase: GEMM column ch = phase*d2s_c + c is stored at pixel offset
                               (phase>>1)*d2s_sy + (phase&1)*d2s_sx, channel c (depth-to-space; sy / sx then step 2 output pixels) */
    int32_t ring;           /* 1: rows on the 1-pixel border ring of the (H, W) grid are neither stored nor counted in stats
                               (the ring is recomputed exactly by the strip convolutions) */
    int64_t d2s_sy, d2s_sx;
    int32_t z_mod;          /* 0, or images per strip side: batch index z addresses image z % z_mod at offset (z / z_mod) * z_off */
    int32_t stats_c;        /* 0 (= C), or the channel stride of `stats` when it is shared with a wider launch */
    int64_t z_off;
} aclgan_out_spec;

/* ---- implicit GEMM plan:  D[pixel][n] = sum_seg sum_tap sum_chunk A_tap[pixel][64] * B[n][k(tap,chunk) + 64] ---- */
typedef struct aclgan_igemm_plan {
    aclgan_tmap_spec a[2][ACLGAN_MAX_AVARIANTS]; /* [plane][variant] rank-4 maps (c, x, y, z) */
    aclgan_tmap_spec b[2];                       /* [plane] rank-2 maps (k, n) */
    int32_t planes;      /* 1 | 2 */
    int32_t nseg;        /* 1 | 3 */
    int32_t n_avariants; /* maps used in a[.][] */
    int32_t block_n;     /* UMMA N (16,32,64,128,256) */
    int32_t n_tiles;     /* tiles along n */
    int32_t box_x, box_y, box_z; /* A box: box_x*box_y*box_z == 128 pixels */
    int32_t tiles_x, tiles_y, tiles_z;
    int32_t cchunks;     /* 64-wide K chunks per tap */
    int32_t num_taps;
    int32_t tap_dx[ACLGAN_MAX_TAPS];
    int32_t tap_dy[ACLGAN_MAX_TAPS];
    int32_t tap_var[ACLGAN_MAX_TAPS]; /* which a[.][variant] map the tap loads through */
    int32_t tap_bk[ACLGAN_MAX_TAPS];  /* first k element of the tap in B */
    int32_t flat;        /* 0: tile rows decode as (x,y,z) box coordinates; 1: q = x0 + row, decoded with pitches */
    int32_t flat_w;      /* flat: x = q % flat_w */
    int32_t flat_img;    /* flat: z = q / flat_img, y = (q % flat_img) / flat_w */
    int32_t n_groups;    /* 1, or the 4 output-parity phases of a stride-2 data gradient merged into one launch */
    int32_t group_taps;  /* taps per group: group g uses taps [g*group_taps, (g+1)*group_taps) */
    int64_t group_off[4];/* element offset added to out.off for the rows of group g */
    /* segment mode (stride-1 convolutions): the A rows of the k taps of ONE filter row are consecutive pixels, so
     * they are staged ONCE per 64-channel chunk as a segment of seg_rows = 128 + k - 1 (rounded up to 8) pixels and
     * tap j of the segment reads rows [tap_row, tap_row + 128) of it - k times less activation traffic from L2.
     * Tap t belongs to segment t / seg_taps.  Executed by the segment kernel; the per-tap fields above describe the
     * same computation for the plain kernel. */
    int32_t seg_mode;    /* 0 | 1 */
    int32_t seg_rows;    /* pixels per staged segment (multiple of 8, <= 256) */
    int32_t num_segs;    /* segments per 64-channel chunk (= filter rows) */
    int32_t seg_taps;    /* taps per segment (= filter columns) */
    int32_t seg_dx[16], seg_dy[16]; /* box offset of segment s relative to the tile origin (like tap_dx / tap_dy) */
    int32_t tap_row[ACLGAN_MAX_TAPS]; /* first segment row of tap t */
    aclgan_tmap_spec a_seg[2];        /* [plane] the variant-0 map with a seg_rows-pixel box */
    /* fold mode (forward of the few-output-channel final conv; default, env ACLGAN_FOLD=0 disables): the k taps of a filter row
     * are folded into the N dimension (weight row n = kw*8 + co, one tap per filter ROW), tiles are 128 flattened positions
     * stepping by tile_step = 128 - 8, and the epilogue sums the diagonal out[q][co] = sum_kw P[q + kw][kw*8 + co]. */
    int32_t fold;        /* 0, or the filter width k */
    int32_t tile_step;   /* flattened positions between consecutive tiles (fold mode) */
    aclgan_out_spec out;
} aclgan_igemm_plan;

/* ---- weight-gradient plan:  acc[m][n] = sum_pixels Mop[pixel][m] * Nop_tap[pixel][n]   (both operands MN-major:
 *      a smem row is one pixel, 64 channels wide), accumulated with fp32 atomics into dw[m][tap][n] ---- */
typedef struct aclgan_wgrad_plan {
    aclgan_tmap_spec mop[2][ACLGAN_MAX_AVARIANTS]; /* [plane][variant] rank-4 (c, x, y, z); box = 64 ch x 64 pixels */
    aclgan_tmap_spec nop[2][ACLGAN_MAX_AVARIANTS];
    int32_t planes, nseg;        /* 1/1 or 2/3: (m_hi,n_hi), (m_hi,n_lo), (m_lo,n_hi) */
    int32_t n_mvariants, n_nvariants;
    int32_t m_chunks;            /* 64-channel chunks of the M operand per tile (1 | 2); UMMA M is always 128 */
    int32_t n_chunks;            /* 64-channel chunks of the N operand per tile (1..4);  UMMA N = 64 * n_chunks */
    int32_t m_tiles, n_tiles;
    int32_t box_x, box_y, box_z; /* pixel box (product 64 or 128 = pixels reduced per pipeline stage) */
    int32_t blocks_x, blocks_y, blocks_z; /* pixel blocks covering the reduction grid */
    int32_t ksplit;              /* CTAs per (tap, m_tile, n_tile), each reducing a contiguous range of pixel blocks */
    int32_t num_taps;
    int32_t m_dx[ACLGAN_MAX_TAPS], m_dy[ACLGAN_MAX_TAPS], m_var[ACLGAN_MAX_TAPS];
    int32_t n_dx[ACLGAN_MAX_TAPS], n_dy[ACLGAN_MAX_TAPS], n_var[ACLGAN_MAX_TAPS];
    int32_t tap_out[ACLGAN_MAX_TAPS]; /* tap slot in dw */
    uint64_t dw;                 /* fp32, element (m, tap, n) at m*dw_sm + tap*dw_st + n */
    int64_t dw_sm, dw_st;
    int32_t M, Nn;               /* valid rows / columns */
    /* segment mode (stride-1 convolutions whose output rows are a multiple of 64 pixels): one CTA reduces ALL seg_taps
     * taps of one filter ROW (num_taps = filter rows): per 64-pixel block of an output row the dY tile is staged once and
     * the conv-input pixels once as a segment of seg_rows >= 64 + seg_taps - 1 pixels; tap kw multiplies the dY tile
     * with rows [kw, kw + 64) of the segment into its own accumulator (seg_taps * 64 * n_chunks <= 512 TMEM columns).
     * The shifted operand is N when seg_on_m == 0, else M; its box-load map is seg_map.  A filter row may be split into
     * several entries (so that the accumulators fit TMEM with 128-column tiles): entry t covers the seg_cnt[t] taps
     * kw = seg_kw0[t] .. of filter row m_dy/n_dy[t]; output tap slot of its j-th tap = tap_out[t] + j. */
    int32_t seg_mode, seg_rows, seg_taps, seg_on_m;
    int32_t seg_kw0[ACLGAN_MAX_TAPS], seg_cnt[ACLGAN_MAX_TAPS];
    aclgan_tmap_spec seg_map[2];
    /* vertical segments (small-channel window layers, where a tap is a filter ROW): the 64-pixel block is box_x x box_y pixels
     * (16 x 4), the segment box_x x (box_y + k - 1) pixels, and tap kh starts seg_step = box_x rows further into it (0 = 1) */
    int32_t seg_step;
    int32_t pad_;
} aclgan_wgrad_plan;

/* ---- padded NHWC activation handle ---- */
typedef struct aclgan_act {
    uint64_t data[2]; /* plane pointers */
    int32_t planes;   /* 1 | 2 */
    int32_t n, h, w;  /* logical extent */
    int32_t c;        /* stored channels per pixel (multiple of 8) */
    int32_t pad;      /* stored border width */
} aclgan_act;

/* ---- convolution descriptors ---- */
typedef struct aclgan_conv_desc {
    int32_t cin, cout;   /* logical channels (reference nn.Conv2d in/out channels, networks.py:363) */
    int32_t k;           /* square kernel size */
    int32_t stride;      /* 1 | 2 */
    int32_t pad;         /* reflect padding applied by the block (networks.py:318-325); stored in the input plane */
    int32_t window;      /* ACLGAN_WINDOW_*: small-C side handled as a 64-wide pixel window */
} aclgan_conv_desc;

int aclgan_version(void);
const char* aclgan_build_info(void);

/* plan builders: pure host code, usable without a GPU (unit-tested by CPU emulation) */
int aclgan_plan_conv_fwd(const aclgan_conv_desc* cd, const aclgan_act* x, const uint64_t w[2],
                         const aclgan_out_spec* out, aclgan_igemm_plan* plan);
/* stride 2: phase = 0..3 builds one output-parity phase (out describes that phase's strided view); phase = -1 merges
 * all four into one launch (out describes phase 0; the other phases are reached through group_off) */
int aclgan_plan_conv_dgrad(const aclgan_conv_desc* cd, const aclgan_act* dy, const uint64_t wt[2], int phase,
                           const aclgan_out_spec* out, aclgan_igemm_plan* plan);
int aclgan_plan_conv_wgrad(const aclgan_conv_desc* cd, const aclgan_act* dy, const aclgan_act* x, uint64_t dw,
                           aclgan_wgrad_plan* plan);
/* packed-weight geometry (elements): rows x k_total of the forward (K-major over cin) and
 * transposed (K-major over cout) packings the plans above expect */
int aclgan_packed_weight_shape(const aclgan_conv_desc* cd, int transposed, int64_t* rows, int64_t* k_total);
/* index of logical weight element W[co][ci][kh][kw] inside the packed buffers; -1 if not stored */
int64_t aclgan_packed_weight_index(const aclgan_conv_desc* cd, int transposed, int co, int ci, int kh, int kw);
/* the fp32 weight-gradient buffer written by the wgrad plan has exactly the layout of one of the two packings:
 * returns 0 (forward packing) or 1 (transposed packing) */
int aclgan_wgrad_layout(const aclgan_conv_desc* cd);

/* launches (networks.py:363,366 Conv2d forward; autograd of it for dgrad / wgrad) */
int aclgan_igemm_launch(const aclgan_igemm_plan* plan, void* stream);
int aclgan_fold_launch(const aclgan_igemm_plan* plan, int repeat, void* stream);   /* fold-mode plans (igemm_launch forwards them) */
/* 1 when the epilogue of this plan can accumulate out.stats (otherwise run aclgan_norm_stats on the output) */
int aclgan_igemm_stats_supported(const aclgan_igemm_plan* plan);
int aclgan_wgrad_launch(const aclgan_wgrad_plan* plan, void* stream);
/* same plan launched `repeat` times back to back (tensor maps encoded once): device-side kernel timing */
int aclgan_igemm_launch_repeat(const aclgan_igemm_plan* plan, int repeat, void* stream);
int aclgan_wgrad_launch_repeat(const aclgan_wgrad_plan* plan, int repeat, void* stream);


/* ================= element-wise / reduction kernels around the convolutions ================= */

enum { ACLGAN_NORM_NONE = 0, ACLGAN_NORM_IN = 1, ACLGAN_NORM_ADAIN = 2, ACLGAN_NORM_LN = 3 };
enum { ACLGAN_MASK_NONE = 0, ACLGAN_MASK_FROM_Z = 1, ACLGAN_MASK_FROM_OUT = 2 };

/* dense NHWC tensor (raw conv output / gradient), bf16 (kind 0) or fp32 (kind 1) */
typedef struct aclgan_tensor4 {
    uint64_t ptr;
    int32_t kind;
    int32_t n, h, w, c;
} aclgan_tensor4;

/* NCHW fp32 image(s) -> reflect-padded NHWC planes with 8|16 stored channels (reference: images.cuda(),
 * train.py:67, nn.ReflectionPad2d of the first Conv2dBlock networks.py:319, torch.cat of the dis_2 pair
 * trainer.py:132-133,279-280) */
typedef struct aclgan_pack_img_args {
    uint64_t src0, src1;   /* fp32 NCHW; src1 optional (second image of a concatenated pair) */
    int32_t c0, c1;        /* channels taken from src0 / src1 */
    int32_t n, h, w;
    aclgan_act dst;        /* dst.c in {8,16}, dst.pad = reflect width */
} aclgan_pack_img_args;
int aclgan_pack_img(const aclgan_pack_img_args* a, void* stream);

/* generic layout conversion at the API boundary of the inference path (AdaINGen.encode returns / decode takes the content
 * code as an NCHW fp32 tensor: reference networks.py:141-152, test.py:96-106): fp32 NCHW [n][c][h][w] -> reflect-padded NHWC
 * plane(s) with dst->c >= c stored channels (zero beyond c), and back (interior, first c channels, hi + lo) */
int aclgan_pack_nchw(uint64_t src, int32_t c, const aclgan_act* dst, void* stream);
int aclgan_unpack_plane(const aclgan_act* src, int32_t c, uint64_t dst, void* stream);

/* per-(n,c) sum and sum of squares of a raw conv output, accumulated in fp64 (first half of
 * nn.InstanceNorm2d / F.batch_norm / LayerNorm statistics: networks.py:333,499-501,525-529) */
int aclgan_norm_stats(const aclgan_tensor4* y, uint64_t sums /* double [n][c][2], zeroed by the caller */, void* stream);

/* statistics -> per-(n,c) scale / shift (+ saved mean / inverse deviation for the backward pass).
 *   IN    : scale = rstd, shift = -mean*rstd                       (biased var, eps inside sqrt)
 *   ADAIN : scale = rstd*w[n,c], shift = b[n,c] - mean*scale       (networks.py:497-503)
 *   LN    : per sample over C*H*W, UNBIASED std, eps outside: scale = g[c]/(std+eps)  (networks.py:525-535) */
typedef struct aclgan_norm_finalize_args {
    int32_t mode, n, c, hw;  /* c = stored channels */
    int32_t c_valid;         /* logical channels (<= c); stored padding channels get scale = shift = 0 */
    float eps;
    uint64_t sums;          /* double [n][c][2] */
    uint64_t w, b;          /* ADAIN: fp32 [n][c_valid]; LN: fp32 [c_valid] gamma / beta; IN: unused */
    uint64_t scale, shift;  /* out fp32 [n][c] */
    uint64_t mean, inv;     /* out fp32 [n][c] (LN: broadcast per sample) */
    uint64_t sigma;         /* out fp32 [n] (LN only: the unbiased std) */
    int64_t wb_stride;      /* ADAIN: elements between consecutive samples of w / b (0 = c_valid): the parameters are read
                               in place from the MLP output row (networks.py:154-163 slices, never copied) */
    int32_t stat_groups;    /* LN after a sub-pixel up-convolution: `sums` holds stat_groups (4) phase groups of c_valid channels
                               per sample, i.e. [n][stat_groups * c_valid][2]; 0 / 1 = plain [n][c][2] */
    int32_t pad_;
} aclgan_norm_finalize_args;
int aclgan_norm_finalize(const aclgan_norm_finalize_args* a, void* stream);

/* out = act(y*scale + shift) (+ residual), written as the NEXT conv's reflect-padded (optionally 2x nearest
 * upsampled) input plane(s)  (networks.py:367-370 norm+activation, :309 residual add, :256 nn.Upsample,
 * :319 ReflectionPad2d of the consumer) */
typedef struct aclgan_apply_args {
    aclgan_tensor4 y;
    uint64_t scale, shift;  /* fp32 [n][c], or 0 for identity */
    int32_t act;
    float slope;
    int32_t has_res;
    aclgan_act res;         /* residual source plane (same n,h,w,c), any pad */
    int32_t upsample;       /* 1 | 2 */
    aclgan_act dst;         /* n, h*upsample, w*upsample, c; pad = consumer's reflect width */
} aclgan_apply_args;
int aclgan_norm_apply(const aclgan_apply_args* a, void* stream);
/* aclgan_norm_finalize + aclgan_norm_apply (a->scale / a->shift == f->scale / f->shift) as ONE launch whenever the row-structured
 * apply kernel takes the plane: every CTA derives the coefficients of its own channels from the statistics, CTA 0 of each image
 * stores them (and mean / inv / sigma) for the backward pass; otherwise the two launches */
int aclgan_norm_finalize_apply(const aclgan_norm_finalize_args* f, const aclgan_apply_args* a, void* stream);

/* backward of  pad/upsample -> activation -> norm  for one block.
 * g = fold(gp) (+ gr): gradient w.r.t. the block's logical output, gathered from the gradient of the padded
 * (and upsampled) plane; dz = g * act'(z);  norm blocks: dy = ca*dz + cb*yhat + cc with yhat = (y-mean)*inv.
 * `reduce` produces T1 = sum dz, T2 = sum dz*yhat per (n,c) in fp64; `apply` writes dy as zero-bordered plane(s). */
typedef struct aclgan_block_bwd_args {
    uint64_t gp;            /* gradient of the padded plane [n][u*h+2p][u*w+2p][c] or 0 */
    int32_t g_kind;         /* 0 bf16, 1 fp32 (gp and gr) */
    int32_t gp_pad, upsample;
    uint64_t gr;            /* dense [n][h][w][c] additional gradient (residual branch) or 0 */
    int32_t n, h, w, c;
    int32_t mask_mode;      /* ACLGAN_MASK_* */
    float slope;
    aclgan_tensor4 y;       /* raw conv output (MASK_FROM_Z / norm blocks) */
    uint64_t scale, shift;  /* fp32 [n][c] used in the forward pass */
    aclgan_act out;         /* forward output plane (MASK_FROM_OUT) */
    int32_t norm;           /* 0 | 1 */
    uint64_t mean, inv;     /* fp32 [n][c] */
    uint64_t sums;          /* double [n][c][2]: T1, T2 (written by reduce) */
    uint64_t ca, cb, cc;    /* fp32 [n][c] (read by apply when norm) */
    aclgan_act dy;          /* apply output, dy.pad = zero border */
    uint64_t dbias;         /* apply, blocks without norm: fp32 [dbias_n] += sum over (n,h,w) of dz (conv bias gradient), or 0 */
    int32_t dbias_n;
} aclgan_block_bwd_args;
int aclgan_block_bwd_reduce(const aclgan_block_bwd_args* a, void* stream);
int aclgan_block_bwd_apply(const aclgan_block_bwd_args* a, void* stream);

/* T1/T2 -> (ca, cb, cc) and the parameter gradients of the norm layer */
typedef struct aclgan_norm_bwd_finalize_args {
    int32_t mode, n, c, hw;
    int32_t c_valid;
    uint64_t sums;          /* double [n][c][2] */
    uint64_t inv, sigma;    /* forward saves */
    uint64_t w;             /* ADAIN: fp32 [n][c] weight; LN: fp32 [c] gamma */
    uint64_t ca, cb, cc;    /* out fp32 [n][c] */
    uint64_t dw, db;        /* out: ADAIN fp32 [n][c] (assigned); LN fp32 [c] (accumulated: +=) ; IN unused */
    int64_t wb_stride;      /* ADAIN: sample stride of w, dw and db in elements (0 = c_valid) */
    /* LN only: gradient of the conv bias in front of the LayerNorm (NOT cancelled: the statistics span all channels):
     * dbias[c] += sum_n ca*T1 + cb*sum_hw(yhat) + cc*HW with sum_hw(yhat) = (S1fwd - HW*mean)*inv; 0 = skip */
    uint64_t fsums;         /* double [n][c][2]: the FORWARD statistics (S1 = sum of y) */
    uint64_t mean;          /* fp32 [n][c] forward mean */
    uint64_t dbias;         /* fp32 [c_valid] accumulated with atomics, or 0 */
    int32_t fstat_groups;   /* fsums is [n][fstat_groups * c_valid][2] (phase groups of the sub-pixel up-convolution); 0 / 1 = [n][c][2] */
    int32_t pad_;
} aclgan_norm_bwd_finalize_args;
int aclgan_norm_bwd_finalize(const aclgan_norm_bwd_finalize_args* a, void* stream);
/* aclgan_norm_bwd_finalize + aclgan_block_bwd_apply (same sums / ca / cb / cc) as ONE launch whenever the row-structured apply
 * kernel takes the plane: every CTA derives (ca, cb, cc) of its own channels from T1 / T2, CTA 0 of each image performs the
 * parameter-gradient side effects; otherwise the two launches */
int aclgan_norm_bwd_finalize_apply(const aclgan_norm_bwd_finalize_args* f, const aclgan_block_bwd_args* a, void* stream);

/* gradient of an NCHW fp32 image (optionally through tanh: d * (1 - out^2)) -> zero-bordered 8-channel planes
 * feeding the final conv's dgrad / wgrad; also accumulates the bias gradient (sum over n,h,w) */
typedef struct aclgan_img_grad_pack_args {
    uint64_t dimg, out_img; /* fp32 NCHW [n][c][h][w]; out_img = tanh output or 0 */
    int32_t n, c, h, w;
    aclgan_act dy;          /* dy.c = 8 */
    uint64_t dbias;         /* fp32 [c], accumulated */
} aclgan_img_grad_pack_args;
int aclgan_img_grad_pack(const aclgan_img_grad_pack_args* a, void* stream);

/* first-layer dgrad output (fp32 [n][h+2p][w+2p][cs], gradient of the padded image plane) -> NCHW fp32 image
 * gradient with the reflect padding folded back (adjoint of networks.py:319) */
typedef struct aclgan_img_grad_unpack_args {
    uint64_t src;
    int32_t n, c, h, w, cs, pad;
    int32_t c_off;          /* first stored channel taken from src (second image of a concatenated pair) */
    uint64_t dst;           /* fp32 NCHW [n][c][h][w] */
    int32_t accumulate;
} aclgan_img_grad_unpack_args;
int aclgan_img_grad_unpack(const aclgan_img_grad_unpack_args* a, void* stream);

/* OIHW fp32 master weight -> packed bf16 plane(s); the packed index is affine in (co,ci,kh,kw) */
typedef struct aclgan_pack_weight_args {
    uint64_t w;
    int32_t co, ci, kh, kw;
    int64_t base, s_co, s_ci, s_kh, s_kw;
    uint64_t dst[2];
    int32_t planes;
} aclgan_pack_weight_args;
int aclgan_pack_weight(const aclgan_pack_weight_args* a, void* stream);


/* ---- fused multi-tensor Adam (torch.optim.Adam semantics with L2 weight decay: reference trainer.py:39-42,170,293)
 *      + re-derivation of the packed bf16 weight planes in the same pass.  One table entry per parameter tensor. */
typedef struct aclgan_adam_tensor {
    uint64_t p, m, v;        /* fp32 contiguous: parameter (OIHW / dense), exp_avg, exp_avg_sq */
    uint64_t g;              /* fp32 gradient base pointer */
    int32_t d[4];            /* logical dims (co, ci, kh, kw); lower-rank tensors are left-padded with 1 */
    int64_t gs[4];           /* gradient element strides per dim (the arena keeps conv grads in packed layout) */
    int64_t goff;            /* gradient element offset */
    uint64_t pk[2][2];       /* [packing: forward, transposed][plane] bf16 destinations or 0 */
    int64_t aff[2][5];       /* [packing] base, s_co, s_ci, s_kh, s_kw */
    int32_t planes;
    int32_t pad_;
} aclgan_adam_tensor;

/* hyper (device, fp32[8]): lr, beta1, beta2, eps, weight_decay, grad_scale, step (advanced on device), unused.
 * chunks (device, int32[2*n_chunks]): (tensor id, unit) per CTA, unit = 0 .. aclgan_adam_units(tensor) - 1.  A unit is a
 * (32 co x 32 ci x all taps) tile of a conv weight - gradient, master / moments and each packed plane are then streamed in their
 * own fastest order through a shared-memory transpose - or 1024 consecutive elements of a dense tensor. */
int aclgan_adam_units(const aclgan_adam_tensor* t);   /* host-side: CTAs needed for this table entry */
int aclgan_adam_step(uint64_t table, uint64_t chunks, int32_t n_chunks, uint64_t hyper, void* stream);
int aclgan_adam_advance(uint64_t hyper, void* stream);

/* ================= sub-pixel up-convolution (csrc/upconv.cu) =================
 * nearest-2x-upsample -> ReflectionPad2d(2) -> Conv2d 5x5 (reference networks.py:256-257) as ONE 3x3 convolution of the
 * reflect-pad-1 source plane with 4*Cout output channels (the 4 output phases folded into N, depth-to-space epilogue:
 * aclgan_out_spec.d2s_*, .ring) plus exact 5x5 convolutions on four thin border strips (aclgan_out_spec.z_mod / z_off).
 * The GEMMs are ordinary plans; these entry points are the layout kernels around them. */
typedef struct aclgan_up_derive_args {
    uint64_t w5, bias;        /* fp32 [co][ci][5][5] master weight, fp32 [co] bias */
    int32_t co, ci, planes, pad_;
    uint64_t pk[2][2];        /* [forward, transposed packing][plane] bf16 destinations of the phase weights (4co, ci, 3, 3):
                                 Wp[(py*2+px)*co + o][i][u][v] = sum of W5[o][i][a][b] over the taps a (b) that read source row (column)
                                 offset u-1 (v-1) in phase py (px): phase 0: {0,1},{2,3},{4}; phase 1: {0},{1,2},{3,4} */
    int64_t aff[2][5];        /* packed index = base + co'*s0 + ci*s1 + u*s2 + v*s3 (aclgan_packed_weight_index of the 3x3 desc) */
    uint64_t bias4;           /* out fp32 [4*co]: the bias tiled over the four phases, or 0 */
} aclgan_up_derive_args;
int aclgan_up_derive_weights(const aclgan_up_derive_args* a, void* stream);

/* forward: strips of the exactly padded up-sampled plane.  rows: [2n][2][2W] pad 2 = up-sampled rows -2..3 (side 0) and
 * 2H-4..2H+1 (side 1), columns -2..2W+1; cols: [2n][2][2H-4] pad 2, TRANSPOSED (strip row = up-sampled column -2..3 /
 * 2W-4..2W+1, strip column = up-sampled row 0..2H-1): both fully written, borders included */
typedef struct aclgan_up_strips_args {
    aclgan_act src;           /* source plane [n][H][W][C] (any pad; only the interior is read) */
    aclgan_act rows, cols;
} aclgan_up_strips_args;
int aclgan_up_gather_strips(const aclgan_up_strips_args* a, void* stream);

/* backward: dense dY [n][2H][2W][cs] (pad 0) -> (a) space-to-depth plane [n][H][W][4*cout] pad 2 with the ring source pixels
 * and the border zero, (b) ring rows [2n][2][2W] pad 4, (c) ring columns [2n][2][2H-4] pad 4 transposed (zero borders) */
typedef struct aclgan_up_dy_pack_args {
    aclgan_act dy;
    int32_t cout, pad_;       /* valid channels per phase (multiple of 8) */
    aclgan_act s2d, rows, cols;
} aclgan_up_dy_pack_args;
int aclgan_up_dy_pack(const aclgan_up_dy_pack_args* a, void* stream);

/* backward: input gradients of the strip convolutions, gathered back onto the source pixels they were copied from and added
 * to the interior of the padded-source-plane gradient g [n][H+2][W+2][C] */
typedef struct aclgan_up_scatter_args {
    uint64_t g, grows, gcols; /* grows [2n][6][2W+4][C], gcols [2n][6][2H][C]; all of `kind` (0 bf16, 1 fp32) */
    int32_t kind, n, h, w, c, pad_;
} aclgan_up_scatter_args;
int aclgan_up_scatter_strips(const aclgan_up_scatter_args* a, void* stream);

/* backward: dW5[o][i][a][b] += sum over the 4 phases of dWp[phase*co + o][i][u(py, a)][v(px, b)] (both in packed layouts) */
typedef struct aclgan_up_fold_wgrad_args {
    uint64_t dwp;
    int64_t affp[5];
    uint64_t dw5;
    int64_t aff5[5];
    int32_t co, ci;
} aclgan_up_fold_wgrad_args;
int aclgan_up_fold_wgrad(const aclgan_up_fold_wgrad_args* a, void* stream);

/* ================= small operators: csrc/smallops.cu ================= */

/* AvgPool2d(3, stride 2, padding 1, count_include_pad=False) of the multi-scale discriminator's image pyramid on NCHW fp32
 * images (reference networks.py:33,53).  Output extent (h - 1) / 2 + 1.  Bit-identical to the ATen kernel: the window is
 * summed row-major in fp32 and divided once by the number of valid elements (4 | 6 | 9). */
typedef struct aclgan_avgpool_args {
    uint64_t src, dst;   /* fwd: src [planes][h][w] -> dst [planes][ho][wo];  bwd: src = d(dst) [planes][ho][wo] -> dst = d(src) [planes][h][w] */
    int32_t planes;      /* n * c */
    int32_t h, w;        /* extent of the un-pooled image */
    int32_t accumulate;  /* bwd: dst += */
} aclgan_avgpool_args;
int aclgan_avgpool3x3s2_fwd(const aclgan_avgpool_args* a, void* stream);
int aclgan_avgpool3x3s2_bwd(const aclgan_avgpool_args* a, void* stream);

/* style head: AdaptiveAvgPool2d(1) + Conv2d(C, style_dim, 1) (networks.py:222-223) and its backward */
typedef struct aclgan_style_head_args {
    aclgan_act x;             /* last style-encoder plane */
    int32_t c_valid, style_dim;
    uint64_t weight, bias;    /* fp32 [style_dim][c_valid], [style_dim] */
    uint64_t pooled;          /* fp32 [n][c_valid]: written by fwd, read by bwd */
    uint64_t style;           /* fwd out: fp32 [n][style_dim] */
    uint64_t dstyle;          /* bwd in:  fp32 [n][style_dim] */
    uint64_t dweight, dbias;  /* bwd: accumulated (+=), or 0 */
    uint64_t gr;              /* bwd out: dense gradient of the plane [n][h][w][x.c] (assigned) */
    int32_t g_kind;           /* 0 bf16, 1 fp32 */
    int32_t pad_;
} aclgan_style_head_args;
int aclgan_style_head_fwd(const aclgan_style_head_args* a, void* stream);
int aclgan_style_head_bwd(const aclgan_style_head_args* a, void* stream);

/* MLP of Linear(+ReLU) layers, ReLU on all but the last (networks.py:280-292: 8 -> 256 -> 256 -> n_adain) */
typedef struct aclgan_mlp_args {
    int32_t n, n_layers;      /* samples, layers (<= 4) */
    int32_t dims[5];          /* in, hidden ..., out */
    int32_t pad_;
    uint64_t w[4], b[4];      /* fp32 [dims[l+1]][dims[l]], [dims[l+1]] */
    uint64_t h[5];            /* fp32 activations [n][dims[l]]: h[0] input, h[n_layers] output (post-ReLU values for hidden layers) */
    int64_t h0_stride;        /* row stride of h[0] in elements (0 = dims[0]) */
    uint64_t dh[5];           /* bwd: dh[n_layers] = gradient of the output (in); dh[l] scratch [n][dims[l]] for 0 < l < n_layers;
                                 dh[0] = gradient of the input (out) or 0 */
    uint64_t dw[4], db[4];    /* bwd: accumulated (+=), or 0 */
} aclgan_mlp_args;
int aclgan_mlp_fwd(const aclgan_mlp_args* a, void* stream);
int aclgan_mlp_bwd(const aclgan_mlp_args* a, void* stream);

/* discriminator head Conv2d(C, 1, 1) (networks.py:45) fused with the GAN terms of networks.py:60-106 and their gradient seed.
 * gan_kind ACLGAN_GAN_LSGAN: mean((o - t)^2) (networks.py:67,83,98); ACLGAN_GAN_NSGAN: mean of
 * F.binary_cross_entropy(F.sigmoid(o), t) with t in {0, 1} (networks.py:68-72,84-86,99-103; fp32 sigmoid, log clamped at -100 and
 * the 1e-12 floor of its backward as in PyTorch).  The batch holds `groups` image groups of equal size (one discriminator
 * evaluated once over the concatenation of its inputs); group g has target[g], loss weight gweight[g] and loss accumulator
 * slot loss_slot[g]. */
enum { ACLGAN_GAN_LSGAN = 0, ACLGAN_GAN_NSGAN = 1 };
typedef struct aclgan_dis_head_args {
    aclgan_act x;             /* last feature plane */
    int32_t c_valid, groups;  /* groups <= 4 */
    uint64_t weight, bias;    /* fp32 [c_valid], [1] */
    uint64_t logits;          /* out fp32 [N][h][w] */
    uint64_t dlogits;         /* out fp32 [N][h][w] = gweight[g] * d term / d o / (n_per * h * w)  (LSGAN: 2 (o - target[g])), or 0 */
    uint64_t loss;            /* double accumulators: loss[loss_slot[g]] += mean over the group of the GAN term, or 0 */
    float target[4], gweight[4];
    int32_t loss_slot[4];
    int32_t gan_kind, pad_;   /* ACLGAN_GAN_* */
} aclgan_dis_head_args;
int aclgan_dis_head_fwd(const aclgan_dis_head_args* a, void* stream);
typedef struct aclgan_dis_head_bwd_args {
    aclgan_act x;
    int32_t c_valid, g_kind;
    uint64_t weight, dlogits;
    uint64_t dweight, dbias;  /* accumulated with atomics, or 0 (generator update: no discriminator weight gradients) */
    uint64_t gr;              /* out: dense gradient of the plane [N][h][w][x.c] (assigned), or 0 */
} aclgan_dis_head_bwd_args;
int aclgan_dis_head_bwd(const aclgan_dis_head_bwd_args* a, void* stream);

/* focus_translation (trainer.py:85-88): dst = fg * m + bg * (1 - m), m = (mask + 1) / 2, fg / mask = channels 0-2 / 3 of the
 * decoder output; same rounding sequence as the reference's fp32 expression.  bwd: adjoint into d(out4) and d(bg). */
typedef struct aclgan_blend_args {
    uint64_t out4;            /* fp32 [n][4][h][w] */
    uint64_t bg;              /* fp32 [n][3][h][w] */
    uint64_t dst;             /* fwd out fp32 [n][3][h][w] */
    int32_t n, h, w;
    int32_t acc_out4, acc_bg; /* bwd: 1 = accumulate (+=), 0 = assign */
    int32_t pad_;
    uint64_t ddst;            /* bwd in  [n][3][h][w] */
    uint64_t dout4;           /* bwd out [n][4][h][w] or 0 */
    uint64_t dbg;             /* bwd out [n][3][h][w] or 0 */
} aclgan_blend_args;
int aclgan_focus_blend_fwd(const aclgan_blend_args* a, void* stream);
int aclgan_focus_blend_bwd(const aclgan_blend_args* a, void* stream);

/* loss reductions into double accumulators (zeroed by the caller once per update) */
enum { ACLGAN_LOSS_L1 = 0, ACLGAN_LOSS_FOCUS = 1 };
typedef struct aclgan_loss_reduce_args {
    int32_t mode;
    int32_t n, ca, c, h, w;   /* a is [n][ca][h][w]; L1 compares its first c channels with b [n][c][h][w]; FOCUS reads channel 3 */
    uint64_t a, b;
    uint64_t acc;             /* double[]: L1: acc[slot] += mean|a - b| (trainer.py:61-62);
                                 FOCUS: acc[slot] += sum(m - upper), acc[slot+1] += sum(lower - m), acc[slot+2] += sum 1/(|m-.5|+eps)
                                 over the WHOLE batch, m = (a[:,3] + 1)/2 (trainer.py:146-151) */
    int32_t slot;
    int32_t acc_da;           /* L1: da (+)= */
    uint64_t da;              /* L1: [n][ca][h][w], channels < c get sign(a - b) * gscale, or 0 */
    float gscale, upper, lower, eps;
} aclgan_loss_reduce_args;
int aclgan_loss_reduce(const aclgan_loss_reduce_args* a, void* stream);

/* second pass of the focus losses: size = delta * (relu(S1)^2 + relu(S2)^2) -> sums[size_slot]; channel 3 of d(out4) (+)=
 * gscale * (2 delta (relu(S1) - relu(S2)) - sign(m - .5) / (|m - .5| + eps)^2)   (trainer.py:149-161) */
typedef struct aclgan_focus_grad_args {
    uint64_t out4, dout4;     /* fp32 [n][4][h][w] */
    int32_t n, h, w;
    int32_t slot, size_slot;  /* sums[slot], sums[slot+1] = S1, S2 (from ACLGAN_LOSS_FOCUS) */
    int32_t acc;
    uint64_t sums;            /* double[] */
    float delta, eps, gscale;
    int32_t pad_;
} aclgan_focus_grad_args;
int aclgan_focus_grad(const aclgan_focus_grad_args* a, void* stream);

/* out[j] = sum_k M[j][k] * acc[k]: every loss_* scalar and weighted total of trainer.py:142-165 / 288-290 in one launch */
int aclgan_loss_combine(uint64_t acc /* double[K] */, uint64_t M /* fp32 [J][K] */, uint64_t out /* fp32 [J] */, int32_t J,
                        int32_t K, void* stream);

/* input pipeline (reference utils.py:43-100): uint8 NHWC batch [n][h][w][3] -> fp32 NCHW [n][3][h][w] in [-1, 1] with an
 * optional per-sample horizontal flip (flip: uint8 [n] or 0) = RandomHorizontalFlip + ToTensor + Normalize(0.5, 0.5), bit-identical
 * to the torchvision ops (u8 -> float, / 255, - 0.5, / 0.5) */
int aclgan_augment_u8(uint64_t src, uint64_t flip, uint64_t dst, int32_t n, int32_t h, int32_t w, void* stream);
/* dst[ch] += sum_n sums[n][ch][0] for ch < c_valid: conv-bias gradient of a no-norm block from the T1 sums of
 * aclgan_block_bwd_reduce (planes too small for the fused-bias apply kernel) */
int aclgan_stats_to_bias(uint64_t sums /* double [n][c][2] */, uint64_t dst /* fp32 [c_valid] */, int32_t n, int32_t c,
                         int32_t c_valid, void* stream);
/* dst = alpha * a + beta * b (b may be 0; dst may alias a or b); kind 0 bf16, 1 fp32: gradient accumulation / alpha * z_2 */
int aclgan_axpby(uint64_t dst, uint64_t a, uint64_t b, float alpha, float beta, int64_t n, int32_t kind, void* stream);
/* cudaMemsetAsync(0) / device-to-device cudaMemcpyAsync on the caller's stream (memset / memcpy nodes under graph capture) */
int aclgan_zero(uint64_t ptr, int64_t bytes, void* stream);
int aclgan_copy(uint64_t dst, uint64_t src, int64_t bytes, void* stream);

/* stream-ordering helper for the host code: wait on an event recorded outside a stream capture (an external event-wait
 * node when `stream` is capturing); returns the cudaError_t */
int aclgan_stream_wait_external_event(void* stream, void* event);

#ifdef __cplusplus
}
#endif
#endif /* ACLGAN_B200_H */
