// Host-side geometry: turns a convolution description into an implicit-GEMM plan (tensor-map specs, filter-tap
// box offsets, tile decomposition, packed-weight indexing).  Pure host code - runs without a GPU, and is
// unit-tested on CPU by emulating the TMA box loads the plan prescribes (tests/test_plans.py).
//
// Geometry notes (all coordinates in the *stored* padded plane [n][H+2p][W+2p][C]):
//  fwd, stride 1 : out(y,x) tap(kh,kw) reads (y+kh, x+kw)            -> one map, box offset (kw, kh)
//  fwd, stride 2 : reads (2y+kh, 2x+kw) = parity map (kh&1,kw&1) at (y+kh/2, x+kw/2)
//  fwd, window   : small-C planes (C=8|16): 64 consecutive elements = 64/C pixels x C channels form one
//                  K chunk, so a whole filter row is one tap; weights are packed [kh][kw*C+ci] with zeros
//  dgrad         : dX_pad = full correlation of the zero-bordered dY with the flipped filter, evaluated over
//                  the flattened (pitch = stored dY width) grid so every tile is 128 consecutive pixels;
//                  stride 2 splits into 4 output-parity phases of 2x2 taps each.
#include <cstdlib>

#include "common.cuh"

namespace aclgan {

static int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}
static int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p *= 2;
    return p;
}

static void choose_block_n(int n_out, int* block_n, int* n_tiles) {
    int padded = round_up(n_out, 16);
    if (padded <= 256) {
        int b = 16;
        while (b < padded) b *= 2;
        *block_n = b;
        *n_tiles = 1;
    } else {
        *block_n = 256;
        *n_tiles = ceil_div(n_out, 256);
    }
}

static void plane_map4(aclgan_tmap_spec* m, uint64_t base, int c_extent, int64_t w_extent, int64_t h_extent, int n,
                       int64_t sx_bytes, int64_t sy_bytes, int64_t sn_bytes, int bx, int by, int bz) {
    memset(m, 0, sizeof(*m));
    m->base = base;
    m->rank = 4;
    m->elem_bytes = 2;
    m->dims[0] = c_extent;  m->dims[1] = w_extent;  m->dims[2] = h_extent;  m->dims[3] = n;
    m->strides[0] = 2;      m->strides[1] = sx_bytes; m->strides[2] = sy_bytes; m->strides[3] = sn_bytes;
    m->box[0] = 64;         m->box[1] = bx;          m->box[2] = by;          m->box[3] = bz;
}

static void weight_map2(aclgan_tmap_spec* m, uint64_t base, int64_t k_total, int64_t rows, int block_n) {
    memset(m, 0, sizeof(*m));
    m->base = base;
    m->rank = 2;
    m->elem_bytes = 2;
    m->dims[0] = k_total; m->dims[1] = rows;
    m->strides[0] = 2;    m->strides[1] = k_total * 2;
    m->box[0] = 64;       m->box[1] = block_n;
}

static void rect_tiles(aclgan_igemm_plan* p, int wo, int ho, int n) {
    int bx = pow2_floor(wo < 128 ? wo : 128);
    int by = 128 / bx;
    int hc = pow2_ceil(ho);
    if (by > hc) by = hc;
    int bz = 128 / (bx * by);
    p->box_x = bx; p->box_y = by; p->box_z = bz;
    p->tiles_x = ceil_div(wo, bx); p->tiles_y = ceil_div(ho, by); p->tiles_z = ceil_div(n, bz);
    p->flat = 0; p->flat_w = 1; p->flat_img = 1;
}

// stored channel count of the K dimension of a packing
// segment mode can be switched off for A/B comparisons (env ACLGAN_SEG=0): the plans then keep the box-per-tap geometry
static bool seg_enabled() {
    const char* e = getenv("ACLGAN_SEG");
    return e == nullptr || atoi(e) != 0;
}

// fold mode of the few-output-channel final conv's forward: on by default since round 2 (measured on the 256x256 bs 8 step:
// 43.12 -> 41.72 ms per step-pair, 256x256 step parity unchanged); ACLGAN_FOLD=0 switches back to one MMA group per tap
static bool fold_enabled(const aclgan_conv_desc* cd) {
    const char* e = getenv("ACLGAN_FOLD");
    return (e == nullptr || atoi(e) != 0) && cd->window == ACLGAN_WINDOW_OUT && cd->stride == 1 && cd->k <= 8 && cd->cout <= 8;
}

// vertical segments of the stride-1 small-channel window layers (ACLGAN_WINDOW_VSEG=0: one 128-pixel box per filter row)
static bool window_vseg_enabled() {
    const char* e = getenv("ACLGAN_WINDOW_VSEG");
    return e == nullptr || atoi(e) != 0;
}

// fold mode of the first (few-INPUT-channel) conv's data gradient: the same trick on the transposed problem (ACLGAN_FOLD_DGRAD=0:
// one 16-column MMA group per tap)
static bool fold_dgrad_enabled(const aclgan_conv_desc* cd) {
    const char* e = getenv("ACLGAN_FOLD_DGRAD");
    return (e == nullptr || atoi(e) != 0) && cd->window == ACLGAN_WINDOW_IN && cd->stride == 1 && cd->k <= 8 && cd->k > 1 &&
           cd->cin <= 8;
}

static bool wgrad_seg_enabled() {
    const char* e = getenv("ACLGAN_WGRAD_SEG");
    return e == nullptr || atoi(e) != 0;
}

static int k_channels(const aclgan_conv_desc* cd, int transposed) {
    return round_up(transposed ? cd->cout : cd->cin, 64);
}
static int window_cs(const aclgan_conv_desc* cd) { return cd->stride == 2 ? 16 : 8; }
// forward packing uses the pixel-window layout for WINDOW_IN convs, the transposed packing for WINDOW_OUT convs
static bool uses_window(const aclgan_conv_desc* cd, int transposed) {
    return transposed ? cd->window == ACLGAN_WINDOW_OUT : cd->window == ACLGAN_WINDOW_IN;
}

}  // namespace aclgan

using namespace aclgan;

extern "C" int aclgan_packed_weight_shape(const aclgan_conv_desc* cd, int transposed, int64_t* rows,
                                          int64_t* k_total) {
    int bn, nt;
    choose_block_n(transposed ? cd->cin : cd->cout, &bn, &nt);
    *rows = (int64_t)bn * nt;
    if (!transposed && fold_enabled(cd)) {          // rows n = kw * 8 + co, k = kh * Cs + ci
        *rows = 64;
        *k_total = (int64_t)cd->k * k_channels(cd, 0);
        return ACLGAN_OK;
    }
    if (transposed && fold_dgrad_enabled(cd)) {     // rows n = (k-1-kw) * 8 + ci, k = (k-1-kh) * Cs + co: the flipped filter, columns folded into N
        *rows = 64;
        *k_total = (int64_t)cd->k * k_channels(cd, 1);
        return ACLGAN_OK;
    }
    if (uses_window(cd, transposed)) *k_total = (int64_t)cd->k * 64;
    else *k_total = (int64_t)cd->k * cd->k * k_channels(cd, transposed);
    return ACLGAN_OK;
}

extern "C" int64_t aclgan_packed_weight_index(const aclgan_conv_desc* cd, int transposed, int co, int ci, int kh,
                                              int kw) {
    int64_t rows, kt;
    aclgan_packed_weight_shape(cd, transposed, &rows, &kt);
    if (!transposed && fold_enabled(cd)) return (int64_t)(kw * 8 + co) * kt + (int64_t)kh * k_channels(cd, 0) + ci;
    if (transposed && fold_dgrad_enabled(cd))
        return (int64_t)((cd->k - 1 - kw) * 8 + ci) * kt + (int64_t)(cd->k - 1 - kh) * k_channels(cd, 1) + co;
    if (!uses_window(cd, transposed)) {
        if (!transposed) return (int64_t)co * kt + (int64_t)(kh * cd->k + kw) * k_channels(cd, 0) + ci;
        return (int64_t)ci * kt + (int64_t)(kh * cd->k + kw) * k_channels(cd, 1) + co;
    }
    const int cs = window_cs(cd);
    if (!transposed) {
        // forward window: K chunk of filter row kh = pixels x .. x+win-1 of the stored plane, C=cs channels each
        return (int64_t)co * kt + (int64_t)kh * 64 + kw * cs + ci;
    }
    // data-gradient window (dY stored with cs=8 channels, zero border k-1): window pixel j <-> kw = k-1-j
    const int j = cd->k - 1 - kw;
    return (int64_t)ci * kt + (int64_t)kh * 64 + j * 8 + co;
}

extern "C" int aclgan_plan_conv_fwd(const aclgan_conv_desc* cd, const aclgan_act* x, const uint64_t w[2],
                                    const aclgan_out_spec* out, aclgan_igemm_plan* p) {
    memset(p, 0, sizeof(*p));
    if (cd->stride != 1 && cd->stride != 2) return ACLGAN_ERR_UNSUPPORTED;
    if (x->pad != cd->pad) return ACLGAN_ERR_SHAPE;
    const int k = cd->k, s = cd->stride;
    const int hp = x->h + 2 * x->pad, wp = x->w + 2 * x->pad;
    if (hp < k || wp < k) return ACLGAN_ERR_SHAPE;
    const int ho = (hp - k) / s + 1, wo = (wp - k) / s + 1;
    const int cs = x->c;
    const int64_t px = (int64_t)cs * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
    p->planes = x->planes;
    p->nseg = x->planes == 2 ? 3 : 1;
    choose_block_n(cd->cout, &p->block_n, &p->n_tiles);
    rect_tiles(p, wo, ho, x->n);
    int64_t rows, kt;
    aclgan_packed_weight_shape(cd, 0, &rows, &kt);
    for (int pl = 0; pl < x->planes; ++pl) weight_map2(&p->b[pl], w[pl], kt, rows, p->block_n);

    if (fold_enabled(cd)) {
        // filter columns folded into N: one tap per filter row over the flattened padded grid, overlapping 128-row tiles
        if (cs % 64 != 0 || cs != k_channels(cd, 0)) return ACLGAN_ERR_SHAPE;
        const int64_t total = (int64_t)x->n * hp * wp;
        p->block_n = 64; p->n_tiles = 1;
        p->fold = k;
        p->tile_step = 120;
        p->box_x = 128; p->box_y = 1; p->box_z = 1;
        p->tiles_x = (int)((total + p->tile_step - 1) / p->tile_step); p->tiles_y = 1; p->tiles_z = 1;
        p->flat = 1; p->flat_w = wp; p->flat_img = hp * wp;
        p->cchunks = cs / 64;
        p->num_taps = k;
        p->n_avariants = 1;
        for (int pl = 0; pl < x->planes; ++pl) {
            plane_map4(&p->a[pl][0], x->data[pl], cs, total, 1, 1, px, total * px, total * px, 128, 1, 1);
            weight_map2(&p->b[pl], w[pl], kt, rows, 64);
        }
        for (int kh = 0; kh < k; ++kh) {
            p->tap_dx[kh] = kh * wp;
            p->tap_bk[kh] = kh * cs;
        }
    } else if (cd->window != ACLGAN_WINDOW_IN && s == 1 && k > 1 && k <= 8 && seg_enabled()) {
        // stride 1: segment mode.  Tiles are 128 consecutive pixels of one output row when the rows are long enough,
        // otherwise 128 consecutive positions of the flattened padded input grid (outputs at the k - 1 right-most
        // columns / bottom rows of that grid are computed and dropped by the epilogue's extent check).
        if (cs % 64 != 0 || cs != k_channels(cd, 0)) return ACLGAN_ERR_SHAPE;
        p->cchunks = cs / 64;
        p->num_taps = k * k;
        p->n_avariants = 1;
        p->seg_mode = 1;
        p->seg_rows = round_up(128 + k - 1, 8);
        p->num_segs = k;
        p->seg_taps = k;
        const bool row_tiles = (wo % 128 == 0);
        const int64_t total = (int64_t)x->n * hp * wp;
        if (row_tiles) {
            p->box_x = 128; p->box_y = 1; p->box_z = 1;
            p->tiles_x = wo / 128; p->tiles_y = ho; p->tiles_z = x->n;
        } else {
            p->box_x = 128; p->box_y = 1; p->box_z = 1;
            p->tiles_x = (int)((total + 127) / 128); p->tiles_y = 1; p->tiles_z = 1;
            p->flat = 1; p->flat_w = wp; p->flat_img = hp * wp;
        }
        for (int pl = 0; pl < x->planes; ++pl) {
            if (row_tiles) {
                plane_map4(&p->a[pl][0], x->data[pl], cs, wp, hp, x->n, px, row, img, 128, 1, 1);
                plane_map4(&p->a_seg[pl], x->data[pl], cs, wp, hp, x->n, px, row, img, p->seg_rows, 1, 1);
            } else {
                plane_map4(&p->a[pl][0], x->data[pl], cs, total, 1, 1, px, total * px, total * px, 128, 1, 1);
                plane_map4(&p->a_seg[pl], x->data[pl], cs, total, 1, 1, px, total * px, total * px, p->seg_rows, 1, 1);
            }
        }
        for (int kh = 0; kh < k; ++kh) {
            p->seg_dx[kh] = row_tiles ? 0 : kh * wp;
            p->seg_dy[kh] = row_tiles ? kh : 0;
            for (int kw = 0; kw < k; ++kw) {
                const int t = kh * k + kw;
                p->tap_dx[t] = row_tiles ? kw : kh * wp + kw;
                p->tap_dy[t] = row_tiles ? kh : 0;
                p->tap_var[t] = 0;
                p->tap_bk[t] = t * cs;
                p->tap_row[t] = kw;
            }
        }
    } else if (cd->window != ACLGAN_WINDOW_IN) {
        if (cs % 64 != 0 || cs != k_channels(cd, 0)) return ACLGAN_ERR_SHAPE;
        if (k * k > ACLGAN_MAX_TAPS) return ACLGAN_ERR_UNSUPPORTED;
        p->cchunks = cs / 64;
        p->num_taps = k * k;
        p->n_avariants = (s == 1) ? 1 : 4;
        for (int pl = 0; pl < x->planes; ++pl) {
            if (s == 1) {
                plane_map4(&p->a[pl][0], x->data[pl], cs, wp, hp, x->n, px, row, img, p->box_x, p->box_y, p->box_z);
            } else {
                for (int ph = 0; ph < 2; ++ph)
                    for (int pw = 0; pw < 2; ++pw)
                        plane_map4(&p->a[pl][ph * 2 + pw], x->data[pl] + (uint64_t)(ph * row + pw * px), cs,
                                   (wp - pw + 1) / 2, (hp - ph + 1) / 2, x->n, 2 * px, 2 * row, img, p->box_x,
                                   p->box_y, p->box_z);
            }
        }
        for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
                const int t = kh * k + kw;
                p->tap_dx[t] = (s == 1) ? kw : kw / 2;
                p->tap_dy[t] = (s == 1) ? kh : kh / 2;
                p->tap_var[t] = (s == 1) ? 0 : (kh & 1) * 2 + (kw & 1);
                p->tap_bk[t] = t * cs;
            }
    } else {
        const int wcs = window_cs(cd);
        if (cs != wcs || cd->cin > cs || k > 64 / cs) return ACLGAN_ERR_SHAPE;
        p->cchunks = 1;
        p->num_taps = k;
        p->n_avariants = (s == 1) ? 1 : 2;
        // stride 1: VERTICAL segment mode.  Tiles are 16 x 8 output pixels; the k taps (= filter rows) of a tile read the
        // input rows y .. y + 8 + k - 2, so ONE box of 16 x (8 + k - 1) window rows is staged per tile and tap kh is the
        // 128-row window starting 16 * kh rows into it (a multiple of the 1 KB swizzle atom) - k times fewer activation bytes
        // cross L2 -> shared memory than with one 128-pixel box per tap, which is what bounded these layers
        const bool vseg = (s == 1) && k > 1 && 16 * (8 + k - 1) <= 256 && wo >= 16 && ho >= 8 && window_vseg_enabled();
        if (vseg) {
            p->box_x = 16; p->box_y = 8; p->box_z = 1;
            p->tiles_x = ceil_div(wo, 16); p->tiles_y = ceil_div(ho, 8); p->tiles_z = x->n;
            p->seg_mode = 1;
            p->seg_rows = 16 * (8 + k - 1);
            p->num_segs = 1;
            p->seg_taps = k;
            p->seg_dx[0] = 0; p->seg_dy[0] = 0;
            for (int kh = 0; kh < k; ++kh) p->tap_row[kh] = 16 * kh;
            for (int pl = 0; pl < x->planes; ++pl)
                plane_map4(&p->a_seg[pl], x->data[pl], 64, wp, hp, x->n, px, row, img, 16, 8 + k - 1, 1);
        }
        for (int pl = 0; pl < x->planes; ++pl) {
            if (s == 1) {
                // window starting at stored pixel x covers pixels x .. x+7 (the caller provides >= 64 elements of
                // zeroed slack behind the plane; inside the plane the overrun hits finite data times zero weights)
                plane_map4(&p->a[pl][0], x->data[pl], 64, wp, hp, x->n, px, row, img, p->box_x, p->box_y, p->box_z);
            } else {
                for (int ph = 0; ph < 2; ++ph)
                    plane_map4(&p->a[pl][ph], x->data[pl] + (uint64_t)(ph * row), 64, (wp + 1) / 2, (hp - ph + 1) / 2,
                               x->n, 2 * px, 2 * row, img, p->box_x, p->box_y, p->box_z);
            }
        }
        for (int kh = 0; kh < k; ++kh) {
            p->tap_dx[kh] = 0;
            p->tap_dy[kh] = (s == 1) ? kh : kh / 2;
            p->tap_var[kh] = (s == 1) ? 0 : (kh & 1);
            p->tap_bk[kh] = kh * 64;
        }
    }
    p->out = *out;
    if (p->out.N <= 0) { p->out.N = x->n; p->out.H = ho; p->out.W = wo; }
    p->n_groups = 1;
    p->group_taps = p->num_taps;
    return ACLGAN_OK;
}

extern "C" int aclgan_plan_conv_dgrad(const aclgan_conv_desc* cd, const aclgan_act* dy, const uint64_t wt[2],
                                      int phase, const aclgan_out_spec* out, aclgan_igemm_plan* p) {
    memset(p, 0, sizeof(*p));
    const int k = cd->k, s = cd->stride;
    if (s != 1 && s != 2) return ACLGAN_ERR_UNSUPPORTED;
    if (s == 2 && (k % 2 != 0)) return ACLGAN_ERR_UNSUPPORTED;
    const int pz = (s == 1) ? k - 1 : k / 2 - 1;          // zero border the caller stored around dY
    if (dy->pad != pz) return ACLGAN_ERR_SHAPE;
    const int ho = dy->h, wo = dy->w;
    const int hz = ho + 2 * pz, wz = wo + 2 * pz;
    const int cs = dy->c;
    const int64_t total = (int64_t)dy->n * hz * wz;
    p->planes = dy->planes;
    p->nseg = dy->planes == 2 ? 3 : 1;
    choose_block_n(cd->cin, &p->block_n, &p->n_tiles);
    p->box_x = 128; p->box_y = 1; p->box_z = 1;
    p->tiles_x = (int)((total + 127) / 128); p->tiles_y = 1; p->tiles_z = 1;
    p->flat = 1; p->flat_w = wz; p->flat_img = hz * wz;
    p->n_avariants = 1;
    int64_t rows, kt;
    aclgan_packed_weight_shape(cd, 1, &rows, &kt);
    for (int pl = 0; pl < dy->planes; ++pl) weight_map2(&p->b[pl], wt[pl], kt, rows, p->block_n);

    if (fold_dgrad_enabled(cd)) {
        // few input channels (first conv): dX_pad[q][ci] = sum_{kh', kw'} dYz[q + kh' Wz + kw'][co] * W[co][ci][k-1-kh'][k-1-kw'] is the
        // FORWARD fold problem (see aclgan_plan_conv_fwd) on the zero-bordered dY plane with the flipped filter: the k filter
        // columns are folded into N (weight row n = kw' * 8 + ci), one tap per filter row, diagonal sum in the epilogue -
        // k times fewer MMAs than one 16-column MMA group per tap (309 -> ~90 us on the 256 x 256 batch-8 plane)
        if (cs % 64 != 0 || cs != k_channels(cd, 1)) return ACLGAN_ERR_SHAPE;
        p->block_n = 64; p->n_tiles = 1;
        p->fold = k;
        p->tile_step = 120;
        p->tiles_x = (int)((total + p->tile_step - 1) / p->tile_step);
        p->cchunks = cs / 64;
        p->num_taps = k;
        for (int pl = 0; pl < dy->planes; ++pl) {
            plane_map4(&p->a[pl][0], dy->data[pl], cs, total, 1, 1, (int64_t)cs * 2, total * cs * 2, total * cs * 2, 128, 1, 1);
            weight_map2(&p->b[pl], wt[pl], kt, rows, 64);
        }
        for (int t = 0; t < k; ++t) {
            p->tap_dx[t] = t * wz;
            p->tap_bk[t] = t * cs;
        }
        p->out = *out;
        if (p->out.C > 8) p->out.C = 8;        // (the folded weight rows hold 8 channel slots per filter column)
        p->n_groups = 1;
        p->group_taps = p->num_taps;
        return ACLGAN_OK;
    }
    if (cd->window != ACLGAN_WINDOW_OUT) {
        if (cs % 64 != 0 || cs != k_channels(cd, 1)) return ACLGAN_ERR_SHAPE;
        p->cchunks = cs / 64;
        for (int pl = 0; pl < dy->planes; ++pl)
            plane_map4(&p->a[pl][0], dy->data[pl], cs, total, 1, 1, (int64_t)cs * 2, total * cs * 2, total * cs * 2,
                       128, 1, 1);
        int t = 0;
        if (s == 1) {
            for (int kh = 0; kh < k; ++kh)
                for (int kw = 0; kw < k; ++kw, ++t) {
                    p->tap_dx[t] = (k - 1 - kh) * wz + (k - 1 - kw);
                    p->tap_bk[t] = (kh * k + kw) * cs;
                    p->tap_row[t] = k - 1 - kw;
                }
            if (k > 1 && k <= 8 && seg_enabled()) {
                p->seg_mode = 1;
                p->seg_rows = round_up(128 + k - 1, 8);
                p->num_segs = k;
                p->seg_taps = k;
                for (int kh = 0; kh < k; ++kh) p->seg_dx[kh] = (k - 1 - kh) * wz;
                for (int pl = 0; pl < dy->planes; ++pl)
                    plane_map4(&p->a_seg[pl], dy->data[pl], cs, total, 1, 1, (int64_t)cs * 2, total * cs * 2, total * cs * 2,
                               p->seg_rows, 1, 1);
            }
        } else {
            for (int ph = (phase < 0 ? 0 : phase); ph <= (phase < 0 ? 3 : phase); ++ph) {
                const int pa = ph >> 1, pb = ph & 1;
                for (int kh = pa; kh < k; kh += 2)
                    for (int kw = pb; kw < k; kw += 2, ++t) {
                        const int a = (kh - pa) / 2, b = (kw - pb) / 2;
                        p->tap_dx[t] = (pz - a) * wz + (pz - b);
                        p->tap_bk[t] = (kh * k + kw) * cs;
                    }
            }
        }
        p->num_taps = t;
    } else {
        // dY stored with 8 channels per pixel; one tap per filter row, K chunk = 8 pixels x 8 channels
        if (s != 1 || cs != 8 || cd->cout > 8 || k > 8) return ACLGAN_ERR_SHAPE;
        p->cchunks = 1;
        p->num_taps = k;
        for (int pl = 0; pl < dy->planes; ++pl)
            plane_map4(&p->a[pl][0], dy->data[pl], 64, total, 1, 1, 16, total * 16, total * 16, 128, 1, 1);
        for (int kh = 0; kh < k; ++kh) {
            p->tap_dx[kh] = (k - 1 - kh) * wz;       // window pixel j = (k-1-kw) is part of the K chunk
            p->tap_bk[kh] = kh * 64;
        }
        if (k > 1 && 16 * (8 + k - 1) <= 256 && wz >= 16 && hz >= 8 && window_vseg_enabled()) {
            // vertical segment mode (see aclgan_plan_conv_fwd): 16 x 8 tiles of the zero-bordered grid, one staged box of
            // 16 x (8 + k - 1) window rows per tile; tap kh reads grid rows y + (k - 1 - kh), i.e. starts 16 (k - 1 - kh) rows
            // into the box (a DEscending arithmetic progression of start rows)
            const int gw = (out->W > 0 && out->W < wz) ? out->W : wz, gh = (out->H > 0 && out->H < hz) ? out->H : hz;
            p->flat = 0; p->flat_w = 1; p->flat_img = 1;
            p->box_x = 16; p->box_y = 8; p->box_z = 1;
            p->tiles_x = ceil_div(gw, 16); p->tiles_y = ceil_div(gh, 8); p->tiles_z = dy->n;
            p->seg_mode = 1;
            p->seg_rows = 16 * (8 + k - 1);
            p->num_segs = 1;
            p->seg_taps = k;
            p->seg_dx[0] = 0; p->seg_dy[0] = 0;
            for (int kh = 0; kh < k; ++kh) {
                p->tap_row[kh] = 16 * (k - 1 - kh);
                p->tap_dx[kh] = 0;
                p->tap_dy[kh] = k - 1 - kh;
            }
            for (int pl = 0; pl < dy->planes; ++pl) {
                plane_map4(&p->a[pl][0], dy->data[pl], 64, wz, hz, dy->n, 16, (int64_t)wz * 16, (int64_t)hz * wz * 16, 16, 8, 1);
                plane_map4(&p->a_seg[pl], dy->data[pl], 64, wz, hz, dy->n, 16, (int64_t)wz * 16, (int64_t)hz * wz * 16, 16,
                           8 + k - 1, 1);
            }
        }
    }
    p->out = *out;
    p->n_groups = 1;
    p->group_taps = p->num_taps;
    if (s == 2 && phase < 0) {
        // all four output-parity phases in one launch: `out` is the strided view of phase (0,0)
        if (p->num_taps > ACLGAN_MAX_TAPS || (p->num_taps % 4) != 0) return ACLGAN_ERR_UNSUPPORTED;
        p->n_groups = 4;
        p->group_taps = p->num_taps / 4;
        for (int ph = 0; ph < 4; ++ph) p->group_off[ph] = (ph >> 1) * (out->sy / 2) + (ph & 1) * (out->sx / 2);
    }
    return ACLGAN_OK;
}

// ------------------------------------------------------------------------------------------------ wgrad
extern "C" int aclgan_wgrad_layout(const aclgan_conv_desc* cd) {
    if (cd->window == ACLGAN_WINDOW_IN) return 0;
    if (cd->window == ACLGAN_WINDOW_OUT) return 1;
    return (cd->cout >= 128 || cd->cout >= cd->cin) ? 0 : 1;
}

namespace aclgan {
// interior view of a (possibly bordered) plane as a (c, x, y, z) map with a 64-pixel box
static void interior_map(aclgan_tmap_spec* m, const aclgan_act* a, int pl, int c_extent, int bx, int by, int bz) {
    const int64_t px = (int64_t)a->c * 2, row = (int64_t)(a->w + 2 * a->pad) * px,
                  img = (int64_t)(a->h + 2 * a->pad) * row;
    plane_map4(m, a->data[pl] + (uint64_t)(a->pad * row + a->pad * px), c_extent, a->w, a->h, a->n, px, row, img, bx,
               by, bz);
}
// conv-input maps + per-tap offsets, shared logic with the forward plan (regular layouts only)
static void input_maps(aclgan_tmap_spec (*maps)[ACLGAN_MAX_AVARIANTS], int* nvar, const aclgan_act* x, int s, int bx,
                       int by, int bz) {
    const int hp = x->h + 2 * x->pad, wp = x->w + 2 * x->pad;
    const int64_t px = (int64_t)x->c * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
    *nvar = (s == 1) ? 1 : 4;
    for (int pl = 0; pl < x->planes; ++pl) {
        if (s == 1) plane_map4(&maps[pl][0], x->data[pl], x->c, wp, hp, x->n, px, row, img, bx, by, bz);
        else
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw)
                    plane_map4(&maps[pl][ph * 2 + pw], x->data[pl] + (uint64_t)(ph * row + pw * px), x->c,
                               (wp - pw + 1) / 2, (hp - ph + 1) / 2, x->n, 2 * px, 2 * row, img, bx, by, bz);
    }
}
}  // namespace aclgan

extern "C" int aclgan_plan_conv_wgrad(const aclgan_conv_desc* cd, const aclgan_act* dy, const aclgan_act* x,
                                      uint64_t dw, aclgan_wgrad_plan* p) {
    memset(p, 0, sizeof(*p));
    const int k = cd->k, s = cd->stride;
    if (s != 1 && s != 2) return ACLGAN_ERR_UNSUPPORTED;
    if (x->pad != cd->pad || x->planes != dy->planes || x->n != dy->n) return ACLGAN_ERR_SHAPE;
    const int hp = x->h + 2 * x->pad, wp = x->w + 2 * x->pad;
    const int ho = (hp - k) / s + 1, wo = (wp - k) / s + 1;
    if (dy->h != ho || dy->w != wo) return ACLGAN_ERR_SHAPE;
    const int layout = aclgan_wgrad_layout(cd);
    int64_t rows, kt;
    aclgan_packed_weight_shape(cd, layout, &rows, &kt);
    p->planes = x->planes;
    p->nseg = x->planes == 2 ? 3 : 1;
    p->dw = dw;
    p->dw_sm = kt;

    // reduction grid and its pixel boxes: 64 pixels per pipeline stage, or 128 when the operands are narrow (at most
    // three 64-channel chunks per stage) so that the per-stage barrier / issue overhead is amortised
    const int gw = (cd->window == ACLGAN_WINDOW_OUT) ? wp : wo, gh = ho;
    int chunks_guess;
    if (cd->window != ACLGAN_WINDOW_NONE) chunks_guess = 2;
    else {
        const int cm = layout == 0 ? dy->c : x->c, cn = layout == 0 ? x->c : dy->c;
        chunks_guess = (cm / 64 >= 2 ? 2 : 1) + (cn / 64 > 4 ? 4 : cn / 64);
    }
    const int pix = (chunks_guess <= 3 && (int64_t)gw * gh * x->n >= 4096) ? 128 : 64;
    int bx = pow2_floor(gw < pix ? gw : pix);
    int by = pix / bx;
    if (by > pow2_ceil(gh)) by = pow2_ceil(gh);
    const int bz = pix / (bx * by);
    p->box_x = bx; p->box_y = by; p->box_z = bz;
    p->blocks_x = ceil_div(gw, bx); p->blocks_y = ceil_div(gh, by); p->blocks_z = ceil_div(x->n, bz);

    if (cd->window == ACLGAN_WINDOW_NONE) {
        if (k * k > ACLGAN_MAX_TAPS || x->c % 64 || dy->c % 64) return ACLGAN_ERR_SHAPE;
        p->num_taps = k * k;
        aclgan_tmap_spec(*xm)[ACLGAN_MAX_AVARIANTS] = layout == 0 ? p->nop : p->mop;
        aclgan_tmap_spec(*ym)[ACLGAN_MAX_AVARIANTS] = layout == 0 ? p->mop : p->nop;
        int nvar = 1;
        input_maps(xm, &nvar, x, s, bx, by, bz);
        for (int pl = 0; pl < dy->planes; ++pl) interior_map(&ym[pl][0], dy, pl, dy->c, bx, by, bz);
        int* xdx = layout == 0 ? p->n_dx : p->m_dx;
        int* xdy = layout == 0 ? p->n_dy : p->m_dy;
        int* xvar = layout == 0 ? p->n_var : p->m_var;
        for (int kh = 0; kh < k; ++kh)
            for (int kw = 0; kw < k; ++kw) {
                const int t = kh * k + kw;
                xdx[t] = (s == 1) ? kw : kw / 2;
                xdy[t] = (s == 1) ? kh : kh / 2;
                xvar[t] = (s == 1) ? 0 : (kh & 1) * 2 + (kw & 1);
                p->tap_out[t] = t;
            }
        const int cm = layout == 0 ? dy->c : x->c, cn = layout == 0 ? x->c : dy->c;
        p->n_mvariants = layout == 0 ? 1 : nvar;
        p->n_nvariants = layout == 0 ? nvar : 1;
        p->m_chunks = cm / 64 >= 2 ? 2 : 1;
        p->m_tiles = ceil_div(cm, 128);
        p->n_chunks = cn / 64 > 4 ? 4 : cn / 64;
        p->n_tiles = ceil_div(cn / 64, p->n_chunks);
        p->M = layout == 0 ? cd->cout : cd->cin;
        p->Nn = cn;
        p->dw_st = cn;
    } else if (cd->window == ACLGAN_WINDOW_IN) {
        const int cs = x->c;
        if (cs != window_cs(cd) || dy->c % 64 || k > 64 / cs) return ACLGAN_ERR_SHAPE;
        p->num_taps = k;
        const int64_t px = (int64_t)cs * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
        for (int pl = 0; pl < x->planes; ++pl) {
            interior_map(&p->mop[pl][0], dy, pl, dy->c, bx, by, bz);
            if (s == 1) plane_map4(&p->nop[pl][0], x->data[pl], 64, wp, hp, x->n, px, row, img, bx, by, bz);
            else
                for (int ph = 0; ph < 2; ++ph)
                    plane_map4(&p->nop[pl][ph], x->data[pl] + (uint64_t)(ph * row), 64, (wp + 1) / 2,
                               (hp - ph + 1) / 2, x->n, 2 * px, 2 * row, img, bx, by, bz);
        }
        for (int kh = 0; kh < k; ++kh) {
            p->n_dy[kh] = (s == 1) ? kh : kh / 2;
            p->n_var[kh] = (s == 1) ? 0 : (kh & 1);
            p->tap_out[kh] = kh;
        }
        p->n_mvariants = 1;
        p->n_nvariants = (s == 1) ? 1 : 2;
        p->m_chunks = dy->c / 64 >= 2 ? 2 : 1;
        p->m_tiles = ceil_div(dy->c, 128);
        p->n_chunks = 1; p->n_tiles = 1;
        p->M = cd->cout; p->Nn = 64; p->dw_st = 64;
    } else {
        // WINDOW_OUT: M = conv input (taps = filter rows), N = 8-pixel window of the zero-bordered dY (cs = 8)
        if (s != 1 || dy->c != 8 || dy->pad != k - 1 || x->c % 64 || k > 8) return ACLGAN_ERR_SHAPE;
        p->num_taps = k;
        int nvar = 1;
        input_maps(p->mop, &nvar, x, 1, bx, by, bz);
        const int hz = ho + 2 * dy->pad, wz = wo + 2 * dy->pad;
        for (int pl = 0; pl < dy->planes; ++pl)
            plane_map4(&p->nop[pl][0], dy->data[pl], 64, wz, hz, dy->n, 16, (int64_t)wz * 16, (int64_t)hz * wz * 16, bx,
                       by, bz);
        for (int kh = 0; kh < k; ++kh) {
            p->m_dy[kh] = kh;
            p->n_dy[kh] = dy->pad;
            p->tap_out[kh] = kh;
        }
        p->n_mvariants = 1; p->n_nvariants = 1;
        p->m_chunks = x->c / 64 >= 2 ? 2 : 1;
        p->m_tiles = ceil_div(x->c, 128);
        p->n_chunks = 1; p->n_tiles = 1;
        p->M = cd->cin; p->Nn = 64; p->dw_st = 64;
    }
    if (cd->window == ACLGAN_WINDOW_NONE && s == 1 && k > 1 && k <= 7 && wo % 64 == 0 && wgrad_seg_enabled()) {
        // segment mode: re-plan with 64-pixel row blocks, one CTA-tile per (filter row, m tile, n tile)
        const int cm = layout == 0 ? dy->c : x->c, cn = layout == 0 ? x->c : dy->c;
        const int n_chunks = cn / 64 >= 2 ? 2 : 1;     // 128-column accumulators when the operand is wide enough
        const int per_cta = 512 / (64 * n_chunks);     // taps whose accumulators fit the 512 TMEM columns
        const int groups = ceil_div(k, per_cta);       // a filter row is split into `groups` entries of ~equal size
        if (k * groups <= ACLGAN_MAX_TAPS) {
            p->seg_mode = 1;
            p->seg_taps = ceil_div(k, groups);
            p->seg_rows = round_up(64 + k - 1, 8);
            p->seg_on_m = layout == 0 ? 0 : 1;
            p->box_x = 64; p->box_y = 1; p->box_z = 1;
            p->blocks_x = wo / 64; p->blocks_y = ho; p->blocks_z = x->n;
            p->num_taps = k * groups;
            p->m_chunks = cm / 64 >= 2 ? 2 : 1;
            p->m_tiles = ceil_div(cm, 128);
            p->n_chunks = n_chunks;
            p->n_tiles = ceil_div(cn / 64, n_chunks);
            const int64_t px = (int64_t)x->c * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
            for (int pl = 0; pl < x->planes; ++pl) {
                aclgan_tmap_spec* xm = layout == 0 ? &p->nop[pl][0] : &p->mop[pl][0];
                aclgan_tmap_spec* ym = layout == 0 ? &p->mop[pl][0] : &p->nop[pl][0];
                plane_map4(xm, x->data[pl], x->c, wp, hp, x->n, px, row, img, 64, 1, 1);
                plane_map4(&p->seg_map[pl], x->data[pl], x->c, wp, hp, x->n, px, row, img, p->seg_rows, 1, 1);
                interior_map(ym, dy, pl, dy->c, 64, 1, 1);
            }
            p->n_mvariants = 1; p->n_nvariants = 1;
            int t = 0;
            for (int kh = 0; kh < k; ++kh) {
                int kw0 = 0;
                for (int g = 0; g < groups; ++g, ++t) {
                    const int cnt = (k - kw0 + (groups - g) - 1) / (groups - g);      // remaining taps spread evenly
                    p->m_dx[t] = p->m_dy[t] = p->n_dx[t] = p->n_dy[t] = 0;
                    p->m_var[t] = p->n_var[t] = 0;
                    if (layout == 0) p->n_dy[t] = kh; else p->m_dy[t] = kh;
                    p->seg_kw0[t] = kw0;
                    p->seg_cnt[t] = cnt;
                    p->tap_out[t] = kh * k + kw0;
                    kw0 += cnt;
                }
            }
        }
    }
    if (cd->window == ACLGAN_WINDOW_IN && s == 1 && k > 1 && k * 64 <= 512 && 16 * (4 + k - 1) <= 256 && wo >= 16 && ho >= 4 &&
        window_vseg_enabled()) {
        // vertical segment mode of the stride-1 pixel-window layer: a tap is a filter ROW, so the shifted operand (the window
        // rows of the conv input) is staged once per 16 x 4 pixel block as 16 x (4 + k - 1) window rows and tap kh starts
        // 16 * kh rows into it; ONE CTA-tile reduces all k taps (k x 64 accumulator columns)
        p->seg_mode = 1;
        p->seg_taps = k;
        p->seg_step = 16;
        p->seg_rows = 16 * (4 + k - 1);
        p->seg_on_m = 0;
        p->box_x = 16; p->box_y = 4; p->box_z = 1;
        p->blocks_x = ceil_div(wo, 16); p->blocks_y = ceil_div(ho, 4); p->blocks_z = x->n;
        p->num_taps = 1;
        const int64_t px = (int64_t)x->c * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
        for (int pl = 0; pl < x->planes; ++pl) {
            interior_map(&p->mop[pl][0], dy, pl, dy->c, 16, 4, 1);
            plane_map4(&p->nop[pl][0], x->data[pl], 64, wp, hp, x->n, px, row, img, 16, 4, 1);
            plane_map4(&p->seg_map[pl], x->data[pl], 64, wp, hp, x->n, px, row, img, 16, 4 + k - 1, 1);
        }
        p->n_mvariants = 1; p->n_nvariants = 1;
        p->m_dx[0] = p->m_dy[0] = p->n_dx[0] = p->n_dy[0] = 0;
        p->m_var[0] = p->n_var[0] = 0;
        p->seg_kw0[0] = 0;
        p->seg_cnt[0] = k;
        p->tap_out[0] = 0;
    }
    if (cd->window == ACLGAN_WINDOW_OUT && s == 1 && k > 1 && k * 64 <= 512 && 16 * (4 + k - 1) <= 256 && wp >= 16 && ho >= 4 &&
        window_vseg_enabled()) {
        // the same for the few-output-channel final conv: here the SHIFTED operand is M (the conv input rows y + kh), the
        // 8-pixel window of the zero-bordered dY is the fixed N operand
        p->seg_mode = 1;
        p->seg_taps = k;
        p->seg_step = 16;
        p->seg_rows = 16 * (4 + k - 1);
        p->seg_on_m = 1;
        p->box_x = 16; p->box_y = 4; p->box_z = 1;
        p->blocks_x = ceil_div(wp, 16); p->blocks_y = ceil_div(ho, 4); p->blocks_z = x->n;
        p->num_taps = 1;
        int nvar = 1;
        input_maps(p->mop, &nvar, x, 1, 16, 4, 1);
        const int64_t px = (int64_t)x->c * 2, row = (int64_t)wp * px, img = (int64_t)hp * row;
        const int hz = ho + 2 * dy->pad, wz = wo + 2 * dy->pad;
        for (int pl = 0; pl < x->planes; ++pl) {
            plane_map4(&p->seg_map[pl], x->data[pl], x->c, wp, hp, x->n, px, row, img, 16, 4 + k - 1, 1);
            plane_map4(&p->nop[pl][0], dy->data[pl], 64, wz, hz, dy->n, 16, (int64_t)wz * 16, (int64_t)hz * wz * 16, 16, 4, 1);
        }
        p->n_mvariants = 1; p->n_nvariants = 1;
        p->m_dx[0] = p->m_dy[0] = p->n_dx[0] = 0;
        p->n_dy[0] = dy->pad;
        p->m_var[0] = p->n_var[0] = 0;
        p->seg_kw0[0] = 0;
        p->seg_cnt[0] = k;
        p->tap_out[0] = 0;
    }
    // split-K so that the grid is ONE wave of CTAs (<= number of SMs): a grid of 162 CTAs on 148 SMs would run two
    // waves and take twice as long as 144
    const int blocks_total = p->blocks_x * p->blocks_y * p->blocks_z;
    const int tiles = p->num_taps * p->m_tiles * p->n_tiles;
    int ks = num_sms() / tiles;
    if (ks > blocks_total / 4) ks = blocks_total / 4;
    if (ks < 1) ks = 1;
    p->ksplit = ks;
    return ACLGAN_OK;
}
