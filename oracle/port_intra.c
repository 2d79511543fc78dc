/* oracle/port_intra.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of OpenVVC's luma intra prediction: vvc_intra_pred (libovvc/rcn_intra.c:484-560) with the reference
 * sample fill of libovvc/rcn_fill_ref.c (fill_ref_left_0 :71-145, fill_ref_above_0 :320-390, filter_ref_samples
 * :47-68), planar / DC with PDPC (libovvc/rcn_intra_dc_planar.c:151-236), the angular dispatch (rcn_intra.c:55-481)
 * and kernels (libovvc/rcn_intra_angular.c:102-930; angle tables libovvc/data_rcn_angular.c:37-60).
 * The horizontal class is the vertical class on the transposed block with the roles of the two reference arrays
 * swapped, which is how it is written here.
 */
#include "oracle.h"
#include "mip_tables.h"
#include <stdlib.h>
#include <string.h>

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static int floor_log2u(unsigned v) { int n = -1; while (v) { v >>= 1; n++; } return n; }

static const int angle_table[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29,
                                     32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
static const int inv_angle_table[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170, 1024, 910, 819, 712, 630, 565,
                                         512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
static const uint8_t intra_filter_thr[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };
/* cubic interpolation taps = the chroma MC filter (rcn_intra_angular.c:43-77) */
static const int8_t cubic[32][4] = {
    { 0, 64, 0, 0 }, { -1, 63, 2, 0 }, { -2, 62, 4, 0 }, { -2, 60, 7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
    { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 }, { -6, 49, 24, -3 },
    { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 }, { -4, 36, 36, -4 }, { -4, 33, 39, -4 },
    { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 }, { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 },
    { -2, 16, 54, -4 }, { -2, 15, 55, -4 }, { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 },
    { 0, 4, 62, -2 }, { 0, 2, 63, -1 } };

static inline int pdpc_w(int scale, int pos) { int s = (pos << 1) >> scale; return s > 5 ? 0 : 32 >> s; }   /* vvc_pdpc_w[scale][pos] */

enum { REF_OFF = 128, REF_LEN = 128 + 2 * 128 + 16 };

/* fill_ref_left_0 / fill_ref_above_0 with offset 0: ref[0] = corner, ref[1 + i] = i-th neighbour, 2 * len + 1 samples + 4
 * of wide-angle padding.  `step` walks along the neighbour line, `p` points at the corner sample. */
/* us = log2 of the availability unit in samples of this plane: 2 for luma, 1 for 4:2:0 chroma (fill_ref_*_0_chroma,
 * rcn_fill_ref.c:147-230, 392-480).  left_chroma: fill_ref_left_0_chroma pads with the mid value when only the corner
 * unit is available (its padding value is only updated inside the copy loop, :193-200). */
static void fill_ref_off(uint16_t *ref, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                         uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int us, int left_chroma, int off);
static void fill_ref(uint16_t *ref, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                     uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int us, int left_chroma)
{
    fill_ref_off(ref, p, step, other_step, log2_len, log2_other, map, other_map, pos, other_pos, bitdepth, us, left_chroma, 0);
}
/* off != 0 (ISP, rcn_fill_ref.c:71-145 / 320-390 with offset_y / offset_x): the array is filled for the whole coding block
 * starting `off` samples before the partition; only the corner sample, the "nothing available" padding value and the
 * length of the wide-angle padding depend on it */
static void fill_ref_off(uint16_t *ref, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                         uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int us, int left_chroma, int off)
{
    int un = 1 << us, nb_units = ((1 << (log2_len + 1)) >> us) + 1, ref_len = (1 << (log2_len + 1)) + 1;
    uint64_t mask = (1ull << (nb_units + 1)) - 1, avl = (map >> (pos >> us)) & mask, navl = avl ^ mask;
    if (!navl) {
        for (int i = 0; i < ref_len; i++) ref[i] = p[i * step];
    } else if (avl) {
        int nb_avl = 64 - __builtin_clzll(avl), i;
        ref[0] = (avl & 1) ? p[off * step] : p[step];
        for (i = 1; i < nb_avl; i++)
            for (int k = 0; k < un; k++) ref[1 + un * (i - 1) + k] = p[(1 + un * (i - 1) + k) * step];
        uint16_t pad = ref[un * (nb_avl - 1)];
        if (left_chroma && nb_avl == 1) pad = (uint16_t)(1 << (bitdepth - 1));
        for (; i < nb_units; i++)
            for (int k = 0; k < un; k++) ref[1 + un * (i - 1) + k] = pad;
    } else {
        int nb_units_o = ((1 << (log2_other + 1)) >> us) + 1;
        uint64_t avl_o = (other_map >> (other_pos >> us)) & ((1ull << (nb_units_o + 1)) - 1);
        uint16_t pad = avl_o ? p[other_step + off * step] : (uint16_t)(1 << (bitdepth - 1));     /* first sample of the other line */
        for (int i = 0; i < ref_len; i++) ref[i] = pad;
    }
    for (int i = 0; i < 4 + off; i++) ref[ref_len + i] = ref[ref_len - 1];
}

/* fill_ref_left_0_mref / fill_ref_above_0_mref (rcn_fill_ref.c:228-318, 493-590) for reference line m = mrl_idx >= 1:
 * raw[0] is the sample at distance m + 1 from the block both ways (p), the corner group is m + 1 samples, then 4-sample
 * units; raw has 2 * len + 1 + (m + 1) samples + ratio * (m + 1) of wide-angle padding.  is_left selects the two places
 * where the reference's left and above variants differ.  Cases the decoder cannot reach (MRL needs y0 > 0 inside the CTU,
 * vcl_coding_unit.c:1377): left units available without the corner unit -- restated as written (:264-270, the unit copy
 * then starts m rows early); no above unit at all -- the reference leaves m + 1 entries unwritten (:553-575), here they
 * get the padding value. */
static void fill_ref_mref(uint16_t *raw, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                          uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int m, int is_left)
{
    int nb_units = ((1 << (log2_len + 1)) >> 2) + 1, ref_len = (1 << (log2_len + 1)) + 1 + (m + 1);
    int ratio = (1 << log2_len) >> log2_other; if (ratio < 1) ratio = 1;
    uint64_t mask = (1ull << (nb_units + 1)) - 1, avl = (map >> (pos >> 2)) & mask, navl = avl ^ mask;
    if (!navl) {
        for (int i = 0; i < ref_len; i++) raw[i] = p[i * step];
    } else if (avl) {
        int nb_avl = 64 - __builtin_clzll(avl), i;
        const uint16_t *q = p;
        uint16_t *d = raw;
        uint16_t pad = (uint16_t)(1 << (bitdepth - 1));
        if (avl & 1) { for (i = 0; i < m + 1; i++) { *d++ = *q; q += step; } }
        else {
            q += is_left ? step : (m + 1) * step;
            for (i = 0; i < m + 1; i++) *d++ = *q;
        }
        if (!is_left) pad = q[-step];                              /* above: src_line[-1] after the copy loop (:541) */
        for (i = 1; i < nb_avl; i++) { for (int k = 0; k < 4; k++) d[k] = q[k * step]; pad = q[3 * step]; q += 4 * step; d += 4; }
        for (; i < nb_units; i++) { for (int k = 0; k < 4; k++) d[k] = pad; d += 4; }
    } else {
        int nb_units_o = ((1 << (log2_other + 1)) >> 2) + 1;
        uint64_t avl_o = (other_map >> (other_pos >> 2)) & ((1ull << (nb_units_o + 1)) - 1);
        uint16_t pad = (uint16_t)(1 << (bitdepth - 1));
        if (avl_o) pad = is_left ? p[(1 + m) * other_step] : p[m * step + (m + 1) * other_step];
        for (int i = 0; i < ref_len; i++) raw[i] = pad;
    }
    /* wide-angle padding as written (:312-316, :583-587): it starts ON the last sample, which is therefore replaced by its
     * neighbour in every case */
    for (int i = 0; i < ratio * (m + 1); i++) raw[ref_len - 1 + i] = raw[ref_len - 2 + i];
}

/* vvc_intra_pred_multi_ref (rcn_intra.c:645-786): no reference filtering, no PDPC, cubic interpolation only, every row
 * m lines further along the prediction direction (intra_angular_{v,h}_cubic_mref, rcn_intra_angular.c:932-1005) */
static void mrl_class(int *out, uint16_t *M, const uint16_t *S, int lw, int lh, int mode_idx, int maxv, int m)
{
    int W = 1 << lw, H = 1 << lh, a = abs(mode_idx);
    int angle = a == 16 ? 32 : angle_table[a];
    if (mode_idx < 0) {
        int inv = inv_angle_table[a];
        for (int k = -1, sum = 256; k >= -H; k--) { sum += inv; M[-m + k] = S[-m + imin(sum >> 9, H)]; }   /* :709-713 on the raw arrays */
        angle = -angle;
    }
    for (int y = 0; y < H; y++) {
        int delta = (y + 1 + m) * angle, di = delta >> 5, df = delta & 31;
        for (int x = 0; x < W; x++) {
            const uint16_t *r = M + x + di;
            if (a == 0) out[y * W + x] = M[x + 1];
            else if (!(angle & 31)) out[y * W + x] = r[1];
            else out[y * W + x] = clipi((r[0] * cubic[df][0] + r[1] * cubic[df][1] + r[2] * cubic[df][2] + r[3] * cubic[df][3] + 32) >> 6, 0, maxv);
        }
    }
}

/* filter_ref_samples, rcn_fill_ref.c:47-68 */
static void filter_ref(uint16_t *dst, const uint16_t *src, const uint16_t *src2, int length)
{
    dst[0] = (uint16_t)((src2[1] + (src[0] << 1) + src[1] + 2) >> 2);
    for (int i = 1; i < length; i++) dst[i] = (uint16_t)((src[i + 1] + (src[i] << 1) + src[i - 1] + 2) >> 2);
    dst[length] = src[length];
}

/* vertical-class angular prediction of a W x H block (W along the main reference): out[y * W + x].
 * M = main reference (M[0] corner, negative indices writable), S = side reference, tl = ref_above[0]. */
static void angular_class2(int *out, uint16_t *M, uint16_t *S, int tl, int lw, int lh, int mode_idx, int maxv, int chroma, int isp_lh);
static void angular_class(int *out, uint16_t *M, uint16_t *S, int tl, int lw, int lh, int mode_idx, int maxv, int chroma)
{
    angular_class2(out, M, S, tl, lw, lh, mode_idx, maxv, chroma, -1);
}
/* isp_lh >= 0: intra_angular_cubic_{v,h} of the ISP path (rcn_intra.c:288-452): cubic interpolation only, no reference
 * filtering, and the PDPC decisions test log2_pb_h of the BLOCK (isp_lh) -- also in the horizontal variant, where every other
 * quantity is the transposed one */
static void angular_class2(int *out, uint16_t *M, uint16_t *S, int tl, int lw, int lh, int mode_idx, int maxv, int chroma, int isp_lh)
{
    int W = 1 << lw, H = 1 << lh, a = abs(mode_idx), isp = isp_lh >= 0;
    int gauss = !isp && !chroma && (lw + lh) > 5 && a > intra_filter_thr[(lw + lh) >> 1];
    uint16_t fM[REF_LEN], fS[REF_LEN];
    int cpd = lw > 1 && lh > 1;                                    /* chroma: PDPC only when both sides are >= 4 (rcn_intra.c:954-1008) */
    if (a == 0) {                                                  /* pure vertical: intra_angular_ver(_pdpc) */
        int pd = isp ? isp_lh > 1 : (chroma ? cpd : (gauss || lh > 1)), scale = (lw + lh - 2) >> 2;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                out[y * W + x] = pd ? clipi((pdpc_w(scale, x) * (S[y + 1] - tl) + (M[x + 1] << 6) + 32) >> 6, 0, maxv) : M[x + 1];
        return;
    }
    if (mode_idx == 16) {                                          /* pure diagonal: intra_angular_vdia(_pdpc) */
        int pd = isp ? isp_lh > 1 : (chroma ? cpd : (gauss || lh > 1)), scale = imin(2, lh - 2);
        if (gauss) {
            filter_ref(fM + REF_OFF, M, S, 2 * W); filter_ref(fS + REF_OFF, S, M, 2 * H);
            M = fM + REF_OFF; S = fS + REF_OFF;
        }
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int v = M[x + y + 2];
                if (pd && x < imin(3 << scale, W)) v = clipi(v + (((32 >> ((2 * x) >> scale)) * (S[y + x + 2] - v) + 32) >> 6), 0, maxv);
                out[y * W + x] = v;
            }
        return;
    }
    int angle = angle_table[a], inv = inv_angle_table[a], frac = (angle & 31) != 0;
    int pdpc_scale = imin(2, lh - (floor_log2u(3 * inv - 2) - 8));
    int use_pdpc = 0;
    if (mode_idx < 0) {
        if (gauss && !frac) {                                      /* integer slope: [1 2 1]-filtered references */
            filter_ref(fM + REF_OFF, M, S, 2 * W); filter_ref(fS + REF_OFF, S, M, 2 * H);
            M = fM + REF_OFF; S = fS + REF_OFF;
        }
        for (int k = -1, sum = 256; k >= -H; k--) { sum += inv; M[k] = S[imin(sum >> 9, H)]; }     /* rcn_intra.c:118-121 */
        angle = -angle;
    } else if (pdpc_scale < 0 || (isp ? isp_lh < 2 : ((!gauss && lh < 2) || (chroma && !cpd)))) {
        if (gauss && !frac) { filter_ref(fM + REF_OFF, M, S, 2 * W); M = fM + REF_OFF; }
    } else {
        if (gauss && !frac) {
            filter_ref(fM + REF_OFF, M, S, 2 * W); filter_ref(fS + REF_OFF, S, M, 2 * H);
            M = fM + REF_OFF; S = fS + REF_OFF;
        }
        use_pdpc = 1;
    }
    for (int y = 0; y < H; y++) {
        int delta = (y + 1) * angle, di = delta >> 5, df = delta & 31;
        for (int x = 0; x < W; x++) {
            int v;
            if (!frac) v = M[x + di + 1];
            else {
                const uint16_t *r = M + x + di;
                if (chroma) v = clipi(r[1] + ((df * (r[2] - r[1]) + 16) >> 5), 0, maxv);    /* intra_angular_v_c, rcn_intra_angular.c:270-294 */
                else if (gauss) v = (r[0] * (16 - (df >> 1)) + r[1] * (32 - (df >> 1)) + r[2] * (16 + (df >> 1)) + r[3] * (df >> 1) + 32) >> 6;
                else       v = clipi((r[0] * cubic[df][0] + r[1] * cubic[df][1] + r[2] * cubic[df][2] + r[3] * cubic[df][3] + 32) >> 6, 0, maxv);
            }
            if (use_pdpc && x < imin(3 << pdpc_scale, W)) {
                int wL = 32 >> ((x << 1) >> pdpc_scale), left = S[y + ((256 + (x + 1) * inv) >> 9) + 1];
                v = clipi(v + ((wL * (left - v) + 32) >> 6), 0, maxv);
            }
            out[y * W + x] = v;
        }
    }
}

/* ---- cross-component linear model (rcn_intra_cclm.c:57-903) -------------------------------------------------------------
 * One luma sample per chroma position: 6-tap [1 2 1; 1 2 1] over two luma rows (compute_lm_subsample :626-657) or, with
 * sps_chroma_vertical_collocated_flag, the 5-tap cross (compute_lm_subsample_cl :297-324); y points at luma (2i, 2j). */
static int lm_luma(const uint16_t *y, ptrdiff_t st, int coll, int pad_left, int pad_abv)
{
    int l = pad_left ? 0 : -1;
    if (coll) return (y[pad_abv ? 0 : -st] + 4 * y[0] + y[l] + y[1] + y[st] + 4) >> 3;
    return (y[l] + 2 * y[0] + y[1] + y[st + l] + 2 * y[st] + y[st + 1] + 4) >> 3;
}

typedef struct { int a, b, shift; } LmParams;

/* compute_lm_params, :72-106 */
static LmParams lm_params_c(int min_l, int min_c, int max_c, int v, int log2_rng_l)
{
    LmParams p;
    int range_c = max_c - min_c;
    int l2c = range_c ? floor_log2u((unsigned)abs(range_c)) + 1 : 0;
    p.a = (range_c * v + ((1 << l2c) >> 1)) >> l2c;
    p.shift = 3 + log2_rng_l - l2c;
    if (p.shift < 1) { p.shift = 1; p.a = p.a ? (p.a < 0 ? -15 : 15) : 0; }
    p.b = min_c - ((p.a * min_l) >> p.shift);
    return p;
}

static void cclm_block(uint16_t *const planes[3], const int stride[3], int bitdepth, int x0, int y0, int xl, int yl,
                       int lw, int lh, int mode, uint64_t lft_map, uint64_t abv_map, int coll)
{
    static const uint8_t div_lut[16] = { 0, 7, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1, 1, 1, 0 };
    int w = 1 << lw, h = 1 << lh, maxv = (1 << bitdepth) - 1;
    ptrdiff_t sl = stride[0], sc = stride[1];
    const uint16_t *Y = planes[0] + (size_t)(2 * y0) * sl + 2 * x0;
    uint16_t *C[2] = { planes[1] + (size_t)y0 * sc + x0, planes[2] + (size_t)y0 * stride[2] + x0 };
    /* derive_availibilty_status, :57-69: any unit along the block's own width / height */
    int abv = ((abv_map >> ((xl >> 1) + 1)) & ((1ull << (w >> 1)) - 1)) != 0;
    int lft = ((lft_map >> ((yl >> 1) + 1)) & ((1ull << (h >> 1)) - 1)) != 0;
    int nb_abv = 0, nb_lft = 0, abv_step = 1, lft_step = 1;
    if (mode == 67) {                                              /* intra_cclm(_cl), :659-744 */
        if (abv) { nb_abv = imin(w, (abv + !lft) << 1); abv_step = w >> (abv + !lft); }
        if (lft) { nb_lft = imin(h, (lft + !abv) << 1); lft_step = h >> (lft + !abv); }
    } else if (mode == 69) {                                       /* intra_mdlm_abv(_cl), :746-815: above and above-right */
        if (abv) {
            int nb_units = (w + imin(h, w)) >> 1;
            uint64_t usable = (abv_map >> ((xl >> 1) + 1)) & ((1ull << nb_units) - 1);
            int len = __builtin_ctzll(~usable) << 1;
            nb_abv = imin(len, 4); abv_step = len >> 2;
        }
    } else {                                                       /* intra_mdlm_lft(_cl), :817-883: left and below-left */
        if (lft) {
            int nb_units = (h + imin(h, w)) >> 1;
            uint64_t usable = (lft_map >> ((yl >> 1) + 1)) & ((1ull << nb_units) - 1);
            int len = __builtin_ctzll(~usable) << 1;
            nb_lft = imin(len, 4); lft_step = len >> 2;
        }
    }
    if (abv_step < 1) abv_step = 1;
    if (lft_step < 1) lft_step = 1;
    int sy[4], sc_[2][4], n = 0;
    for (int i = 0; i < nb_abv; i++, n++) {                        /* sub_sample_lm_ref_abv0 / _abv(_cl), :188-212, 267-295, 588-624 */
        int x = (abv_step >> 1) + i * abv_step, pad_left = x == 0 && !lft;
        if (yl == 0) {                                             /* first line of the CTU: one luma row, [1 2 1] */
            const uint16_t *r = Y - sl + 2 * x;
            sy[n] = (r[pad_left ? 0 : -1] + 2 * r[0] + r[1] + 2) >> 2;
        } else sy[n] = lm_luma(Y - 2 * sl + 2 * x, sl, coll, pad_left, 0);
        sc_[0][n] = C[0][-sc + x]; sc_[1][n] = C[1][-(ptrdiff_t)stride[2] + x];
    }
    for (int i = 0; i < nb_lft; i++, n++) {                        /* sub_sample_lm_ref_lft(_cl), :214-265, 551-586 */
        int y = (lft_step >> 1) + i * lft_step;
        sy[n] = lm_luma(Y + (ptrdiff_t)(2 * y) * sl - 2, sl, coll, 0, y == 0 && !abv);
        sc_[0][n] = C[0][(ptrdiff_t)y * sc - 1]; sc_[1][n] = C[1][(ptrdiff_t)y * stride[2] - 1];
    }
    LmParams prm[2] = { { 0, 1 << (bitdepth - 1), 0 }, { 0, 1 << (bitdepth - 1), 0 } };
    if (n) {
        /* sort_average_lm_ref_samples, :137-186 */
        int min_l, max_l, min_c[2], max_c[2];
        if (n == 2) {
            int mi = sy[0] >= sy[1], ma = !mi;
            min_l = sy[mi]; max_l = sy[ma];
            for (int k = 0; k < 2; k++) { min_c[k] = sc_[k][mi]; max_c[k] = sc_[k][ma]; }
        } else {
            int idx[4] = { 0, 2, 1, 3 }, *mn = &idx[0], *mx = &idx[2], t;
            if (sy[0] > sy[2]) { t = mn[0]; mn[0] = mn[1]; mn[1] = t; }
            if (sy[1] > sy[3]) { t = mx[0]; mx[0] = mx[1]; mx[1] = t; }
            if (sy[mn[0]] > sy[mx[1]]) { int *tp = mn; mn = mx; mx = tp; }
            if (sy[mn[1]] > sy[mx[0]]) { t = mn[1]; mn[1] = mx[0]; mx[0] = t; }
            min_l = (sy[mn[0]] + sy[mn[1]] + 1) >> 1; max_l = (sy[mx[0]] + sy[mx[1]] + 1) >> 1;
            for (int k = 0; k < 2; k++) {
                min_c[k] = (sc_[k][mn[0]] + sc_[k][mn[1]] + 1) >> 1;
                max_c[k] = (sc_[k][mx[0]] + sc_[k][mx[1]] + 1) >> 1;
            }
        }
        /* derive_cclm_params, :108-135 */
        int range_l = max_l - min_l;
        for (int k = 0; k < 2; k++) { prm[k].a = 0; prm[k].b = min_c[k]; prm[k].shift = 0; }
        if (range_l) {
            int l2 = floor_log2u((unsigned)range_l), norm = ((range_l << 4) >> l2) & 15, v = div_lut[norm] | 8;
            l2 += norm != 0;
            for (int k = 0; k < 2; k++) prm[k] = lm_params_c(min_l, min_c[k], max_c[k], v, l2);
        }
    }
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            int lm = lm_luma(Y + (ptrdiff_t)(2 * j) * sl + 2 * i, sl, coll, i == 0 && !lft, j == 0 && !abv);
            for (int k = 0; k < 2; k++)
                C[k][(ptrdiff_t)j * stride[1 + k] + i] = (uint16_t)clipi(((lm * prm[k].a) >> prm[k].shift) + prm[k].b, 0, maxv);
        }
}

/* one block of one plane; chroma: x0 / y0 / log2 sizes in chroma samples */
static void intra_block(uint16_t *plane, int stride, int bitdepth, int x0, int y0, int xl, int yl, int lw, int lh, int mode_in,
                        uint64_t cols, uint64_t rows, int chroma, int ciip_wt, int mrl, int bdpcm)
{
    int w = 1 << lw, h = 1 << lh, maxv = (1 << bitdepth) - 1, us = chroma ? 1 : 2;
    uint16_t abv_buf[REF_LEN], lft_buf[REF_LEN], *abv = abv_buf + REF_OFF, *lft = lft_buf + REF_OFF;
    uint16_t *dst = plane + (size_t)y0 * stride + x0;
    const uint16_t *corner = dst - (ptrdiff_t)(mrl + 1) * stride - (mrl + 1);
    if (mrl) {          /* abv / lft point at raw + mrl, so that [1 + i] is the i-th sample of the reference line (:676-677) */
        fill_ref_mref(lft - mrl, corner, stride, 1, lh, lw, cols, rows, yl, xl, bitdepth, mrl, 1);
        fill_ref_mref(abv - mrl, corner, 1, stride, lw, lh, rows, cols, xl, yl, bitdepth, mrl, 0);
    } else {
        fill_ref(lft, corner, stride, 1, lh, lw, cols, rows, yl, xl, bitdepth, us, chroma);
        fill_ref(abv, corner, 1, stride, lw, lh, rows, cols, xl, yl, bitdepth, us, 0);
    }
    int *out = malloc(sizeof(int) * w * h);
    int pd = chroma ? (lw > 1 && lh > 1) : !mrl;
    if (bdpcm) {                                                   /* BDPCM block: cubic_{h,v}->pure on the plain references, nothing else
                                                                    * (vvc_intra_pred rcn_intra.c:511-522, vvc_intra_pred_chroma :1043-1074) */
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) out[y * w + x] = bdpcm == 2 ? abv[x + 1] : lft[y + 1];
    } else if (mode_in == 0) {                                            /* intra_planar(_pdpc), rcn_intra_dc_planar.c:106-149, 196-236 */
        uint16_t fa[REF_LEN], fl[REF_LEN];
        const uint16_t *A = abv, *L = lft;
        if (!chroma && !mrl && lw + lh > 5) { filter_ref(fa, abv, lft, w + 4); filter_ref(fl, lft, abv, h + 4); A = fa; L = fl; }
        int ws = lw > 1 ? lw : 1, hs = lh > 1 ? lh : 1, scale = (lw + lh - 2) >> 2;
        if (!pd) { ws = lw; hs = lh; }
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int hor = (L[y + 1] << ws) + (x + 1) * (A[w + 1] - L[y + 1]), ver = (A[x + 1] << hs) + (y + 1) * (L[h + 1] - A[x + 1]);
                int v = ((hor << hs) + (ver << ws) + (1 << (ws + hs))) >> (ws + hs + 1);
                int xw = pdpc_w(scale, x), yw = pdpc_w(scale, y);
                out[y * w + x] = pd ? clipi((xw * L[y + 1] + yw * A[x + 1] + (64 - (xw + yw)) * v + 32) >> 6, 0, maxv) : v;
            }
    } else if (mode_in == 1) {                                     /* intra_dc(_pdpc), :64-104, 151-194 */
        int shift = (lw > lh ? lw : lh) + (lw == lh), dc = 0, scale = (lw + lh - 2) >> 2;
        if (lw >= lh) for (int i = 0; i < w; i++) dc += abv[1 + i];
        if (lw <= lh) for (int i = 0; i < h; i++) dc += lft[1 + i];
        dc = (dc + ((1 << shift) >> 1)) >> shift;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int tw = pdpc_w(scale, x), lwg = pdpc_w(scale, y);
                out[y * w + x] = pd ? clipi((tw * lft[y + 1] + lwg * abv[x + 1] + (64 - (tw + lwg)) * dc + 32) >> 6, 0, maxv) : dc;
            }
    } else {
        /* derive_wide_angular_mode, rcn_intra.c:55-66 */
        static const uint8_t mode_shift_tab[6] = { 0, 6, 10, 12, 14, 15 };
        int mode = mode_in, ms = mode_shift_tab[abs(lw - lh)];
        if (lw > lh && mode < 2 + ms) mode += 65;
        else if (lh > lw && mode > 66 - ms) mode -= 65;
        if (mode >= 34) {
            if (mrl) mrl_class(out, abv, lft, lw, lh, mode - 50, maxv, mrl);
            else angular_class(out, abv, lft, abv[0], lw, lh, mode - 50, maxv, chroma);
        } else {
            int *t = malloc(sizeof(int) * w * h);
            if (mrl) mrl_class(t, lft, abv, lh, lw, 18 - mode, maxv, mrl);
            else angular_class(t, lft, abv, abv[0], lh, lw, 18 - mode, maxv, chroma);
            for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) out[y * w + x] = t[x * h + y];
            free(t);
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = out[y * w + x];
            if (ciip_wt) v = clipi((v * ciip_wt + dst[(size_t)y * stride + x] * (4 - ciip_wt) + 2) >> 2, 0, maxv);   /* rcn_mc.c:1611-1627 */
            dst[(size_t)y * stride + x] = (uint16_t)v;
        }
    free(out);
}

/* Matrix-based intra prediction: rcn_intra_mip -> vvc_intra_pred_mip(_tr) (libovvc/rcn_intra_mip.c:135-384): boundary
 * averaging to 2 / 4 samples per side (:165-183; the transposed variant puts the left boundary first, :283-300), offset
 * removal and the 4x4-class replacement of the first input by the mid value (:185-195), matrix product with the folded
 * rounding term 32 - 32 * sum (mip_matmult, :63-91), transposition of the reduced block for the _tr variant (:337-341),
 * then linear up-sampling, horizontally on the rows that hold the reduced lines and vertically over everything
 * (up_sample, :93-131), the first neighbour being the left / above reference sample. */
static void mip_block(uint16_t *plane, int stride, int bitdepth, int x0, int y0, int xl, int yl, int lw, int lh, int opaque,
                      uint64_t cols, uint64_t rows)
{
    int w = 1 << lw, maxv = (1 << bitdepth) - 1;
    uint16_t abv_buf[REF_LEN], lft_buf[REF_LEN], *abv = abv_buf + REF_OFF, *lft = lft_buf + REF_OFF;
    uint16_t *dst = plane + (size_t)y0 * stride + x0;
    const uint16_t *corner = dst - stride - 1;
    fill_ref(lft, corner, stride, 1, lh, lw, cols, rows, yl, xl, bitdepth, 2, 0);
    fill_ref(abv, corner, 1, stride, lw, lh, rows, cols, xl, yl, bitdepth, 2, 0);
    int transpose = (opaque >> 7) & 1, mode = opaque & 0x3f;
    int nb = (lw > 2 || lh > 2) ? 4 : 2, log2_nb = nb == 4 ? 2 : 1;
    int red_size = lh == 2 || lw == 2 || (lh <= 3 && lw <= 3), red = red_size ? 4 : 8, log2_red = red_size ? 2 : 3;
    int a_red[4], l_red[4], in[8];
    for (int j = 0; j < nb; j++) {
        int sa = 0, sl = 0, lbx = lw - log2_nb, lby = lh - log2_nb;
        for (int i = 0; i < (1 << lbx); i++) sa += abv[1 + i + (j << lbx)];
        for (int i = 0; i < (1 << lby); i++) sl += lft[1 + i + (j << lby)];
        a_red[j] = (sa + ((1 << lbx) >> 1)) >> lbx;
        l_red[j] = (sl + ((1 << lby) >> 1)) >> lby;
    }
    for (int j = 0; j < nb; j++) { in[j] = transpose ? l_red[j] : a_red[j]; in[nb + j] = transpose ? a_red[j] : l_red[j]; }
    int input_offset = in[0], sum = 0;
    if (red_size) in[0] = 1 << (bitdepth - 1);
    for (int i = 0; i < 2 * nb; i++) { in[i] = (int16_t)(in[i] - input_offset); sum += in[i]; }
    int rnd = 32 - 32 * sum;
    const uint8_t *m = (lw == 2 && lh == 2) ? &mip_w_4x4[mode * 16 * 4] : (red_size ? &mip_w_8x8[mode * 16 * 8] : &mip_w_16x16[mode * 64 * 8]);
    int pred[64], P[64];
    for (int pos = 0; pos < red * red; pos++) {
        int v = 0;
        for (int i = 0; i < 2 * nb; i++) v += in[i] * m[pos * 2 * nb + i];
        pred[pos] = clipi(((v + rnd) >> 6) + input_offset, 0, maxv);
    }
    for (int i = 0; i < red; i++) for (int j = 0; j < red; j++) P[i * red + j] = transpose ? pred[j * red + i] : pred[i * red + j];
    int lsx = lw - log2_red, lsy = lh - log2_red, sx = 1 << lsx, sy = 1 << lsy;
    int *H = malloc(sizeof(int) * red * w);                       /* reduced rows after the horizontal pass */
    for (int i = 0; i < red; i++) {
        int before = lft[1 + (i + 1) * sy - 1];
        for (int j = 0; j < red; j++) {
            int after = P[i * red + j];
            for (int pos = 1; pos <= sx; pos++)
                H[i * w + j * sx + pos - 1] = lsx ? (before * (sx - pos) + after * pos + (sx >> 1)) >> lsx : after;
            before = after;
        }
    }
    for (int x = 0; x < w; x++) {
        int before = abv[1 + x];
        for (int j = 0; j < red; j++) {
            int after = H[j * w + x];
            for (int pos = 1; pos <= sy; pos++)
                dst[(size_t)(j * sy + pos - 1) * stride + x] = (uint16_t)(lsy ? (before * (sy - pos) + after * pos + (sy >> 1)) >> lsy : after);
            before = after;
        }
    }
    free(H);
}

/* vvc_intra_pred_isp (rcn_intra.c:567-643): prediction of one partition (lpw x lph, at offset (ox, oy) inside the coding
 * block of the command) from references filled for the whole coding block */
static void isp_pred_block(uint16_t *plane, int stride, int bitdepth, int cm, const OVOIntraCmd *c, uint64_t cols, uint64_t rows,
                           int ox, int oy, int lpw, int lph)
{
    int lcw = c->log2_w, lch = c->log2_h, w = 1 << lpw, h = 1 << lph, maxv = (1 << bitdepth) - 1;
    int x0 = c->x0 + ox, y0 = c->y0 + oy;
    uint16_t abv_buf[REF_LEN + 64], lft_buf[REF_LEN + 64], *abv = abv_buf + REF_OFF, *lft = lft_buf + REF_OFF;
    uint16_t *dst = plane + (size_t)y0 * stride + x0;
    /* left: column x0 - 1 from the block's top - 1 down; above: row y0 - 1 from the block's left - 1 on */
    fill_ref_off(lft, plane + (size_t)(c->y0 - 1) * stride + x0 - 1, stride, 1, lch, lcw, cols, rows, c->y0 & cm, x0 & cm, bitdepth, 2, 0, oy);
    fill_ref_off(abv, plane + (size_t)(y0 - 1) * stride + c->x0 - 1, 1, stride, lcw, lch, rows, cols, c->x0 & cm, y0 & cm, bitdepth, 2, 0, ox);
    abv += ox; lft += oy;
    for (int i = 0; i < 4; i++) lft[(1 << lch) + h + 1 + i] = lft[(1 << lch) + h + i];       /* rcn_intra.c:601-607 */
    for (int i = 0; i < 4; i++) abv[(1 << lcw) + w + 1 + i] = abv[(1 << lcw) + w + i];
    int *out = malloc(sizeof(int) * w * h);
    int pd = lph > 1;
    if (c->mode == 0) {
        int ws = pd ? (lpw > 1 ? lpw : 1) : lpw, hs = pd ? (lph > 1 ? lph : 1) : lph, scale = (lpw + lph - 2) >> 2;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int hor = (lft[y + 1] << ws) + (x + 1) * (abv[w + 1] - lft[y + 1]), ver = (abv[x + 1] << hs) + (y + 1) * (lft[h + 1] - abv[x + 1]);
                int v = ((hor << hs) + (ver << ws) + (1 << (ws + hs))) >> (ws + hs + 1);
                int xw = pdpc_w(scale, x), yw = pdpc_w(scale, y);
                out[y * w + x] = pd ? clipi((xw * lft[y + 1] + yw * abv[x + 1] + (64 - (xw + yw)) * v + 32) >> 6, 0, maxv) : v;
            }
    } else if (c->mode == 1) {
        int shift = (lpw > lph ? lpw : lph) + (lpw == lph), dc = 0, scale = (lpw + lph - 2) >> 2;
        if (lpw >= lph) for (int i = 0; i < w; i++) dc += abv[1 + i];
        if (lpw <= lph) for (int i = 0; i < h; i++) dc += lft[1 + i];
        dc = (dc + ((1 << shift) >> 1)) >> shift;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int tw = pdpc_w(scale, x), lwg = pdpc_w(scale, y);
                out[y * w + x] = pd ? clipi((tw * lft[y + 1] + lwg * abv[x + 1] + (64 - (tw + lwg)) * dc + 32) >> 6, 0, maxv) : dc;
            }
    } else {
        static const uint8_t mode_shift_tab[6] = { 0, 6, 10, 12, 14, 15 };
        int mode = c->mode, ms = mode_shift_tab[abs(lcw - lch)];                   /* wide angles from the CODING BLOCK's shape */
        if (lcw > lch && mode < 2 + ms) mode += 65;
        else if (lch > lcw && mode > 66 - ms) mode -= 65;
        if (mode >= 34) angular_class2(out, abv, lft, abv[0], lpw, lph, mode - 50, maxv, 0, lph);
        else {
            int *t = malloc(sizeof(int) * w * h);
            angular_class2(t, lft, abv, abv[0], lph, lpw, 18 - mode, maxv, 0, lph);
            for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) out[y * w + x] = t[x * h + y];
            free(t);
        }
    }
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[(size_t)y * stride + x] = (uint16_t)out[y * w + x];
    free(out);
}

/* recon_isp_subtree_v / _h (rcn_transform_tree.c:1087-1205): partitions in order, each predicted (vertical split: in groups
 * of at least 4 columns) and then reconstructed with its own transform block */
void ovo_isp_cu(uint16_t *plane, int stride, int bitdepth, int cm, const OVOIntraCmd *c, const OVOIspInfo *info,
                const int16_t *arena, int dep_quant)
{
    int lcw = c->log2_w, lch = c->log2_h, ver = (c->flags & OVO_INTRA_ISP_VER) != 0, maxv = (1 << bitdepth) - 1;
    int lp = (ver ? lcw : lch) - 2, lo = ver ? lch : lcw;                          /* partition size along the split, the other side */
    if (lo < 4 && lp <= 4 - lo) lp = 4 - lo;
    int nb = (1 << (ver ? lcw : lch)) >> lp, q = 0;
    int16_t res[64 * 64];
    for (int i = 0; i < nb; i++) {
        int off = i << lp;
        if (ver) { if (!(off & 3)) { isp_pred_block(plane, stride, bitdepth, cm, c, info->part[q], info->common, off, 0, lp < 2 ? 2 : lp, lch); q++; } }
        else     { isp_pred_block(plane, stride, bitdepth, cm, c, info->common, info->part[q], 0, off, lcw, lp); q++; }
        const OVOTuCmd *t = &info->tu[i];
        if (!(t->cbf_mask & 0x10)) continue;
        ovo_isp_residual(res, arena, t, dep_quant, bitdepth);
        int tw = 1 << t->log2_w, th = 1 << t->log2_h;
        for (int y = 0; y < th; y++)
            for (int x = 0; x < tw; x++) {
                uint16_t *d = plane + (size_t)(t->y0 + y) * stride + t->x0 + x;
                *d = (uint16_t)clipi(*d + res[y * tw + x], 0, maxv);
            }
    }
}

void ovo_intra_picture2(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                        const OVOIntraCmd *cmds, int nb_cmds, const OVOIspInfo *isp, const int16_t *arena, int dep_quant)
{
    int cm = (1 << log2_ctu_s) - 1;
    for (int i = 0; i < nb_cmds; i++) {
        const OVOIntraCmd *c = &cmds[i];
        if (c->flags & OVO_INTRA_ISP) {
            ovo_isp_cu(planes[0], stride[0], bitdepth, cm, c, &isp[c->pad[0] | (c->pad[1] << 8)], arena, dep_quant);
            if (c->flags & OVO_INTRA_CHROMA) {          /* the chroma blocks of the CU are ordinary ones */
                OVOIntraCmd k = *c;
                k.flags &= ~(OVO_INTRA_LUMA | OVO_INTRA_ISP | OVO_INTRA_ISP_VER);
                ovo_intra_picture(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, &k, 1);
            }
        } else ovo_intra_picture(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, c, 1);
    }
}

/* rcn_ibc_l / rcn_ibc_c (rcn_ibc.c:8-140).  The reference copies inside its CTU ring buffer (256 * 128 >> log2_ctu_s luma columns:
 * the current CTU and the ones decoded before it in the row, addressed modulo the ring width; rows modulo the CTU height).  For a
 * block vector that points at samples already reconstructed in the current CTU or in the CTUs the ring still holds -- the only ones
 * a conforming stream uses -- the ring position is the picture position (x0 + bv.x, y0 + bv.y), which is what is restated here.
 * Chroma: position ((x0 + bv.x) >> 1, (y0 + bv.y) >> 1), both planes (:84-89, :107-112). */
static void ibc_block(uint16_t *const planes[3], const int stride[3], const OVOIntraCmd *c)
{
    const int bvx = (int16_t)(c->pad[0] | (c->pad[1] << 8)), bvy = (int16_t)(c->pad[2] | (c->pad[3] << 8));
    const int w = 1 << c->log2_w, h = 1 << c->log2_h;
    if (c->flags & OVO_INTRA_LUMA)
        for (int y = 0; y < h; y++)
            memcpy(planes[0] + (size_t)(c->y0 + y) * stride[0] + c->x0, planes[0] + (size_t)(c->y0 + bvy + y) * stride[0] + c->x0 + bvx, 2 * w);
    if (c->flags & OVO_INTRA_CHROMA)
        for (int k = 1; k < 3; k++)
            for (int y = 0; y < h / 2; y++)
                memcpy(planes[k] + (size_t)((c->y0 >> 1) + y) * stride[k] + (c->x0 >> 1),
                       planes[k] + (size_t)(((c->y0 + bvy) >> 1) + y) * stride[k] + ((c->x0 + bvx) >> 1), w);
}

void ovo_intra_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                       const OVOIntraCmd *cmds, int nb_cmds)
{
    int cm = (1 << log2_ctu_s) - 1;
    (void)pic_w; (void)pic_h;
    for (int i = 0; i < nb_cmds; i++) {
        const OVOIntraCmd *c = &cmds[i];
        if (c->flags & OVO_INTRA_IBC) { ibc_block(planes, stride, c); continue; }
        int ciip = (c->flags & OVO_INTRA_CIIP) ? c->ciip_wt : 0;
        if ((c->flags & OVO_INTRA_LUMA) && (c->flags & OVO_INTRA_MIP))
            mip_block(planes[0], stride[0], bitdepth, c->x0, c->y0, c->x0 & cm, c->y0 & cm, c->log2_w, c->log2_h, c->mode, c->avail_cols, c->avail_rows);
        else if (c->flags & OVO_INTRA_LUMA)
            intra_block(planes[0], stride[0], bitdepth, c->x0, c->y0, c->x0 & cm, c->y0 & cm, c->log2_w, c->log2_h, ciip ? 0 : c->mode,
                        c->avail_cols, c->avail_rows, 0, ciip, ciip ? 0 : c->mrl_idx, (!ciip && (c->pad[4] & 1)) ? 1 + ((c->pad[4] >> 1) & 1) : 0);
        if ((c->flags & OVO_INTRA_CHROMA) && !ciip && c->mode_c >= 67)
            cclm_block(planes, stride, bitdepth, c->x0 >> 1, c->y0 >> 1, (c->x0 & cm) >> 1, (c->y0 & cm) >> 1, c->log2_w - 1, c->log2_h - 1,
                       c->mode_c, c->avail_cols, c->avail_rows, (c->flags & OVO_INTRA_CCLM_COLLOCATED) != 0);
        else if ((c->flags & OVO_INTRA_CHROMA) && !(ciip && c->log2_w <= 2))    /* vvc_intra_pred_chroma (rcn_intra.c:1015-1182): Cb then Cr */
            for (int k = 1; k < 3; k++)
                intra_block(planes[k], stride[k], bitdepth, c->x0 >> 1, c->y0 >> 1, (c->x0 & cm) >> 1, (c->y0 & cm) >> 1,
                            c->log2_w - 1, c->log2_h - 1, ciip ? 0 : c->mode_c, c->avail_cols, c->avail_rows, 1, ciip, 0,
                            (!ciip && (c->pad[4] & 4)) ? 1 + ((c->pad[4] >> 3) & 1) : 0);
    }
}
