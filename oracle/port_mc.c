/* oracle/port_mc.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of OpenVVC's translational motion compensation: dispatch (rcn_mcp_b & co,
 * libovvc/rcn_inter.c:2750-2966), MV clipping / fractional split / reference addressing
 * (rcn_inter.c:96-108, 521-602, 1392-1476, 1479-1554, 1822-1900), border replication
 * (emulate_block_border :148-225 == clamped coordinates) and the separable interpolation kernels
 * (libovvc/rcn_mc.c:381-1098 luma, :1100-1376 chroma, :1380-1608 BCW-weighted).
 * Plus the LMCS LUTs (libovvc/rcn_lmcs.c:84-187) because every luma inter prediction is forward-reshaped
 * inside the MC routines (rcn_inter.c:599, 1552).
 */
#include "oracle.h"
#include <string.h>
#include <stdlib.h>

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* VVC luma 1/16-sample filters; row 15 = the half-sample filter selected by AMVR half-pel (rcn_mc.c:103-123) */
static const int8_t mc_filt_l[16][8] = {
    {  0, 1,  -3, 63,  4,  -2,  1,  0 }, { -1, 2,  -5, 62,  8,  -3,  1,  0 }, { -1, 3,  -8, 60, 13,  -4,  1,  0 },
    { -1, 4, -10, 58, 17,  -5,  1,  0 }, { -1, 4, -11, 52, 26,  -8,  3, -1 }, { -1, 3,  -9, 47, 31, -10,  4, -1 },
    { -1, 4, -11, 45, 34, -10,  4, -1 }, { -1, 4, -11, 40, 40, -11,  4, -1 }, { -1, 4, -10, 34, 45, -11,  4, -1 },
    { -1, 4, -10, 31, 47,  -9,  3, -1 }, { -1, 3,  -8, 26, 52, -11,  4, -1 }, {  0, 1,  -5, 17, 58, -10,  4, -1 },
    {  0, 1,  -4, 13, 60,  -8,  3, -1 }, {  0, 1,  -3,  8, 62,  -5,  2, -1 }, {  0, 1,  -2,  4, 63,  -3,  1,  0 },
    {  0, 3,   9, 20, 20,   9,  3,  0 } };
/* 6-tap variant used for 4x4 (affine sub-)blocks (rcn_mc.c:125-142) */
static const int8_t mc_filt_l4[15][8] = {
    { 0, 1,  -3, 63,  4,  -2, 1, 0 }, { 0, 1,  -5, 62,  8,  -3, 1, 0 }, { 0, 2,  -8, 60, 13,  -4, 1, 0 },
    { 0, 3, -10, 58, 17,  -5, 1, 0 }, { 0, 3, -11, 52, 26,  -8, 2, 0 }, { 0, 2,  -9, 47, 31, -10, 3, 0 },
    { 0, 3, -11, 45, 34, -10, 3, 0 }, { 0, 3, -11, 40, 40, -11, 3, 0 }, { 0, 3, -10, 34, 45, -11, 3, 0 },
    { 0, 3, -10, 31, 47,  -9, 2, 0 }, { 0, 2,  -8, 26, 52, -11, 3, 0 }, { 0, 1,  -5, 17, 58, -10, 3, 0 },
    { 0, 1,  -4, 13, 60,  -8, 2, 0 }, { 0, 1,  -3,  8, 62,  -5, 1, 0 }, { 0, 1,  -2,  4, 63,  -3, 1, 0 } };
/* chroma 1/32-sample filters (rcn_mc.c:67-101) */
static const int8_t mc_filt_c[31][4] = {
    { -1, 63,  2,  0 }, { -2, 62,  4,  0 }, { -2, 60,  7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
    { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 },
    { -6, 49, 24, -3 }, { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 },
    { -4, 36, 36, -4 }, { -4, 33, 39, -4 }, { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 },
    { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 }, { -2, 16, 54, -4 }, { -2, 15, 55, -4 },
    { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1,  7, 60, -2 }, {  0,  4, 62, -2 },
    {  0,  2, 63, -1 } };
static const int8_t bcw_w[5] = { -2, 3, 4, 5, 10 };   /* rcn_inter.c:89 */

typedef struct { const uint16_t *p; int stride, w, h; } RefPlane;

static inline int rpx(const RefPlane *r, int x, int y)
{   /* emulate_block_border (rcn_inter.c:148-225): samples outside the picture repeat the border */
    return r->p[(size_t)clipi(y, 0, r->h - 1) * r->stride + clipi(x, 0, r->w - 1)];
}

/* 14-bit intermediate prediction of one block (put_vvc_pel_pixels / qpel_h / qpel_v / qpel_hv and the
 * epel equivalents, rcn_mc.c:401-419, 901-997, 1208-1300).  ntaps = 8 luma / 4 chroma; fx, fy = filters or NULL. */
static void pred14(int16_t *dst, int dst_stride, const RefPlane *r, int x_int, int y_int, int w, int h,
                   const int8_t *fx, const int8_t *fy, int ntaps, int bitdepth)
{
    int before = ntaps / 2 - 1;
    int sh1 = bitdepth - 8;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v;
            if (!fx && !fy) {
                v = rpx(r, x_int + x, y_int + y) << (14 - bitdepth);
            } else if (fx && !fy) {
                int s = 0;
                for (int k = 0; k < ntaps; k++) s += fx[k] * rpx(r, x_int + x + k - before, y_int + y);
                v = s >> sh1;
            } else if (!fx && fy) {
                int s = 0;
                for (int k = 0; k < ntaps; k++) s += fy[k] * rpx(r, x_int + x, y_int + y + k - before);
                v = s >> sh1;
            } else {
                int s2 = 0;
                for (int j = 0; j < ntaps; j++) {
                    int s = 0;
                    for (int k = 0; k < ntaps; k++)
                        s += fx[k] * rpx(r, x_int + x + k - before, y_int + y + j - before);
                    s2 += fy[j] * (int16_t)(s >> sh1);     /* tmp[] is int16_t, rcn_mc.c:506,515 */
                }
                v = s2 >> 6;
            }
            dst[y * dst_stride + x] = (int16_t)v;
        }
}

typedef struct { int x, y; } MV;

static MV clip_mv(int pos_x, int pos_y, int pic_w, int pic_h, int pb_w, int pb_h, MV mv)
{   /* rcn_inter.c:96-108 */
    int x_max = (pic_w + 2 - pos_x) << 4, y_max = (pic_h + 2 - pos_y) << 4;
    int x_min = -((pb_w + 3 + pos_x) << 4), y_min = -((pb_h + 3 + pos_y) << 4);
    mv.x = clipi(mv.x, x_min, x_max);
    mv.y = clipi(mv.y, y_min, y_max);
    return mv;
}

/* prediction of one component of one PU from one reference into a 14-bit buffer */
static void pu_pred14(int16_t *buf, const OVOMcRefs *refs, int list, int ref_idx, int comp,
                      int pic_w, int pic_h, const OVOMcCmd *c, MV mv, int bitdepth)
{
    int pu_w = 1 << c->log2_w, pu_h = 1 << c->log2_h;
    mv = clip_mv(c->x0, c->y0, pic_w, pic_h, pu_w, pu_h, mv);
    if (comp == 0) {
        RefPlane r = { refs->plane[list][ref_idx][0], refs->stride[list][ref_idx][0], pic_w, pic_h };
        int px = mv.x & 15, py = mv.y & 15;
        int small = (pu_w == 4 && pu_h == 4);                                              /* rcn_mc.c:457 */
        /* half-pel AMVR never coexists with 4x4 (affine) blocks in a conforming stream; the reference would
         * index ov_mc_filters_4[15] out of bounds there, so the combination is left undefined/ignored */
        if ((c->flags & OVO_MC_HPEL_AMVR) && !small) { if (px == 8) px = 16; if (py == 8) py = 16; }   /* rcn_inter.c:572-577 */
        const int8_t *fx = px ? (small ? mc_filt_l4[px - 1] : mc_filt_l[px - 1]) : NULL;
        const int8_t *fy = py ? (small ? mc_filt_l4[py - 1] : mc_filt_l[py - 1]) : NULL;
        pred14(buf, 128, &r, c->x0 + (mv.x >> 4), c->y0 + (mv.y >> 4), pu_w, pu_h, fx, fy, 8, bitdepth);
    } else {
        RefPlane r = { refs->plane[list][ref_idx][comp], refs->stride[list][ref_idx][comp], pic_w >> 1, pic_h >> 1 };
        int px = mv.x & 31, py = mv.y & 31;
        pred14(buf, 128, &r, (c->x0 >> 1) + (mv.x >> 5), (c->y0 >> 1) + (mv.y >> 5), pu_w >> 1, pu_h >> 1,
               px ? mc_filt_c[px - 1] : NULL, py ? mc_filt_c[py - 1] : NULL, 4, bitdepth);
    }
}

/* ---------------------------------------------------------------------------------------------------
 * BDOF (rcn_bdof_mcp_l, rcn_inter.c:1137-1250; libovvc/rcn_prof_bdof.c:60-99, 295-492) and PROF
 * (rcn_prof_mcp_b_l / rcn_prof_motion_compensation_b_l / rcn_prof_mcp_l, rcn_inter.c:1253-1388, 1632-1727,
 * 2865-2919; rcn_prof_bdof.c:101-282)
 * ------------------------------------------------------------------------------------------------- */
enum { PS = 20 };      /* stride of the padded 14-bit prediction / gradient blocks (<= 16 + 2) */

/* 14-bit prediction of a w x h luma block into the interior of a padded block, plus the one-sample ring taken from
 * the nearest INTEGER reference samples (extend_bdof_buff / extend_prof_buff, rcn_prof_bdof.c:295-345, 173-221) */
static void padded_pred14(int16_t *P, const OVOMcRefs *refs, int list, int ref_idx, int pic_w, int pic_h,
                          int x0, int y0, int w, int h, MV mv, int hpel, int bitdepth)
{
    RefPlane r = { refs->plane[list][ref_idx][0], refs->stride[list][ref_idx][0], pic_w, pic_h };
    mv = clip_mv(x0, y0, pic_w, pic_h, w, h, mv);
    int px = mv.x & 15, py = mv.y & 15, small = (w == 4 && h == 4);
    if (hpel && !small) { if (px == 8) px = 16; if (py == 8) py = 16; }
    const int8_t *fx = px ? (small ? mc_filt_l4[px - 1] : mc_filt_l[px - 1]) : NULL;
    const int8_t *fy = py ? (small ? mc_filt_l4[py - 1] : mc_filt_l[py - 1]) : NULL;
    int xi = x0 + (mv.x >> 4), yi = y0 + (mv.y >> 4);
    pred14(P + PS + 1, PS, &r, xi, yi, w, h, fx, fy, 8, bitdepth);
    int ex = (px >> 3) != 0, ey = (py >> 3) != 0, sh = 14 - bitdepth;
    for (int i = 0; i < w + 2; i++) {
        P[i] = (int16_t)(rpx(&r, xi - 1 + i + ex, yi - 1 + ey) << sh);
        P[(h + 1) * PS + i] = (int16_t)(rpx(&r, xi - 1 + i + ex, yi + h + ey) << sh);
    }
    for (int j = 0; j < h; j++) {
        P[(j + 1) * PS] = (int16_t)(rpx(&r, xi - 1 + ex, yi + j + ey) << sh);
        P[(j + 1) * PS + w + 1] = (int16_t)(rpx(&r, xi + w + ex, yi + j + ey) << sh);
    }
}

/* compute_prof_grad, rcn_prof_bdof.c:153-171: gradients of the interior from the padded block */
static void pad_grad(int16_t *gx, int16_t *gy, const int16_t *P, int w, int h)
{
    for (int y = 1; y <= h; y++)
        for (int x = 1; x <= w; x++) {
            gy[y * PS + x] = (int16_t)(((P[(y + 1) * PS + x] - 8192) >> 6) - ((P[(y - 1) * PS + x] - 8192) >> 6));
            gx[y * PS + x] = (int16_t)(((P[y * PS + x + 1] - 8192) >> 6) - ((P[y * PS + x - 1] - 8192) >> 6));
        }
}

/* extend_bdof_grad, rcn_inter.c:840-870: ring = nearest interior sample (columns first, then whole rows) */
static void pad_replicate(int16_t *B, int w, int h)
{
    for (int y = 1; y <= h; y++) { B[y * PS] = B[y * PS + 1]; B[y * PS + w + 1] = B[y * PS + w]; }
    for (int x = 0; x < w + 2; x++) { B[x] = B[PS + x]; B[(h + 1) * PS + x] = B[h * PS + x]; }
}

static int floor_log2u(unsigned v) { int n = -1; while (v) { v >>= 1; n++; } return n; }

/* BDOF on two padded 14-bit predictions whose ring holds the integer reference samples (rcn_bdof_mcp_l tail,
 * rcn_inter.c:1215-1244; also the BDOF branch of rcn_dmvr_mv_refine, :1018-1056) */
static void bdof_core(uint16_t *d, int stride, int16_t *P0, int16_t *P1, int w, int h, int bitdepth)
{
    int maxv = (1 << bitdepth) - 1;
    int16_t gx0[PS * PS], gy0[PS * PS], gx1[PS * PS], gy1[PS * PS];
    pad_grad(gx0, gy0, P0, w, h); pad_grad(gx1, gy1, P1, w, h);
    pad_replicate(gx0, w, h); pad_replicate(gy0, w, h); pad_replicate(gx1, w, h); pad_replicate(gy1, w, h);
    pad_replicate(P0, w, h); pad_replicate(P1, w, h);          /* "reference padding overwrite", rcn_inter.c:1233-1235 */
    int shift = 15 - bitdepth, off = 1 << (shift - 1);
    for (int sy = 0; sy < h; sy += 4)
        for (int sx = 0; sx < w; sx += 4) {
            /* derive_bdof_weights over the 6x6 neighbourhood, rcn_prof_bdof.c:347-422 */
            int sum_ax = 0, sum_ay = 0, sum_dx = 0, sum_dy = 0, sum_sign = 0, wx = 0, wy = 0;
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++) {
                    int o = (sy + i) * PS + sx + j;
                    int ax = (gx0[o] + gx1[o]) >> 1, ay = (gy0[o] + gy1[o]) >> 1;
                    int dr = ((P1[o] - 8192) >> 4) - ((P0[o] - 8192) >> 4);
                    sum_ax += abs(ax); sum_ay += abs(ay);
                    sum_sign += ay < 0 ? -ax : (ay == 0 ? 0 : ax);
                    sum_dx += ax < 0 ? -dr : (ax == 0 ? 0 : dr);
                    sum_dy += ay < 0 ? -dr : (ay == 0 ? 0 : dr);
                }
            if (sum_ax) wx = clipi((sum_dx << 2) >> floor_log2u(sum_ax), -15, 15);
            if (sum_ay) {
                int xo = 0;
                if (wx) { int hi = sum_sign >> 12, lo = sum_sign & 4095; xo = (((wx * hi) << 12) + (wx * lo)) >> 1; }
                wy = clipi(((sum_dy << 2) - xo) >> floor_log2u(sum_ay), -15, 15);
            }
            for (int y = 0; y < 4; y++)                         /* rcn_apply_bdof_subblock, :60-99 */
                for (int x = 0; x < 4; x++) {
                    int o = (sy + y + 1) * PS + sx + x + 1;
                    int b = wx * (gx0[o] - gx1[o]) + wy * (gy0[o] - gy1[o]);
                    int16_t v = (int16_t)((P0[o] + P1[o] + b + off) >> shift);
                    d[(size_t)(sy + y) * stride + sx + x] = (uint16_t)clipi(v, 0, maxv);
                }
        }
}

static void bdof_block(uint16_t *d, int stride, const OVOMcRefs *refs, const OVOMcCmd *c, int pic_w, int pic_h, int bitdepth)
{
    int w = 1 << c->log2_w, h = 1 << c->log2_h;
    int16_t P0[PS * PS], P1[PS * PS];
    MV mv0 = { c->mv0_x, c->mv0_y }, mv1 = { c->mv1_x, c->mv1_y };
    int hpel = (c->flags & OVO_MC_HPEL_AMVR) != 0;
    padded_pred14(P0, refs, 0, c->ref_idx0, pic_w, pic_h, c->x0, c->y0, w, h, mv0, hpel, bitdepth);
    padded_pred14(P1, refs, 1, c->ref_idx1, pic_w, pic_h, c->x0, c->y0, w, h, mv1, hpel, bitdepth);
    bdof_core(d, stride, P0, P1, w, h, bitdepth);
}

/* ---------------------------------------------------------------------------------------------------
 * DMVR (rcn_dmvr_mv_refine, rcn_inter.c:873-1126)
 * ------------------------------------------------------------------------------------------------- */
/* The DMVR reference window: the (w + 7) x (h + 7) samples around the block at the CLIPPED integer MV (picture
 * borders replicated, derive_dmvr_ref_buf_y :432-471), itself extended by replication (padd_dmvr :354-378):
 * sample at window coordinates (x, y), x in -5 .. w + 5 */
typedef struct { RefPlane r; int bx, by, w, h, before, after; } DmvrWin;
static inline int dwin(const DmvrWin *d, int x, int y)
{
    return rpx(&d->r, d->bx + clipi(x, -d->before, d->w + d->after), d->by + clipi(y, -d->before, d->h + d->after));
}

static int div_for_maxq7(int64_t num, int64_t den)
{   /* rcn_inter.c:758-797 */
    int sign = 0, q = 0;
    if (num < 0) { sign = 1; num = -num; }
    den <<= 3;
    if (num >= den) { num -= den; q++; }
    q <<= 1; den >>= 1;
    if (num >= den) { num -= den; q++; }
    q <<= 1;
    if (num >= (den >> 1)) q++;
    return sign ? -q : q;
}

static void dmvr_block(uint16_t *const dst[3], const int dst_stride[3], const OVOMcRefs *refs, const OVOMcCmd *c,
                       int pic_w, int pic_h, int bitdepth, OVODmvrMv *out)
{
    int w = 1 << c->log2_w, h = 1 << c->log2_h, maxv = (1 << bitdepth) - 1;
    MV mv[2] = { { c->mv0_x, c->mv0_y }, { c->mv1_x, c->mv1_y } }, org[2];
    int ridx[2] = { c->ref_idx0, c->ref_idx1 };
    DmvrWin win[2];
    uint16_t bl[2][20 * 20];
    org[0] = mv[0]; org[1] = mv[1];
    for (int r = 0; r < 2; r++) {
        MV mc = clip_mv(c->x0, c->y0, pic_w, pic_h, w, h, mv[r]);
        win[r] = (DmvrWin){ { refs->plane[r][ridx[r]][0], refs->stride[r][ridx[r]][0], pic_w, pic_h },
                            c->x0 + (mc.x >> 4), c->y0 + (mc.y >> 4), w, h, 3, 3 };
        /* bilinear interpolation with two extra samples on every side (put_vvc_*_bilinear*, rcn_mc.c:789-897);
         * the fraction comes from the UNCLIPPED MV (rcn_inter.c:911-915) */
        int fx = mv[r].x & 15, fy = mv[r].y & 15;
        for (int i = 0; i < h + 4; i++)
            for (int j = 0; j < w + 4; j++) {
                int x = j - 2, y = i - 2, v;
                if (!fx && !fy) v = dwin(&win[r], x, y) << (10 - bitdepth);
                else if (fx && !fy) v = ((16 - fx) * dwin(&win[r], x, y) + fx * dwin(&win[r], x + 1, y) + 8) >> 4;
                else if (!fx && fy) v = ((16 - fy) * dwin(&win[r], x, y) + fy * dwin(&win[r], x, y + 1) + 8) >> 4;
                else {
                    int t0 = (int16_t)(((16 - fx) * dwin(&win[r], x, y) + fx * dwin(&win[r], x + 1, y) + 8) >> 4);
                    int t1 = (int16_t)(((16 - fx) * dwin(&win[r], x, y + 1) + fx * dwin(&win[r], x + 1, y + 1) + 8) >> 4);
                    v = ((16 - fy) * t0 + fy * t1 + 8) >> 4;
                }
                bl[r][i * 20 + j] = (uint16_t)v;
            }
    }
    /* SAD on every other row (rcn_dmvr_sad_{8,16}, :615-665) */
    uint64_t sad[25];
    for (int idx = 0; idx < 25; idx++) {
        int dx = idx % 5 - 2, dy = idx / 5 - 2;
        uint64_t s = 0;
        for (int i = 0; i < h; i += 2)
            for (int j = 0; j < w; j++)
                s += abs((int16_t)bl[0][(i + 2 + dy) * 20 + j + 2 + dx] - (int16_t)bl[1][(i + 2 - dy) * 20 + j + 2 - dx]);
        sad[idx] = s;
    }
    uint64_t min_cost = sad[12] - (sad[12] >> 2);
    if (min_cost >= (uint64_t)(w * h)) {
        sad[12] = min_cost;
        int best = 12;
        uint64_t mc = (uint64_t)-1;
        for (int idx = 0; idx < 25; idx++)                      /* dmvr_compute_sads_*, :700-706 */
            if (sad[idx] < mc || (idx == 12 && sad[idx] <= mc)) { mc = sad[idx]; best = idx; }
        int dh = (best % 5 - 2) << 4, dv = (best / 5 - 2) << 4;
        min_cost = sad[best];
        if (abs(dh) != 32 && abs(dv) != 32) {                   /* refine_mv, :799-833 */
            const uint64_t *sb = &sad[best];
            int64_t den_h = (int64_t)(sb[-1] + sb[1] - (sb[0] << 1)), den_v = (int64_t)(sb[-5] + sb[5] - (sb[0] << 1));
            if (den_h) {
                if (sb[-1] != sb[0] && sb[1] != sb[0]) dh += div_for_maxq7((int64_t)(sb[-1] << 4) - (int64_t)(sb[1] << 4), den_h);
                else dh += sb[-1] == sb[0] ? -8 : 8;
            }
            if (den_v) {
                if (sb[-5] != sb[0] && sb[5] != sb[0]) dv += div_for_maxq7((int64_t)((sb[-5] - sb[5]) << 4), den_v);
                else dv += sb[-5] == sb[0] ? -8 : 8;
            }
        }
        mv[0].x = clipi(mv[0].x + dh, -(1 << 17), (1 << 17) - 1); mv[0].y = clipi(mv[0].y + dv, -(1 << 17), (1 << 17) - 1);
        mv[1].x = clipi(mv[1].x - dh, -(1 << 17), (1 << 17) - 1); mv[1].y = clipi(mv[1].y - dv, -(1 << 17), (1 << 17) - 1);
    }
    if (out) { out->mv0_x = mv[0].x; out->mv0_y = mv[0].y; out->mv1_x = mv[1].x; out->mv1_y = mv[1].y; }
    int disable_bdof = (c->aux & 1) ? min_cost < (uint64_t)(2 * w * h) : 1;
    /* final 8-tap prediction from the padded window, shifted by the integer part of the refinement */
    int hpel = (c->flags & OVO_MC_HPEL_AMVR) != 0;
    int16_t P[2][PS * PS];
    for (int r = 0; r < 2; r++) {
        int px = mv[r].x & 15, py = mv[r].y & 15;
        if (hpel) { if (px == 8) px = 16; if (py == 8) py = 16; }
        const int8_t *fx = px ? mc_filt_l[px - 1] : NULL, *fy = py ? mc_filt_l[py - 1] : NULL;
        int ox = (mv[r].x >> 4) - (org[r].x >> 4), oy = (mv[r].y >> 4) - (org[r].y >> 4), sh1 = bitdepth - 8;
        win[r].before = 5; win[r].after = 5;                 /* reads beyond +-3 / w+3 see the replicated border */
        DmvrWin *d = &win[r];
#define DW(x_, y_) rpx(&d->r, d->bx + clipi((x_), -3, d->w + 3), d->by + clipi((y_), -3, d->h + 3))
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int v, X = x + ox, Y = y + oy;
                if (!fx && !fy) v = DW(X, Y) << (14 - bitdepth);
                else if (fx && !fy) { int s = 0; for (int k = 0; k < 8; k++) s += fx[k] * DW(X + k - 3, Y); v = s >> sh1; }
                else if (!fx && fy) { int s = 0; for (int k = 0; k < 8; k++) s += fy[k] * DW(X, Y + k - 3); v = s >> sh1; }
                else {
                    int s2 = 0;
                    for (int j = 0; j < 8; j++) {
                        int s = 0;
                        for (int k = 0; k < 8; k++) s += fx[k] * DW(X + k - 3, Y + j - 3);
                        s2 += fy[j] * (int16_t)(s >> sh1);
                    }
                    v = s2 >> 6;
                }
                P[r][(y + 1) * PS + x + 1] = (int16_t)v;
            }
        if (!disable_bdof) {                                  /* extend_bdof_buff on the shifted window */
            int ex = (px >> 3) != 0, ey = (py >> 3) != 0, sh = 14 - bitdepth;
            for (int i = 0; i < w + 2; i++) {
                P[r][i] = (int16_t)(DW(ox - 1 + i + ex, oy - 1 + ey) << sh);
                P[r][(h + 1) * PS + i] = (int16_t)(DW(ox - 1 + i + ex, oy + h + ey) << sh);
            }
            for (int j = 0; j < h; j++) {
                P[r][(j + 1) * PS] = (int16_t)(DW(ox - 1 + ex, oy + j + ey) << sh);
                P[r][(j + 1) * PS + w + 1] = (int16_t)(DW(ox + w + ex, oy + j + ey) << sh);
            }
        }
#undef DW
    }
    uint16_t *dy = dst[0] + (size_t)c->y0 * dst_stride[0] + c->x0;
    if (disable_bdof) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int o = (y + 1) * PS + x + 1;
                dy[(size_t)y * dst_stride[0] + x] = (uint16_t)clipi((P[0][o] + P[1][o] + (1 << (14 - bitdepth))) >> (15 - bitdepth), 0, maxv);
            }
    } else bdof_core(dy, dst_stride[0], P[0], P[1], w, h, bitdepth);
    /* chroma: window at the clipped ORIGINAL MV, fraction and integer shift from the refined MV (:1079-1124) */
    int wc = w >> 1, hc = h >> 1;
    for (int comp = 1; comp < 3; comp++) {
        int16_t Pc[2][8 * 8];
        for (int r = 0; r < 2; r++) {
            MV mc = clip_mv(c->x0, c->y0, pic_w, pic_h, w, h, org[r]);
            RefPlane rp = { refs->plane[r][ridx[r]][comp], refs->stride[r][ridx[r]][comp], pic_w >> 1, pic_h >> 1 };
            int bx = (c->x0 >> 1) + (mc.x >> 5), by = (c->y0 >> 1) + (mc.y >> 5);
            int px = mv[r].x & 31, py = mv[r].y & 31, sh1 = bitdepth - 8;
            int ox = (mv[r].x >> 5) - (org[r].x >> 5), oy = (mv[r].y >> 5) - (org[r].y >> 5);
            const int8_t *fx = px ? mc_filt_c[px - 1] : NULL, *fy = py ? mc_filt_c[py - 1] : NULL;
#define DC(x_, y_) rpx(&rp, bx + clipi((x_), -1, wc + 1), by + clipi((y_), -1, hc + 1))
            for (int y = 0; y < hc; y++)
                for (int x = 0; x < wc; x++) {
                    int v, X = x + ox, Y = y + oy;
                    if (!fx && !fy) v = DC(X, Y) << (14 - bitdepth);
                    else if (fx && !fy) { int s = 0; for (int k = 0; k < 4; k++) s += fx[k] * DC(X + k - 1, Y); v = s >> sh1; }
                    else if (!fx && fy) { int s = 0; for (int k = 0; k < 4; k++) s += fy[k] * DC(X, Y + k - 1); v = s >> sh1; }
                    else {
                        int s2 = 0;
                        for (int j = 0; j < 4; j++) {
                            int s = 0;
                            for (int k = 0; k < 4; k++) s += fx[k] * DC(X + k - 1, Y + j - 1);
                            s2 += fy[j] * (int16_t)(s >> sh1);
                        }
                        v = s2 >> 6;
                    }
                    Pc[r][y * 8 + x] = (int16_t)v;
                }
#undef DC
        }
        uint16_t *dc = dst[comp] + (size_t)(c->y0 >> 1) * dst_stride[comp] + (c->x0 >> 1);
        for (int y = 0; y < hc; y++)
            for (int x = 0; x < wc; x++)
                dc[(size_t)y * dst_stride[comp] + x] = (uint16_t)clipi((Pc[0][y * 8 + x] + Pc[1][y * 8 + x] + (1 << (14 - bitdepth))) >> (15 - bitdepth), 0, maxv);
    }
}

static void prof_block(uint16_t *d, int stride, const OVOMcRefs *refs, const OVOMcCmd *c, const OVOProfInfo *pi,
                       int pic_w, int pic_h, int bitdepth)
{
    int maxv = (1 << bitdepth) - 1, prof_dir = c->aux & 3, hpel = (c->flags & OVO_MC_HPEL_AMVR) != 0;
    int16_t val[2][16];
    for (int l = 0; l < 2; l++) {
        if (!(c->inter_dir & (1 << l))) continue;
        int16_t P[PS * PS], gx[PS * PS], gy[PS * PS];
        MV mv = l ? (MV){ c->mv1_x, c->mv1_y } : (MV){ c->mv0_x, c->mv0_y };
        padded_pred14(P, refs, l, l ? c->ref_idx1 : c->ref_idx0, pic_w, pic_h, c->x0, c->y0, 4, 4, mv, hpel, bitdepth);
        pad_grad(gx, gy, P, 4, 4);
        const int16_t *dh = l ? pi->dmv_scale_h_1 : pi->dmv_scale_h_0, *dv = l ? pi->dmv_scale_v_1 : pi->dmv_scale_v_0;
        int apply = c->inter_dir == 3 ? (prof_dir >> l) & 1 : 1;     /* uni-prediction always refines (rcn_prof_mcp_l) */
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) {
                int o = (y + 1) * PS + x + 1, v = P[o];
                if (apply) v = (int16_t)(v + clipi(dh[4 * y + x] * gx[o] + dv[4 * y + x] * gy[o], -8192, 8191));   /* rcn_prof, :223-282 */
                val[l][4 * y + x] = (int16_t)v;
            }
    }
    int def = (c->bcw_idx_plus1 == 0 || c->bcw_idx_plus1 == 3), w1 = def ? 0 : bcw_w[c->bcw_idx_plus1 - 1];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            int v, k = 4 * y + x;
            if (c->inter_dir == 3) {
                if (def) v = (val[0][k] + val[1][k] + (1 << (14 - bitdepth))) >> (15 - bitdepth);      /* tmp_prof_mrg, :101-123 */
                else     v = (val[1][k] * w1 + val[0][k] * (8 - w1) + (1 << (16 - bitdepth))) >> (17 - bitdepth);
            } else {
                v = (int16_t)((val[c->inter_dir == 2][k] + (1 << (13 - bitdepth))) >> (14 - bitdepth));
            }
            d[(size_t)y * stride + x] = (uint16_t)clipi(v, 0, maxv);
        }
}

/* ---------------------------------------------------------------------------------------------------
 * GPM (rcn_gpm_b, rcn_inter.c:3118-3146 -> rcn_mc_rpr_b_l / _c with a gpm_ctx, :2525-2592, 2653-2731): two 14-bit
 * uni-predictions (rcn_mcp_bidir0_l / _c) blended with per-sample weights 0..8 taken from the pre-computed masks of
 * rcn_init_gpm_params (libovvc/rcn_gpm.c:150-204): six base masks of 112 x 112, a start offset per (split, size) and a
 * traversal that mirrors the mask for the other angles (rcn_gpm_weights_and_steps, rcn_inter.c:3080-3115); chroma samples
 * the luma mask with step 2 (put_weighted_gpm_bi_pixels, rcn_mc.c:1631-1655).
 * ------------------------------------------------------------------------------------------------- */
static const int8_t gpm_angle2mask[32] = { 0, -1, 1, 2, 3, 4, -1, -1, 5, -1, -1, 4, 3, 2, 1, -1, 0, -1, 1, 2, 3, 4, -1, -1, 5, -1, -1, 4, 3, 2, 1, -1 };
static const int8_t gpm_angle2mirror[32] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2 };
static const int8_t gpm_dis[32] = { 8, 8, 8, 8, 4, 4, 2, 1, 0, -1, -2, -4, -4, -8, -8, -8, -8, -8, -8, -8, -4, -4, -2, -1, 0, 1, 2, 4, 4, 8, 8, 8 };
static const uint8_t gpm_angle[64] = { 0, 0, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 8, 8, 11, 11, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13,
                                       14, 14, 14, 14, 16, 16, 18, 18, 18, 19, 19, 19, 20, 20, 20, 21, 21, 21, 24, 24, 27, 27, 27, 28, 28, 28, 29, 29, 29, 30, 30, 30 };
static const uint8_t gpm_dist[64] = { 1, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 1, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3,
                                      0, 1, 2, 3, 1, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3 };

/* weight of the FIRST prediction at luma position (x, y) of a (1 << lw) x (1 << lh) CU */
static int gpm_weight(int split_dir, int lw, int lh, int x, int y)
{
    int angle = gpm_angle[split_dir], distance = gpm_dist[split_dir], w = 1 << lw, h = 1 << lh;
    int ox = (112 - w) >> 1, oy = (112 - h) >> 1;                              /* g_weightOffset, rcn_gpm.c:182-201 */
    if (distance > 0) {
        if (angle % 16 == 8 || (angle % 16 != 0 && h >= w)) oy += angle < 16 ? (distance * h) >> 3 : -((distance * h) >> 3);
        else                                                ox += angle < 16 ? (distance * w) >> 3 : -((distance * w) >> 3);
    }
    int mirror = gpm_angle2mirror[angle];
    int mx = mirror == 1 ? 111 - ox - x : ox + x, my = mirror == 2 ? 111 - oy - y : oy + y;   /* rcn_inter.c:3100-3114 */
    static const int8_t base_angle[6] = { 0, 2, 3, 4, 5, 8 };                  /* the angle each stored mask was built for */
    int a0 = base_angle[gpm_angle2mask[angle]], dx = a0, dy = (a0 + 8) % 32;
    int rho = (int16_t)((gpm_dis[dx] << 7) + (gpm_dis[dy] << 7));             /* rcn_gpm.c:160-161 */
    int idx = (int16_t)((((mx + 8) << 1) + 1) * gpm_dis[dx] + (int16_t)((((my + 8) << 1) + 1) * gpm_dis[dy]) - rho);
    return clipi((32 + idx + 4) >> 3, 0, 8);
}

static void gpm_block(uint16_t *const dst[3], const int dst_stride[3], const OVOMcRefs *refs, const OVOMcCmd *c,
                      int pic_w, int pic_h, int bitdepth, int16_t *p0, int16_t *p1)
{
    int split = c->aux & 63, l0 = (c->aux >> 8) & 1, l1 = (c->aux >> 9) & 1, maxv = (1 << bitdepth) - 1;
    MV mv0 = { c->mv0_x, c->mv0_y }, mv1 = { c->mv1_x, c->mv1_y };
    for (int comp = 0; comp < 3; comp++) {
        int sh = comp ? 1 : 0, w = (1 << c->log2_w) >> sh, h = (1 << c->log2_h) >> sh;
        uint16_t *d = dst[comp] + (size_t)(c->y0 >> sh) * dst_stride[comp] + (c->x0 >> sh);
        pu_pred14(p0, refs, l0, c->ref_idx0, comp, pic_w, pic_h, c, mv0, bitdepth);
        pu_pred14(p1, refs, l1, c->ref_idx1, comp, pic_w, pic_h, c, mv1, bitdepth);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int w0 = gpm_weight(split, c->log2_w, c->log2_h, x << sh, y << sh);
                d[(size_t)y * dst_stride[comp] + x] = (uint16_t)clipi((p1[y * 128 + x] * (8 - w0) + p0[y * 128 + x] * w0 + (1 << (16 - bitdepth))) >> (17 - bitdepth), 0, maxv);
            }
    }
}

void ovo_mc_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                    const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut)
{
    ovo_mc_picture2(dst, dst_stride, pic_w, pic_h, bitdepth, refs, cmds, nb_cmds, lmcs_fwd_lut, NULL);
}

void ovo_mc_picture2(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                     const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut,
                     const OVOProfInfo *prof_tab)
{
    ovo_mc_picture3(dst, dst_stride, pic_w, pic_h, bitdepth, refs, cmds, nb_cmds, lmcs_fwd_lut, prof_tab, NULL);
}

void ovo_mc_picture3(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                     const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut,
                     const OVOProfInfo *prof_tab, OVODmvrMv *dmvr_out)
{
    int nb_dmvr = 0;
    int16_t *p0 = malloc(2 * 128 * 128 * sizeof(int16_t)), *p1 = p0 + 128 * 128;   /* per call: thread-safe */
    int maxv = (1 << bitdepth) - 1;
    for (int i = 0; i < nb_cmds; i++) {
        const OVOMcCmd *c = &cmds[i];
        if (c->tool == OVO_MC_TOOL_BDOF || c->tool == OVO_MC_TOOL_PROF || c->tool == OVO_MC_TOOL_DMVR) {
            uint16_t *d = dst[0] + (size_t)c->y0 * dst_stride[0] + c->x0;
            if (c->tool == OVO_MC_TOOL_BDOF) bdof_block(d, dst_stride[0], refs, c, pic_w, pic_h, bitdepth);
            else if (c->tool == OVO_MC_TOOL_DMVR) { dmvr_block(dst, dst_stride, refs, c, pic_w, pic_h, bitdepth, dmvr_out ? &dmvr_out[nb_dmvr] : NULL); nb_dmvr++; }
            else prof_block(d, dst_stride[0], refs, c, &prof_tab[c->aux >> 8], pic_w, pic_h, bitdepth);
            if (lmcs_fwd_lut)
                for (int y = 0; y < (1 << c->log2_h); y++)
                    for (int x = 0; x < (1 << c->log2_w); x++) {
                        uint16_t *q = &d[(size_t)y * dst_stride[0] + x];
                        *q = lmcs_fwd_lut[*q & maxv];
                    }
            continue;
        }
        if (c->tool == OVO_MC_TOOL_GPM) {
            gpm_block(dst, dst_stride, refs, c, pic_w, pic_h, bitdepth, p0, p1);
            if (lmcs_fwd_lut)                                             /* rcn_inter.c:2591 */
                for (int y = 0; y < (1 << c->log2_h); y++)
                    for (int x = 0; x < (1 << c->log2_w); x++) {
                        uint16_t *q = dst[0] + (size_t)(c->y0 + y) * dst_stride[0] + c->x0 + x;
                        *q = lmcs_fwd_lut[*q & maxv];
                    }
            continue;
        }
        MV mv0 = { c->mv0_x, c->mv0_y }, mv1 = { c->mv1_x, c->mv1_y };
        /* check_identical_motion, rcn_inter.c:256-267 (POCs compared as uint16_t) */
        int identical = c->inter_dir == 3 &&
                        (uint16_t)refs->poc[0][c->ref_idx0] == (uint16_t)refs->poc[1][c->ref_idx1] &&
                        mv0.x == mv1.x && mv0.y == mv1.y;
        int bi = c->inter_dir == 3 && !identical;
        int uni_list = (c->inter_dir & 2 || identical) ? 1 : 0;          /* rcn_inter.c:2795-2812 */
        for (int comp = 0; comp < 3; comp++) {
            if (comp == 0 && !(c->flags & OVO_MC_LUMA)) continue;
            if (comp != 0 && !(c->flags & OVO_MC_CHROMA)) continue;
            int sh = comp ? 1 : 0;
            int w = (1 << c->log2_w) >> sh, h = (1 << c->log2_h) >> sh;
            uint16_t *d = dst[comp] + (size_t)(c->y0 >> sh) * dst_stride[comp] + (c->x0 >> sh);
            if (bi) {
                pu_pred14(p0, refs, 0, c->ref_idx0, comp, pic_w, pic_h, c, mv0, bitdepth);
                pu_pred14(p1, refs, 1, c->ref_idx1, comp, pic_w, pic_h, c, mv1, bitdepth);
                int def = (c->bcw_idx_plus1 == 0 || c->bcw_idx_plus1 == 3);   /* rcn_inter.c:587 */
                int w1 = def ? 0 : bcw_w[c->bcw_idx_plus1 - 1], w0 = 8 - w1;
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        int a = p0[y * 128 + x], b = p1[y * 128 + x], v;
                        if (def) v = (a + b + (1 << (14 - bitdepth))) >> (15 - bitdepth);   /* rcn_mc.c:433-439 */
                        else     v = (b * w1 + a * w0 + (1 << (16 - bitdepth))) >> (17 - bitdepth);   /* rcn_mc.c:1520-1526 */
                        d[(size_t)y * dst_stride[comp] + x] = (uint16_t)clipi(v, 0, maxv);
                    }
            } else {
                MV mv = uni_list ? mv1 : mv0;
                int ref_idx = uni_list ? c->ref_idx1 : c->ref_idx0;
                pu_pred14(p0, refs, uni_list, ref_idx, comp, pic_w, pic_h, c, mv, bitdepth);
                int shift = 14 - bitdepth, off = 1 << (shift - 1);       /* rcn_mc.c:458-465 */
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++)
                        d[(size_t)y * dst_stride[comp] + x] = (uint16_t)clipi((p0[y * 128 + x] + off) >> shift, 0, maxv);
            }
            if (comp == 0 && lmcs_fwd_lut)                                /* rcn_inter.c:599, 1552 */
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        uint16_t *q = &d[(size_t)y * dst_stride[0] + x];
                        *q = lmcs_fwd_lut[*q & maxv];
                    }
        }
    }
    free(p0);
}

/* ---------------------------------------------------------------------------------------------------
 * Reference picture resampling: rcn_mcp_rpr_l / _bi_l / _c / _bi_c (rcn_inter.c:2025-2513), their combination
 * rcn_mc_rpr_b_l / _b_c (:2524-2731, also the GPM path :3118-3146), rcn_mc_rpr_prof_b_l (:2594-2649), the dispatch on
 * the slice's scale factors in rcn_mcp / rcn_mcp_b / _b_l / _b_c / rcn_prof_mcp_b_l (:2749-2966) and the filter
 * kernels put_vvc_*_rpr_* (rcn_mc.c:535-785).
 *
 * A scaled reference is sampled on a grid of its own: output column c reads reference position
 * (ref_pos_x + c * step) in 1/16 (chroma 1/32) sample units scaled by 2^14, one separable pass per direction with the
 * 14-bit intermediate kept as int16.  The reference position is clipped so that the window stays within four samples
 * of the reference picture (clip_rpr_position), the window border is replicated (emulate_block_border).
 * ------------------------------------------------------------------------------------------------- */
#include "mc_rpr_tables.h"

static int rpr_filter_idx(int scale, int flag_4x4)
{   /* compute_rpr_filter_idx, rcn_inter.c:1990-2006: thresholds 1.25 and 1.75 in 2^14 units */
    int idx = flag_4x4 ? 3 : 0;
    if (scale > (1 << 14) * 7 / 4) idx += 2;
    else if (scale > (1 << 14) * 5 / 4) idx += 1;
    return idx;
}

/* One component of one PU from one SCALED reference.  uni != 0: final clipped samples (rpr_v_uni); else the 14-bit
 * prediction as int16 (rpr_v_bi).  Output stride 128.  pos / pu sizes in samples of the component. */
static void rpr_pred(int16_t *out, const RefPlane *r, int chroma, int pos_x, int pos_y, int pu_w, int pu_h, int flag_4x4,
                     MV mv, int scale_h, int scale_v, int hor_col, int ver_col, int bitdepth, int uni)
{
    const int shift_mv = chroma ? 5 : 4, shift_pos = 14 + shift_mv, offset = 1 << 13, fmask = (1 << shift_mv) - 1;
    const int nt = chroma ? 4 : 8, before = chroma ? 1 : 3;
    int fih = rpr_filter_idx(scale_h, flag_4x4), fiv = rpr_filter_idx(scale_v, flag_4x4);
    /* chroma (rcn_inter.c:2300-2301) looks the same index up in a table of three sets: the 4x4 flag would read past it,
     * a 4x4 luma block has 2x2 chroma blocks that no conforming stream predicts on their own */
    if (chroma) { fih %= 3; fiv %= 3; }
    const int stepX = ((scale_h + 8) >> 4) << 4, stepY = ((scale_v + 8) >> 4) << 4;
    /* 32-bit products as the reference computes them (they wrap beyond 2^31 for 8K positions at 2x) */
    int32_t ref_pos_x = (int32_t)((uint32_t)((pos_x << shift_mv) + mv.x) * (uint32_t)scale_h);
    int32_t ref_pos_y = (int32_t)((uint32_t)((pos_y << shift_mv) + mv.y) * (uint32_t)scale_v);
    if (chroma) {
        ref_pos_x += (1 - hor_col) * 8 * (scale_h - (1 << 14)) + (1 << 8);
        ref_pos_y += (1 - ver_col) * 8 * (scale_v - (1 << 14)) + (1 << 8);
    } else { ref_pos_x += 1 << 7; ref_pos_y += 1 << 7; }
    int ref_x = (ref_pos_x + offset) >> shift_pos, ref_y = (ref_pos_y + offset) >> shift_pos;
    int ref_pu_w = ((ref_pos_x + (((pu_w - 1) * stepX) << shift_mv) + offset) >> shift_pos) - ref_x + 1;
    int ref_pu_h = ((ref_pos_y + (((pu_h - 1) * stepY) << shift_mv) + offset) >> shift_pos) - ref_y + 1;
    if (ref_pu_h < 1) ref_pu_h = 1;
    {   /* clip_rpr_position, rcn_inter.c:2009-2022 */
        int prec_x = ref_pos_x & ((1 << shift_pos) - 1), prec_y = ref_pos_y & ((1 << shift_pos) - 1);
        int x_max = (r->w + 3) << shift_pos, y_max = (r->h + 3) << shift_pos;
        int x_min = -((ref_pu_w + 4) << shift_pos), y_min = -((ref_pu_h + 4) << shift_pos);
        ref_pos_x = clipi(ref_pos_x, x_min + prec_x, x_max + prec_x);
        ref_pos_y = clipi(ref_pos_y, y_min + prec_y, y_max + prec_y);
    }
    ref_x = (ref_pos_x + offset) >> shift_pos; ref_y = (ref_pos_y + offset) >> shift_pos;
    const int rows = ref_pu_h + nt;                       /* the rows the vertical pass can reach: ref_y - before .. */
    int16_t *tmp = malloc((size_t)rows * pu_w * sizeof(int16_t));
    for (int col = 0; col < pu_w; col++) {
        int pos_mv_x = (ref_pos_x + ((col * stepX) << shift_mv) + offset) >> 14;
        int prec_x = pos_mv_x & fmask, xi = pos_mv_x >> shift_mv;
        const int8_t *f = chroma ? mc_rpr_c[fih][prec_x] : mc_rpr_l[fih][prec_x];
        for (int rr = 0; rr < rows; rr++) {
            int y = ref_y - before + rr, v;
            if (!fih && !prec_x) v = rpx(r, xi, y) << (14 - bitdepth);                      /* put_vvc_pel_rpr */
            else {
                int sacc = 0;
                for (int k = 0; k < nt; k++) sacc += f[k] * rpx(r, xi + k - before, y);
                v = sacc >> (bitdepth - 8);                                                  /* put_vvc_{q,e}pel_rpr_h */
            }
            tmp[rr * pu_w + col] = (int16_t)v;
        }
    }
    const int shift = 14 - bitdepth, off = 1 << (shift - 1), maxv = (1 << bitdepth) - 1;
    for (int row = 0; row < pu_h; row++) {
        int pos_mv_y = (ref_pos_y + ((row * stepY) << shift_mv) + offset) >> 14;
        int prec_y = pos_mv_y & fmask, rb = (pos_mv_y >> shift_mv) - ref_y;
        const int8_t *f = chroma ? mc_rpr_c[fiv][prec_y] : mc_rpr_l[fiv][prec_y];
        for (int x = 0; x < pu_w; x++) {
            int v;
            if (!fiv && !prec_y) {
                /* put_vvc_pel_rpr_clip reads the intermediate as UNSIGNED 16 bit (rcn_mc.c:623-642): a negative one saturates high */
                if (uni) v = clipi(((uint16_t)tmp[(rb + before) * pu_w + x] + off) >> shift, 0, maxv);
                else     v = tmp[(rb + before) * pu_w + x];                                  /* put_vvc_pel_rpr_bi */
            } else {
                int sacc = 0;
                for (int k = 0; k < nt; k++) sacc += f[k] * tmp[(rb + k) * pu_w + x];
                if (uni) v = clipi(((sacc >> 6) + off) >> shift, 0, maxv);                   /* put_vvc_*_rpr_clip_v */
                else     v = (int16_t)(sacc >> 6);                                           /* put_vvc_*_rpr_bi_v */
            }
            out[row * 128 + x] = (int16_t)v;
        }
    }
    free(tmp);
}

static int rpr_scaled(const OVOMcScaling *sc, int list, int idx) { return sc->scale[list][idx][0] != (1 << 14) || sc->scale[list][idx][1] != (1 << 14); }

/* 14-bit prediction (or, uni, final samples) of one component from either kind of reference */
static void rpr_any_pred(int16_t *out, const OVOMcRefs *refs, const OVOMcScaling *sc, int list, int ref_idx, int comp,
                         int pic_w, int pic_h, const OVOMcCmd *c, MV mv, int bitdepth, int uni)
{
    if (!rpr_scaled(sc, list, ref_idx)) {                 /* rcn_mcp_bidir0_l / _c; uni: the caller rounds */
        pu_pred14(out, refs, list, ref_idx, comp, pic_w, pic_h, c, mv, bitdepth);
        return;
    }
    const int sh = comp ? 1 : 0;
    RefPlane r = { refs->plane[list][ref_idx][comp], refs->stride[list][ref_idx][comp], sc->ref_w[list][ref_idx] >> sh, sc->ref_h[list][ref_idx] >> sh };
    rpr_pred(out, &r, comp != 0, c->x0 >> sh, c->y0 >> sh, (1 << c->log2_w) >> sh, (1 << c->log2_h) >> sh,
             c->log2_w == 2 && c->log2_h == 2, mv, sc->scale[list][ref_idx][0], sc->scale[list][ref_idx][1],
             sc->chroma_hor_col, sc->chroma_ver_col, bitdepth, uni);
}

void ovo_mc_picture_rpr(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                        const OVOMcRefs *refs, const OVOMcScaling *sc, const OVOMcCmd *cmds, int nb_cmds,
                        const uint16_t *lmcs_fwd_lut, const OVOProfInfo *prof_tab)
{
    int16_t *p0 = malloc(2 * 128 * 128 * sizeof(int16_t)), *p1 = p0 + 128 * 128;
    const int maxv = (1 << bitdepth) - 1;
    for (int i = 0; i < nb_cmds; i++) {
        const OVOMcCmd *c = &cmds[i];
        const int gpm = c->tool == OVO_MC_TOOL_GPM;
        MV mv0 = { c->mv0_x, c->mv0_y }, mv1 = { c->mv1_x, c->mv1_y };
        int l0 = 0, l1 = 1;
        if (gpm) { l0 = (c->aux >> 8) & 1; l1 = (c->aux >> 9) & 1; }
        const int identical = c->tool == OVO_MC_TOOL_PLAIN && c->inter_dir == 3 &&
                              (uint16_t)refs->poc[0][c->ref_idx0] == (uint16_t)refs->poc[1][c->ref_idx1] && mv0.x == mv1.x && mv0.y == mv1.y;
        const int bi = gpm || (c->inter_dir == 3 && !identical);
        const int uni_list = (c->inter_dir & 2 || identical) ? 1 : 0;
        const int s0 = (bi || !uni_list) && rpr_scaled(sc, l0, c->ref_idx0), s1 = (bi || uni_list) && rpr_scaled(sc, l1, c->ref_idx1);
        if ((c->tool != OVO_MC_TOOL_PLAIN && c->tool != OVO_MC_TOOL_GPM && c->tool != OVO_MC_TOOL_PROF) || (!s0 && !s1)) {
            ovo_mc_picture3(dst, dst_stride, pic_w, pic_h, bitdepth, refs, c, 1, lmcs_fwd_lut, prof_tab, NULL);
            continue;
        }
        for (int comp = 0; comp < 3; comp++) {
            if (comp == 0 && !(c->flags & OVO_MC_LUMA)) continue;
            if (comp != 0 && !(c->flags & OVO_MC_CHROMA)) continue;
            const int sh = comp ? 1 : 0, w = (1 << c->log2_w) >> sh, h = (1 << c->log2_h) >> sh;
            uint16_t *d = dst[comp] + (size_t)(c->y0 >> sh) * dst_stride[comp] + (c->x0 >> sh);
            if (!bi) {                                    /* rcn_mcp_rpr_l / _c; a PROF command of a scaled reference is plain (:2895-2917) */
                rpr_any_pred(p0, refs, sc, uni_list, uni_list ? c->ref_idx1 : c->ref_idx0, comp, pic_w, pic_h, c, uni_list ? mv1 : mv0, bitdepth, 1);
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) d[(size_t)y * dst_stride[comp] + x] = (uint16_t)p0[y * 128 + x];
            } else {
                const int prof_dir = c->tool == OVO_MC_TOOL_PROF ? (int)(c->aux & 3) : 0;
                for (int l = 0; l < 2; l++) {
                    int16_t *pb = l ? p1 : p0;
                    const int list = l ? l1 : l0, scaled = l ? s1 : s0, ridx = l ? c->ref_idx1 : c->ref_idx0;
                    const MV mv = l ? mv1 : mv0;
                    if (!scaled && ((prof_dir >> l) & 1)) {
                        /* rcn_mc_rpr_prof_b_l (apply_prof, :2880): the one unscaled list of a mixed bi-prediction is refined by
                         * rcn_prof_mcp_bi_l when its prof_dir bit is set; the scaled list never is */
                        const OVOProfInfo *pi = &prof_tab[c->aux >> 8];
                        int16_t P[PS * PS], gx[PS * PS], gy[PS * PS];
                        padded_pred14(P, refs, list, ridx, pic_w, pic_h, c->x0, c->y0, 4, 4, mv, (c->flags & OVO_MC_HPEL_AMVR) != 0, bitdepth);
                        pad_grad(gx, gy, P, 4, 4);
                        const int16_t *dh = l ? pi->dmv_scale_h_1 : pi->dmv_scale_h_0, *dv = l ? pi->dmv_scale_v_1 : pi->dmv_scale_v_0;
                        for (int y = 0; y < 4; y++)
                            for (int x = 0; x < 4; x++) {
                                int o = (y + 1) * PS + x + 1;
                                pb[y * 128 + x] = (int16_t)(P[o] + clipi(dh[4 * y + x] * gx[o] + dv[4 * y + x] * gy[o], -8192, 8191));
                            }
                        continue;
                    }
                    rpr_any_pred(pb, refs, sc, list, ridx, comp, pic_w, pic_h, c, mv, bitdepth, 0);
                }
                const int def = gpm || c->bcw_idx_plus1 == 0 || c->bcw_idx_plus1 == 3;
                const int w1 = def ? 0 : bcw_w[c->bcw_idx_plus1 - 1], w0 = 8 - w1;
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        int a = p0[y * 128 + x], b = p1[y * 128 + x], v;
                        if (gpm) {                        /* put_weighted_gpm_bi_pixels */
                            int g0 = gpm_weight(c->aux & 63, c->log2_w, c->log2_h, x << sh, y << sh);
                            v = (b * (8 - g0) + a * g0 + (1 << (16 - bitdepth))) >> (17 - bitdepth);
                        } else if (def) v = (a + b + (1 << (14 - bitdepth))) >> (15 - bitdepth);              /* put_vvc_qpel_rpr_bi_sum */
                        else            v = (a * w0 + b * w1 + (1 << (16 - bitdepth))) >> (17 - bitdepth);   /* put_vvc_qpel_rpr_weighted, denom 3 */
                        d[(size_t)y * dst_stride[comp] + x] = (uint16_t)clipi(v, 0, maxv);
                    }
            }
            if (comp == 0 && lmcs_fwd_lut)
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        uint16_t *q = &d[(size_t)y * dst_stride[0] + x];
                        *q = lmcs_fwd_lut[*q & maxv];
                    }
        }
    }
    free(p0);
}

/* ---------------------------------------------------------------------------------------------------
 * LMCS look-up tables (rcn_lmcs.c:84-187, 189-203)
 * ------------------------------------------------------------------------------------------------- */
void ovo_lmcs_build_luts(OVOLmcsLuts *luts, const OVOLmcsData *data, int bitdepth)
{
    const int nb_wnd = 16, prec = 11, rnd = 1 << (prec - 1);
    int rng = 1 << bitdepth, log2_wnd = bitdepth - 4, smp_wnd = rng >> 4;
    int min_idx = data->lmcs_min_bin_idx, max_idx_p1 = nb_wnd - data->lmcs_delta_max_bin_idx;
    int cw_delta[16] = { 0 };
    uint16_t fwd_step[16] = { 0 }, bwd_step[16] = { 0 }, bnd[17] = { 0 };
    int maxv = rng - 1;

    for (int i = min_idx; i < max_idx_p1; i++)
        cw_delta[i] = data->lmcs_delta_sign_cw_flag[i] ? -data->lmcs_delta_abs_cw[i] : data->lmcs_delta_abs_cw[i];
    for (int i = min_idx; i < max_idx_p1; i++) {          /* compute_windows_scale_steps */
        int wnd_sz = smp_wnd + (int16_t)cw_delta[i];
        if (wnd_sz) {
            fwd_step[i] = (uint16_t)(((wnd_sz << prec) + (1 << (log2_wnd - 1))) >> log2_wnd);
            bwd_step[i] = (uint16_t)((smp_wnd << prec) / wnd_sz);
        }
        bnd[i + 1] = (uint16_t)(bnd[i] + wnd_sz);
    }
    for (int i = max_idx_p1; i < nb_wnd; i++) bnd[i + 1] = bnd[i];

    for (int val = 0; val < rng; val++) {                 /* derive_forward_lut */
        int idx = val >> log2_wnd;
        int nb_step = val - (idx << log2_wnd);
        int v = (int16_t)bnd[idx] + (((int32_t)fwd_step[idx] * nb_step + rnd) >> prec);
        luts->fwd_lut[val] = (uint16_t)clipi(v, 0, maxv);
    }
    for (int val = 0; val < rng; val++) {                 /* derive_backward_lut + get_bwd_idx */
        int i = min_idx;
        for (; i < max_idx_p1; i++)
            if (val < bnd[i + 1]) break;
        int idx = i < nb_wnd - 1 ? i : nb_wnd - 1;
        int nb_step = val - bnd[idx];
        int v = (idx << log2_wnd) + (((int32_t)bwd_step[idx] * nb_step + rnd) >> prec);
        luts->bwd_lut[val] = (uint16_t)clipi(v, 0, maxv);
    }
    memcpy(luts->wnd_bnd, bnd, sizeof(bnd));
}

void ovo_lmcs_reshape_plane(uint16_t *plane, int stride, int w, int h, const uint16_t *lut, int bitdepth)
{   /* rcn_lmcs_reshape_luma_blk_lut, rcn_lmcs.c:274-284 */
    int msk = (1 << bitdepth) - 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            plane[(size_t)y * stride + x] = lut[plane[(size_t)y * stride + x] & msk];
}

/* rcn_lmcs_compute_chroma_scale (libovvc/rcn_lmcs.c:204-343): src points at the top-left luma sample of the 64x64 region.
 * The reference walks each mask until it is empty (so the run length is the position of the highest set bit), pads the
 * run to 16 units with its last sample, and averages over 64 or 128 samples. */
int ovo_lmcs_chroma_scale(const uint16_t *src, int stride, unsigned abv_mask, unsigned lft_mask,
                          const OVOLmcsLuts *luts, const OVOLmcsData *d, int bitdepth)
{
    unsigned sum = 0;
    int n_abv = 0, n_lft = 0, nb_units, log2_nb = 0, avg, i;
    const uint16_t *p = src - stride;
    while (abv_mask) { sum += p[0] + p[1] + p[2] + p[3]; p += 4; n_abv++; abv_mask >>= 1; }
    if (n_abv) { sum += 4u * p[-1] * (16 - n_abv); n_abv = 16; }
    p = src - 1;
    while (lft_mask) { sum += p[0] + p[stride] + p[2 * stride] + p[3 * stride]; p += 4 * stride; n_lft++; lft_mask >>= 1; }
    if (n_lft) { sum += 4u * p[-stride] * (16 - n_lft); n_lft = 16; }
    nb_units = n_abv + n_lft;
    while (nb_units) { log2_nb++; nb_units >>= 1; }
    avg = log2_nb ? (int)((sum + (1u << log2_nb)) >> (log2_nb + 1)) : 1 << (bitdepth - 1);
    int min_idx = d->lmcs_min_bin_idx, max_idx_p1 = 16 - d->lmcs_delta_max_bin_idx;
    for (i = min_idx; i < max_idx_p1; i++) if (avg < luts->wnd_bnd[i + 1]) break;      /* get_bwd_idx, :84-94 */
    if (i > 15) i = 15;
    int wnd_sz = (int)luts->wnd_bnd[i + 1] - (int)luts->wnd_bnd[i];
    int offset = d->lmcs_delta_sign_crs_flag ? -(int)d->lmcs_delta_abs_crs : (int)d->lmcs_delta_abs_crs;
    if (wnd_sz == 0 || wnd_sz + offset == 0) return 1 << 11;                           /* (the reference would divide by zero) */
    return (int16_t)(uint16_t)((1 << (bitdepth - 4 + 11)) / (wnd_sz + offset));         /* uint16_t field read as int16_t */
}
