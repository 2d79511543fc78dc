/* oracle/port_df.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of OpenVVC's deblocking filter (libovvc/rcn_df.c) on picture-level maps (OVODbfMaps): rcn_dbf_ctu
 * (:2170-2198) / rcn_dbf_truncated_ctu (:2200-2228), vvc_dbf_ctu_hor / _ver (:1941-2013, :2088-2167), filter-length
 * derivation (:1876-1939), edge decisions filter_vertical_edge / filter_horizontal_edge (:1434-1510, :2016-2085),
 * compute_dbf_limits (:172-193), the long / strong / weak luma filters (:217-958) and the chroma path (:961-1430).
 * CTUs are walked in raster order with the reference's schedule (vertical edges of the CTU, then its horizontal
 * edges shifted two units to the left), so the result is the reference's even where edges could interact.
 */
#include "oracle.h"
#include <stdlib.h>

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

static const uint16_t tc_lut[67] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      3,   4,   4,   4,   4,   5,   5,   5,   5,   7,   7,   8,   9,  10,  10,  11,
     13,  14,  15,  17,  19,  21,  24,  25,  29,  33,  36,  41,  45,  51,  57,  64,
     71,  80,  89, 100, 112, 125, 141, 157, 177, 198, 222, 250, 280, 314, 352, 395, 0 /* sic: 67th entry is implicit */ };
static const uint8_t beta_lut[65] = {
     0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9, 10, 11, 12, 13,
    14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56,
    58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88, 0 /* sic */ };

/* which filters ran (test coverage only): 0 long, 1 strong-short, 2 weak, 3 chroma strong, 4 chroma weak, 5 + lp/2*4+lq/2 */
static long dbf_stats[32];
void ovo_dbf_stats(long *out, int reset) { for (int i = 0; i < 32; i++) { out[i] = dbf_stats[i]; if (reset) dbf_stats[i] = 0; } }

typedef struct {
    const OVODbfMaps *m;
    int n_u, w_u, h_u, bitdepth;
} Ctx;

/* bit of map k at vertical edge column X, unit row Y / horizontal edge row Y, unit column X (0 outside) */
static inline int vb(const Ctx *c, int k, int X, int Y)
{
    if (X < 0 || X > c->w_u || Y < 0 || Y >= c->h_u) return 0;
    return (c->m->ver[k][(Y / c->n_u) * (c->w_u + 1) + X] >> (Y % c->n_u)) & 1;
}
static inline int hb(const Ctx *c, int k, int X, int Y)
{
    if (Y < 0 || Y > c->h_u || X < 0 || X >= c->w_u) return 0;
    return (c->m->hor[k][(X / c->n_u) * (c->h_u + 1) + Y] >> (X % c->n_u)) & 1;
}
/* direction-agnostic: e = index of the edge along the filtering axis, s = index of the unit along the edge */
static inline int eb(const Ctx *c, int hor, int k, int e, int s) { return hor ? hb(c, k, s, e) : vb(c, k, e, s); }

static void limits(const Ctx *c, int qp, int bs, int *beta, int *tc)
{   /* compute_dbf_limits, rcn_df.c:172-193 */
    int tc_idx = clipi(qp + 2 * (bs - 1) + c->m->tc_offset, 0, 66), beta_idx = clipi(qp + c->m->beta_offset, 0, 64);
    *tc = c->bitdepth >= 10 ? tc_lut[tc_idx] << (c->bitdepth - 10) : (tc_lut[tc_idx] + (1 << (9 - c->bitdepth))) >> (10 - c->bitdepth);
    *beta = beta_lut[beta_idx] << (c->bitdepth - 8);
}

/* sample access across the edge: P(i) = i-th sample on the P side (0 nearest), Q(i) likewise; `a` = step across */
#define P(i) ((int)s[-(ptrdiff_t)((i) + 1) * a])
#define Q(i) ((int)s[(ptrdiff_t)(i) * a])

static int dp_of(const uint16_t *s, ptrdiff_t a) { return iabs(P(2) - 2 * P(1) + P(0)); }
static int dq_of(const uint16_t *s, ptrdiff_t a) { return iabs(Q(0) - 2 * Q(1) + Q(2)); }

static int strong_l0(const uint16_t *s, ptrdiff_t a, int beta, int tc, int lp, int lq)
{   /* use_strong_filter_l0, :81-123 */
    int sp3 = iabs(P(3) - P(0)), sq3 = iabs(Q(3) - Q(0));
    if (lp == 7)      { sp3 += iabs(P(4) - P(5) - P(6) + P(7)); sp3 += iabs(P(3) - P(7)) + 1; sp3 >>= 1; }
    else if (lp == 5) { sp3 += iabs(P(3) - P(5)) + 1; sp3 >>= 1; }
    if (lq == 7)      { sq3 += iabs(Q(4) - Q(5) - Q(6) + Q(7)); sq3 += iabs(Q(7) - Q(3)) + 1; sq3 >>= 1; }
    else if (lq == 5) { sq3 += iabs(Q(5) - Q(3)) + 1; sq3 >>= 1; }
    return (sp3 + sq3) < (beta * 3 >> 5) && iabs(P(0) - Q(0)) < ((tc * 5 + 1) >> 1);
}
static int strong_l1(const uint16_t *s, ptrdiff_t a, int beta, int tc)
{   /* use_strong_filter_l1, :125-139 */
    return (iabs(P(3) - P(0)) + iabs(Q(3) - Q(0))) < (beta >> 3) && iabs(P(0) - Q(0)) < ((tc * 5 + 1) >> 1);
}

static const int8_t db7[7] = { 59, 50, 41, 32, 23, 14, 5 }, db5[5] = { 58, 45, 32, 19, 6 }, db3[3] = { 53, 32, 11 };
static const int8_t tcw7[7] = { 6, 5, 4, 3, 2, 1, 1 }, tcw3[3] = { 6, 4, 2 };

/* long filters filter_{h,v}_P_Q, :217-848: one line */
static void long_line(uint16_t *s, ptrdiff_t a, int tc, int lp, int lq)
{
    int p[8], q[8], mid;
    for (int i = 0; i < 8; i++) { p[i] = P(i); q[i] = Q(i); }
    if (lp == 7 && lq == 7)      mid = (2 * (p[0] + q[0]) + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
    else if (lp + lq == 12)      mid = (2 * (p[0] + p[1] + q[0] + q[1]) + p[2] + p[3] + p[4] + p[5] + q[2] + q[3] + q[4] + q[5] + 8) >> 4;
    else if (lp == 5 && lq == 5) mid = (2 * (p[0] + p[1] + p[2] + q[0] + q[1] + q[2]) + p[3] + p[4] + q[3] + q[4] + 8) >> 4;
    else if (lp == 7 && lq == 3) mid = (2 * (p[0] + q[0]) + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + q[0] + 3 * q[1] + 2 * q[2] + 8) >> 4;
    else if (lp == 3 && lq == 7) mid = (2 * (p[0] + q[0]) + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + p[0] + 3 * p[1] + 2 * p[2] + 8) >> 4;
    else                         mid = (p[0] + p[1] + p[2] + p[3] + q[0] + q[1] + q[2] + q[3] + 4) >> 3;      /* 5/3 and 3/5 */
    int ref_p = (p[lp - 1] + p[lp] + 1) >> 1, ref_q = (q[lq - 1] + q[lq] + 1) >> 1;
    const int8_t *dbp = lp == 7 ? db7 : (lp == 5 ? db5 : db3), *dbq = lq == 7 ? db7 : (lq == 5 ? db5 : db3);
    const int8_t *twp = lp == 3 ? tcw3 : tcw7, *twq = lq == 3 ? tcw3 : tcw7;
    for (int i = 0; i < lp; i++) {
        int cv = (tc * twp[i]) >> 1;
        s[-(ptrdiff_t)(i + 1) * a] = (uint16_t)clipi((mid * dbp[i] + ref_p * (64 - dbp[i]) + 32) >> 6, p[i] - cv, p[i] + cv);
    }
    for (int i = 0; i < lq; i++) {
        int cv = (tc * twq[i]) >> 1;
        s[(ptrdiff_t)i * a] = (uint16_t)clipi((mid * dbq[i] + ref_q * (64 - dbq[i]) + 32) >> 6, q[i] - cv, q[i] + cv);
    }
}

static void strong_small_line(uint16_t *s, ptrdiff_t a, int tc)
{   /* filter_luma_strong_small_{h,v}, :849-900 */
    int p0 = P(0), p1 = P(1), p2 = P(2), p3 = P(3), q0 = Q(0), q1 = Q(1), q2 = Q(2), q3 = Q(3);
    s[-3 * a] = (uint16_t)clipi((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
    s[-2 * a] = (uint16_t)clipi((p2 + p1 + p0 + q0 + 2) >> 2, p1 - 2 * tc, p1 + 2 * tc);
    s[-1 * a] = (uint16_t)clipi((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, p0 - 3 * tc, p0 + 3 * tc);
    s[0]      = (uint16_t)clipi((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, q0 - 3 * tc, q0 + 3 * tc);
    s[a]      = (uint16_t)clipi((p0 + q0 + q1 + q2 + 2) >> 2, q1 - 2 * tc, q1 + 2 * tc);
    s[2 * a]  = (uint16_t)clipi((p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3, q2 - tc, q2 + tc);
}

static void weak_line(uint16_t *s, ptrdiff_t a, int tc, int ext_p, int ext_q, int maxv)
{   /* filter_luma_weak_{h,v}, :902-958 */
    int p0 = P(0), p1 = P(1), p2 = P(2), q0 = Q(0), q1 = Q(1), q2 = Q(2);
    int tc2_p = ext_p ? tc >> 1 : 0, tc2_q = ext_q ? tc >> 1 : 0;
    int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
    if (iabs(delta) < tc * 10) {
        delta = clipi(delta, -tc, tc);
        int d1 = clipi((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -tc2_p, tc2_p);
        int d2 = clipi((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -tc2_q, tc2_q);
        s[-2 * a] = (uint16_t)clipi(p1 + d1, 0, maxv);
        s[-1 * a] = (uint16_t)clipi(p0 + delta, 0, maxv);
        s[0]      = (uint16_t)clipi(q0 - delta, 0, maxv);
        s[a]      = (uint16_t)clipi(q1 + d2, 0, maxv);
    }
}

/* filter_vertical_edge / filter_horizontal_edge, :1434-1510 / :2016-2085: one 4-line segment.
 * a = step across the edge, l = step along it. */
static void luma_segment(const Ctx *c, uint16_t *s0, ptrdiff_t a, ptrdiff_t l, int beta, int tc, int lp, int lq)
{
    const uint16_t *s = s0;
    uint16_t *s3 = s0 + 3 * l;
    int dp0 = dp_of(s0, a), dq0 = dq_of(s0, a), dp3 = dp_of(s3, a), dq3 = dq_of(s3, a);
    int d0 = dp0 + dq0, d3 = dp3 + dq3, maxv = (1 << c->bitdepth) - 1;
    (void)s;
    if (d0 + d3 >= beta) return;
    int strong_large = 0;
    if (lp > 3 || lq > 3) {
        int dp0L = dp0, dq0L = dq0, dp3L = dp3, dq3L = dq3;
        if (lp > 3) { dp0L = (dp0L + dp_of(s0 - 3 * a, a) + 1) >> 1; dp3L = (dp3L + dp_of(s3 - 3 * a, a) + 1) >> 1; }
        if (lq > 3) { dq0L = (dq0L + dq_of(s0 + 3 * a, a) + 1) >> 1; dq3L = (dq3L + dq_of(s3 + 3 * a, a) + 1) >> 1; }
        int d0L = dp0L + dq0L, d3L = dp3L + dq3L;
        strong_large = (d0L + d3L < beta) && d0L < ((beta + 0x10) >> 5) && d3L < ((beta + 0x10) >> 5) &&
                       strong_l0(s0, a, beta, tc, lp, lq) && strong_l0(s3, a, beta, tc, lp, lq);
    }
    if (strong_large) {
        for (int i = 0; i < 4; i++) long_line(s0 + i * l, a, tc, lp, lq);
        dbf_stats[0]++; dbf_stats[5 + (lp >> 1) * 4 + (lq >> 1)]++;
        return;
    }
    int sw = lp > 2 && d0 < ((beta + 4) >> 3) && d3 < ((beta + 4) >> 3) && strong_l1(s0, a, beta, tc) && strong_l1(s3, a, beta, tc);
    if (sw) {
        for (int i = 0; i < 4; i++) strong_small_line(s0 + i * l, a, tc);
        dbf_stats[1]++;
    } else {
        dbf_stats[2]++;
        int side = (beta + (beta >> 1)) >> 3;
        int ext_p = (dp0 + dp3) < side && lp > 1, ext_q = (dq0 + dq3) < side && lp > 1;      /* sic: max_l_p for both */
        for (int i = 0; i < 4; i++) weak_line(s0 + i * l, a, tc, ext_p, ext_q, maxv);
    }
}

/* one luma unit edge: hor = 0 vertical edge at column e, unit row u; hor = 1 horizontal edge at row e, unit column u */
static void luma_unit(const Ctx *c, uint16_t *plane, int stride, int hor, int e, int u)
{
    if (!((eb(c, hor, OVO_DBF_EDGE_TU, e, u) | eb(c, hor, OVO_DBF_EDGE_SB, e, u)) &
          (eb(c, hor, OVO_DBF_BS1, e, u) | eb(c, hor, OVO_DBF_BS2, e, u)))) return;
    int bs = 1 + eb(c, hor, OVO_DBF_BS2, e, u);
    int X = hor ? u : e, Y = hor ? e : u;
    int qp_p = hor ? c->m->qp[0][(Y - 1) * c->w_u + X] : c->m->qp[0][Y * c->w_u + X - 1];
    int qp = (qp_p + c->m->qp[0][Y * c->w_u + X] + 1) >> 1, beta, tc;
    limits(c, qp, bs, &beta, &tc);
    if (!(tc || beta)) return;
    /* set_edge_context + derive_filter_length, :1876-1939 */
    int i = e % c->n_u;
    int tu = eb(c, hor, OVO_DBF_EDGE_TU, e, u), sb = eb(c, hor, OVO_DBF_EDGE_SB, e, u);
    int large_p = 0, large_q = 0;
    if (i % 4 == 0) {
        int any_p = 0, any_q = 0;
        for (int k = 1; k <= 7; k++) { any_p |= eb(c, hor, OVO_DBF_EDGE_TU, e - k, u); any_q |= eb(c, hor, OVO_DBF_EDGE_TU, e + k, u); }
        large_p = !any_p; large_q = !any_q;
        if (hor && i < 7) large_p = 0;                         /* :2144: no long P side across the CTU row above */
    }
    if (sb && !tu) large_p = large_q = 0;
    int small = eb(c, hor, OVO_DBF_EDGE_TU, e - 1, u) | eb(c, hor, OVO_DBF_EDGE_TU, e + 1, u) |
                eb(c, hor, OVO_DBF_EDGE_SB, e - 1, u) | eb(c, hor, OVO_DBF_EDGE_SB, e + 1, u);
    int aff1 = sb && (eb(c, hor, OVO_DBF_EDGE_TU, e - 2, u) | eb(c, hor, OVO_DBF_EDGE_TU, e + 2, u)) && !tu;
    /* affine_p / affine_q: the unit on either side belongs to an affine CU (affine_map.ver[i], [i + 1]) */
    int aff_p = eb(c, hor, OVO_DBF_AFFINE, e - 1, u), aff_q = eb(c, hor, OVO_DBF_AFFINE, e, u);
    int lp, lq;
    if (small) lp = lq = 1;
    else if (aff1) lp = lq = 2;
    else { lp = large_p ? (aff_p ? 5 : 7) : 3; lq = large_q ? (aff_q ? 5 : 7) : 3; }
    uint16_t *s = plane + (size_t)(Y * 4) * stride + X * 4;
    luma_segment(c, s, hor ? stride : 1, hor ? 1 : stride, beta, tc, lp, lq);
}

/* chroma: one unit (2 chroma lines) of the edge of plane comp (1 Cb, 2 Cr); :961-1430 */
static void chroma_unit(const Ctx *c, uint16_t *plane, int stride, int comp, int hor, int e, int u)
{
    int bs1k = comp == 1 ? OVO_DBF_BS1_CB : OVO_DBF_BS1_CR;
    int bs2 = eb(c, hor, OVO_DBF_BS2_C, e, u), bs1 = eb(c, hor, bs1k, e, u);
    if (!eb(c, hor, OVO_DBF_EDGE_C, e, u) || !(bs2 | bs1)) return;
    int any = 0;                                                /* derive_large_map_from_ngh, :1087-1105 */
    for (int k = 1; k <= 3; k++) any |= eb(c, hor, OVO_DBF_EDGE_C, e - k, u) | eb(c, hor, OVO_DBF_EDGE_C, e + k, u);
    int large = !any;
    if (!(bs2 | (bs1 & large))) return;
    int X = hor ? u : e, Y = hor ? e : u;
    const uint8_t *qpm = c->m->qp[comp];
    int qp_p = hor ? qpm[(Y - 1) * c->w_u + X] : qpm[Y * c->w_u + X - 1];
    int qp = (qp_p + qpm[Y * c->w_u + X] + 1) >> 1, beta, tc, maxv = (1 << c->bitdepth) - 1;
    limits(c, qp, 1 + bs2, &beta, &tc);
    if (tc == 0 || beta == 0) return;
    ptrdiff_t a = hor ? stride : 1, l = hor ? 1 : stride;
    uint16_t *s0 = plane + (size_t)(Y * 2) * stride + X * 2, *s1 = s0 + l;
    int ctb_b = hor && (e % c->n_u == 0);                       /* horizontal edge on a CTU boundary: 2-line P side */
    int strong = 0;
    if (large) {
        const uint16_t *s;
        int dp0, dp3;
        s = s0; dp0 = iabs((ctb_b ? P(1) : P(2)) - 2 * P(1) + P(0));      /* compute_dp_c, :142-146 */
        s = s1; dp3 = iabs((ctb_b ? P(1) : P(2)) - 2 * P(1) + P(0));
        int dq0 = dq_of(s0, a), dq3 = dq_of(s1, a);
        int d0 = dp0 + dq0, d3 = dp3 + dq3;
        strong = (d0 + d3 < beta) && 2 * d0 < (beta >> 2) && 2 * d3 < (beta >> 2);
        for (int k = 0; k < 2 && strong; k++) {                /* use_strong_filter_c / _c2, :961-995 */
            s = k ? s1 : s0;
            int p3 = ctb_b ? P(1) : P(3);
            strong = (iabs(p3 - P(0)) + iabs(Q(3) - Q(0))) < (beta >> 3) && iabs(P(0) - Q(0)) < ((tc * 5 + 1) >> 1);
        }
    }
    dbf_stats[strong ? 3 : 4]++;
    for (int k = 0; k < 2; k++) {
        uint16_t *s = k ? s1 : s0;
        int p0 = P(0), p1 = P(1), p2 = P(2), p3 = P(3), q0 = Q(0), q1 = Q(1), q2 = Q(2), q3 = Q(3);
        if (strong && ctb_b) {
            s[-1 * a] = (uint16_t)clipi((3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
            s[0]      = (uint16_t)clipi((2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            s[a]      = (uint16_t)clipi((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
            s[2 * a]  = (uint16_t)clipi((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
        } else if (strong) {
            s[-3 * a] = (uint16_t)clipi((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
            s[-2 * a] = (uint16_t)clipi((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3, p1 - tc, p1 + tc);
            s[-1 * a] = (uint16_t)clipi((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
            s[0]      = (uint16_t)clipi((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            s[a]      = (uint16_t)clipi((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
            s[2 * a]  = (uint16_t)clipi((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
        } else {
            int delta = clipi(((q0 << 2) - (p0 << 2) + p1 - q1 + 4) >> 3, -tc, tc);
            s[-1 * a] = (uint16_t)clipi(p0 + delta, 0, maxv);
            s[0]      = (uint16_t)clipi(q0 - delta, 0, maxv);
        }
    }
}

void ovo_dbf_picture(uint16_t *const planes[3], const int stride[3], int bitdepth, const OVODbfMaps *maps)
{
    Ctx c = { maps, (1 << maps->log2_ctu_s) >> 2, maps->width >> 2, maps->height >> 2, bitdepth };
    int n_u = c.n_u, ncw = (c.w_u + n_u - 1) / n_u, nch = (c.h_u + n_u - 1) / n_u;
    for (int cy = 0; cy < nch; cy++)
        for (int cx = 0; cx < ncw; cx++) {
            int nuw = c.w_u - cx * n_u < n_u ? c.w_u - cx * n_u : n_u, nuh = c.h_u - cy * n_u < n_u ? c.h_u - cy * n_u : n_u;
            int last_x = cx == ncw - 1;
            /* horizontal-edge pass of a CTU lags two units behind (vvc_dbf_ctu_ver, :2099-2106) */
            int hx0 = cx * n_u - 2 < 0 ? 0 : cx * n_u - 2, hx1 = cx * n_u + nuw - 2 + (last_x ? 2 : 0);
            /* luma: vertical edges, then horizontal edges */
            for (int i = cx ? 0 : 1; i < nuw; i++)
                for (int y = 0; y < nuh; y++) luma_unit(&c, planes[0], stride[0], 0, cx * n_u + i, cy * n_u + y);
            for (int i = cy ? 0 : 1; i < nuh; i++)
                for (int x = hx0; x < hx1; x++) luma_unit(&c, planes[0], stride[0], 1, cy * n_u + i, x);
            /* chroma: edges on the 16-luma-sample grid */
            for (int comp = 1; comp < 3; comp++)
                for (int i = cx ? 0 : 1; i < (nuw + 3) >> 2; i++)
                    for (int y = 0; y < nuh; y++) chroma_unit(&c, planes[comp], stride[comp], comp, 0, cx * n_u + 4 * i, cy * n_u + y);
            for (int comp = 1; comp < 3; comp++)
                for (int i = cy ? 0 : 1; i < (nuh + 3) >> 2; i++)
                    for (int x = hx0; x < hx1; x++) chroma_unit(&c, planes[comp], stride[comp], comp, 1, cy * n_u + 4 * i, x);
        }
}
