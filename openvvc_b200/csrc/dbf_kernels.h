// dbf_kernels.h -- deblocking filter of a whole picture (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: rcn_dbf_ctu / rcn_dbf_truncated_ctu (libovvc/rcn_df.c:2170-2228) called per CTU right after
// the CTU is written to the frame; vvc_dbf_ctu_hor/_ver (:1941-2013, :2088-2167) + vvc_dbf_chroma_hor/_ver
// (:1149-1430); filter-length derivation (:1876-1939); decisions filter_vertical_edge / filter_horizontal_edge
// (:1434-1510, :2016-2085); limits (:172-193); filters (:217-1085).
//
// The reference filters CTU by CTU (vertical edges, then horizontal edges lagging two units) which equals "all
// vertical edges of the picture, then all horizontal edges" (SURVEY.md Appendix B).  Here each pass is ONE launch
// over the picture-level maps of OVGPUDbfMaps: one thread per 4-sample luma edge unit (2-sample chroma unit, grid.z =
// component), lanes along x so that the rows a warp touches are contiguous.  Edge units of one direction read and
// write disjoint samples (filter lengths are bounded by the distance to the next edge), so the pass runs in place.
#pragma once
#include "dev.h"

enum { DBF_EDGE_TU = 0, DBF_EDGE_SB, DBF_BS1, DBF_BS2, DBF_AFFINE, DBF_EDGE_C, DBF_BS1_CB, DBF_BS1_CR, DBF_BS2_C, DBF_NB_MAPS };

OVD_CONST uint16_t dbf_tc_lut[67] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      3,   4,   4,   4,   4,   5,   5,   5,   5,   7,   7,   8,   9,  10,  10,  11,
     13,  14,  15,  17,  19,  21,  24,  25,  29,  33,  36,  41,  45,  51,  57,  64,
     71,  80,  89, 100, 112, 125, 141, 157, 177, 198, 222, 250, 280, 314, 352, 395, 0 };   // entry 66 is 0 in the reference
OVD_CONST uint8_t dbf_beta_lut[65] = {
     0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9, 10, 11, 12, 13,
    14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56,
    58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88, 0 };
OVD_CONST int8_t dbf_db7[7] = { 59, 50, 41, 32, 23, 14, 5 };
OVD_CONST int8_t dbf_db5[5] = { 58, 45, 32, 19, 6 };
OVD_CONST int8_t dbf_db3[3] = { 53, 32, 11 };
OVD_CONST int8_t dbf_tcw7[7] = { 6, 5, 4, 3, 2, 1, 1 };
OVD_CONST int8_t dbf_tcw3[3] = { 6, 4, 2 };

struct DbfMapsDev {
    const uint32_t *ver[DBF_NB_MAPS];     // [cy][X], X in 0..w_u
    const uint32_t *hor[DBF_NB_MAPS];     // [cx][Y], Y in 0..h_u
    const uint8_t *qp[3];                 // [Y][X]
    int n_u, w_u, h_u;
    int beta_offset, tc_offset;
};

template <int HOR>
struct DbfK {
#ifndef DBF_MIN_BLOCKS
#define DBF_MIN_BLOCKS 6                   /* measured 3 / 4 / 5 / 6 CTAs per SM: vertical edges 38.5 / 34.1 / 30.3 / 29.2 us per 4K picture */
#endif
    static const int THREADS = 256, MIN_BLOCKS = DBF_MIN_BLOCKS, PHASES = 1;
    static const int BX = 32, BY = 8;     // lanes along x (units); 8 unit rows / edge rows per CTA

    struct Args {
        uint16_t *plane[3];
        int pitch[3];                     // samples
        int bitdepth;
        DbfMapsDev m;
    };
    struct State { int dummy; };

    // bit of map k at edge index e (column for vertical edges, row for horizontal ones), unit s along the edge
    OVD int eb(const DbfMapsDev &m, int k, int e, int s)
    {
        if (HOR) {
            if (e < 0 || e > m.h_u || s < 0 || s >= m.w_u) return 0;
            return (m.hor[k][(s / m.n_u) * (m.h_u + 1) + e] >> (s % m.n_u)) & 1;
        }
        if (e < 0 || e > m.w_u || s < 0 || s >= m.h_u) return 0;
        return (m.ver[k][(s / m.n_u) * (m.w_u + 1) + e] >> (s % m.n_u)) & 1;
    }

    OVD void limits(const Args &a, int qp, int bs, int &beta, int &tc)
    {
        const int tc_idx = ovd_clamp(qp + 2 * (bs - 1) + a.m.tc_offset, 0, 66), beta_idx = ovd_clamp(qp + a.m.beta_offset, 0, 64);
        tc = a.bitdepth >= 10 ? dbf_tc_lut[tc_idx] << (a.bitdepth - 10) : (dbf_tc_lut[tc_idx] + (1 << (9 - a.bitdepth))) >> (10 - a.bitdepth);
        beta = dbf_beta_lut[beta_idx] << (a.bitdepth - 8);
    }

    OVD int iabs(int v) { return v < 0 ? -v : v; }
    // samples across the edge: s points at q0, `ac` = step across the edge
    OVD int P(const uint16_t *s, ptrdiff_t ac, int i) { return s[-(ptrdiff_t)(i + 1) * ac]; }
    OVD int Q(const uint16_t *s, ptrdiff_t ac, int i) { return s[(ptrdiff_t)i * ac]; }
    OVD int dp_of(const uint16_t *s, ptrdiff_t ac) { return iabs(P(s, ac, 2) - 2 * P(s, ac, 1) + P(s, ac, 0)); }
    OVD int dq_of(const uint16_t *s, ptrdiff_t ac) { return iabs(Q(s, ac, 0) - 2 * Q(s, ac, 1) + Q(s, ac, 2)); }

    OVD int strong_l0(const uint16_t *s, ptrdiff_t ac, int beta, int tc, int lp, int lq)
    {
        int sp3 = iabs(P(s, ac, 3) - P(s, ac, 0)), sq3 = iabs(Q(s, ac, 3) - Q(s, ac, 0));
        if (lp == 7)      { sp3 += iabs(P(s, ac, 4) - P(s, ac, 5) - P(s, ac, 6) + P(s, ac, 7)); sp3 += iabs(P(s, ac, 3) - P(s, ac, 7)) + 1; sp3 >>= 1; }
        else if (lp == 5) { sp3 += iabs(P(s, ac, 3) - P(s, ac, 5)) + 1; sp3 >>= 1; }
        if (lq == 7)      { sq3 += iabs(Q(s, ac, 4) - Q(s, ac, 5) - Q(s, ac, 6) + Q(s, ac, 7)); sq3 += iabs(Q(s, ac, 7) - Q(s, ac, 3)) + 1; sq3 >>= 1; }
        else if (lq == 5) { sq3 += iabs(Q(s, ac, 5) - Q(s, ac, 3)) + 1; sq3 >>= 1; }
        return (sp3 + sq3) < (beta * 3 >> 5) && iabs(P(s, ac, 0) - Q(s, ac, 0)) < ((tc * 5 + 1) >> 1);
    }
    OVD int strong_l1(const uint16_t *s, ptrdiff_t ac, int beta, int tc)
    {
        return (iabs(P(s, ac, 3) - P(s, ac, 0)) + iabs(Q(s, ac, 3) - Q(s, ac, 0))) < (beta >> 3) &&
               iabs(P(s, ac, 0) - Q(s, ac, 0)) < ((tc * 5 + 1) >> 1);
    }

    OVD void long_line(uint16_t *s, ptrdiff_t ac, int tc, int lp, int lq)
    {
        int p[8], q[8], mid;
#pragma unroll
        for (int i = 0; i < 8; i++) { p[i] = i <= lp ? P(s, ac, i) : 0; q[i] = i <= lq ? Q(s, ac, i) : 0; }
        if (lp == 7 && lq == 7)      mid = (2 * (p[0] + q[0]) + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
        else if (lp + lq == 12)      mid = (2 * (p[0] + p[1] + q[0] + q[1]) + p[2] + p[3] + p[4] + p[5] + q[2] + q[3] + q[4] + q[5] + 8) >> 4;
        else if (lp == 5 && lq == 5) mid = (2 * (p[0] + p[1] + p[2] + q[0] + q[1] + q[2]) + p[3] + p[4] + q[3] + q[4] + 8) >> 4;
        else if (lp == 7 && lq == 3) mid = (2 * (p[0] + q[0]) + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + q[0] + 3 * q[1] + 2 * q[2] + 8) >> 4;
        else if (lp == 3 && lq == 7) mid = (2 * (p[0] + q[0]) + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + p[0] + 3 * p[1] + 2 * p[2] + 8) >> 4;
        else                         mid = (p[0] + p[1] + p[2] + p[3] + q[0] + q[1] + q[2] + q[3] + 4) >> 3;
        int ref_p = 0, ref_q = 0;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            if (i == lp) ref_p = (p[i - 1] + p[i] + 1) >> 1;
            if (i == lq) ref_q = (q[i - 1] + q[i] + 1) >> 1;
        }
#pragma unroll
        for (int i = 0; i < 7; i++) {
            if (i < lp) {
                const int db = lp == 7 ? dbf_db7[i] : (lp == 5 ? dbf_db5[i < 5 ? i : 0] : dbf_db3[i < 3 ? i : 0]);
                const int cv = (tc * (lp == 3 ? dbf_tcw3[i < 3 ? i : 0] : dbf_tcw7[i])) >> 1;
                s[-(ptrdiff_t)(i + 1) * ac] = (uint16_t)ovd_clamp((mid * db + ref_p * (64 - db) + 32) >> 6, p[i] - cv, p[i] + cv);
            }
            if (i < lq) {
                const int db = lq == 7 ? dbf_db7[i] : (lq == 5 ? dbf_db5[i < 5 ? i : 0] : dbf_db3[i < 3 ? i : 0]);
                const int cv = (tc * (lq == 3 ? dbf_tcw3[i < 3 ? i : 0] : dbf_tcw7[i])) >> 1;
                s[(ptrdiff_t)i * ac] = (uint16_t)ovd_clamp((mid * db + ref_q * (64 - db) + 32) >> 6, q[i] - cv, q[i] + cv);
            }
        }
    }

    OVD void strong_small_line(uint16_t *s, ptrdiff_t ac, int tc)
    {
        const int p0 = P(s, ac, 0), p1 = P(s, ac, 1), p2 = P(s, ac, 2), p3 = P(s, ac, 3);
        const int q0 = Q(s, ac, 0), q1 = Q(s, ac, 1), q2 = Q(s, ac, 2), q3 = Q(s, ac, 3);
        s[-3 * ac] = (uint16_t)ovd_clamp((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
        s[-2 * ac] = (uint16_t)ovd_clamp((p2 + p1 + p0 + q0 + 2) >> 2, p1 - 2 * tc, p1 + 2 * tc);
        s[-1 * ac] = (uint16_t)ovd_clamp((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, p0 - 3 * tc, p0 + 3 * tc);
        s[0]       = (uint16_t)ovd_clamp((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, q0 - 3 * tc, q0 + 3 * tc);
        s[ac]      = (uint16_t)ovd_clamp((p0 + q0 + q1 + q2 + 2) >> 2, q1 - 2 * tc, q1 + 2 * tc);
        s[2 * ac]  = (uint16_t)ovd_clamp((p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3, q2 - tc, q2 + tc);
    }

    OVD void weak_line(uint16_t *s, ptrdiff_t ac, int tc, int ext_p, int ext_q, int maxv)
    {
        const int p0 = P(s, ac, 0), p1 = P(s, ac, 1), p2 = P(s, ac, 2), q0 = Q(s, ac, 0), q1 = Q(s, ac, 1), q2 = Q(s, ac, 2);
        int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
        if (iabs(delta) < tc * 10) {
            delta = ovd_clamp(delta, -tc, tc);
            s[-1 * ac] = (uint16_t)ovd_clamp(p0 + delta, 0, maxv);
            s[0]       = (uint16_t)ovd_clamp(q0 - delta, 0, maxv);
            // the reference also rewrites p1 / q1 with a zero correction when the side is not extended; skipping the
            // store is equivalent and keeps the write set minimal
            if (ext_p) s[-2 * ac] = (uint16_t)ovd_clamp(p1 + ovd_clamp((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -(tc >> 1), tc >> 1), 0, maxv);
            if (ext_q) s[ac]      = (uint16_t)ovd_clamp(q1 + ovd_clamp((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -(tc >> 1), tc >> 1), 0, maxv);
        }
    }

    OVD void luma_unit(const Args &a, int e, int u)
    {
        const DbfMapsDev &m = a.m;
        const int tu = eb(m, DBF_EDGE_TU, e, u), sb = eb(m, DBF_EDGE_SB, e, u);
        const int bs2 = eb(m, DBF_BS2, e, u);
        if (!((tu | sb) & (eb(m, DBF_BS1, e, u) | bs2))) return;
        const int X = HOR ? u : e, Y = HOR ? e : u;
        const int qp_p = HOR ? m.qp[0][(Y - 1) * m.w_u + X] : m.qp[0][Y * m.w_u + X - 1];
        const int qp = (qp_p + m.qp[0][Y * m.w_u + X] + 1) >> 1;
        int beta, tc;
        limits(a, qp, 1 + bs2, beta, tc);
        if (!(tc | beta)) return;
        // set_edge_context + derive_filter_length (rcn_df.c:1876-1939)
        const int i = e % m.n_u;
        int large_p = 0, large_q = 0;
        if ((i & 3) == 0) {
            int any_p = 0, any_q = 0;
#pragma unroll
            for (int k = 1; k <= 7; k++) { any_p |= eb(m, DBF_EDGE_TU, e - k, u); any_q |= eb(m, DBF_EDGE_TU, e + k, u); }
            large_p = !any_p; large_q = !any_q;
            if (HOR && i < 7) large_p = 0;                    // :2144
        }
        if (sb && !tu) large_p = large_q = 0;
        const int small = eb(m, DBF_EDGE_TU, e - 1, u) | eb(m, DBF_EDGE_TU, e + 1, u) | eb(m, DBF_EDGE_SB, e - 1, u) | eb(m, DBF_EDGE_SB, e + 1, u);
        const int aff1 = sb && (eb(m, DBF_EDGE_TU, e - 2, u) | eb(m, DBF_EDGE_TU, e + 2, u)) && !tu;
        int lp, lq;
        if (small) lp = lq = 1;
        else if (aff1) lp = lq = 2;
        else {
            lp = large_p ? (eb(m, DBF_AFFINE, e - 1, u) ? 5 : 7) : 3;
            lq = large_q ? (eb(m, DBF_AFFINE, e, u) ? 5 : 7) : 3;
        }
        // filter_vertical_edge / filter_horizontal_edge (:1434-1510, :2016-2085)
        const ptrdiff_t ac = HOR ? a.pitch[0] : 1, al = HOR ? 1 : a.pitch[0];
        uint16_t *s0 = a.plane[0] + (size_t)(Y * 4) * a.pitch[0] + X * 4, *s3 = s0 + 3 * al;
        const int dp0 = dp_of(s0, ac), dq0 = dq_of(s0, ac), dp3 = dp_of(s3, ac), dq3 = dq_of(s3, ac);
        const int d0 = dp0 + dq0, d3 = dp3 + dq3, maxv = (1 << a.bitdepth) - 1;
        if (d0 + d3 >= beta) return;
        int strong_large = 0;
        if (lp > 3 || lq > 3) {
            int dp0L = dp0, dq0L = dq0, dp3L = dp3, dq3L = dq3;
            if (lp > 3) { dp0L = (dp0L + dp_of(s0 - 3 * ac, ac) + 1) >> 1; dp3L = (dp3L + dp_of(s3 - 3 * ac, ac) + 1) >> 1; }
            if (lq > 3) { dq0L = (dq0L + dq_of(s0 + 3 * ac, ac) + 1) >> 1; dq3L = (dq3L + dq_of(s3 + 3 * ac, ac) + 1) >> 1; }
            const int d0L = dp0L + dq0L, d3L = dp3L + dq3L;
            strong_large = (d0L + d3L < beta) && d0L < ((beta + 0x10) >> 5) && d3L < ((beta + 0x10) >> 5) &&
                           strong_l0(s0, ac, beta, tc, lp, lq) && strong_l0(s3, ac, beta, tc, lp, lq);
        }
        if (strong_large) {
            for (int k = 0; k < 4; k++) long_line(s0 + k * al, ac, tc, lp, lq);
            return;
        }
        const int sw = lp > 2 && d0 < ((beta + 4) >> 3) && d3 < ((beta + 4) >> 3) && strong_l1(s0, ac, beta, tc) && strong_l1(s3, ac, beta, tc);
        if (sw) {
            for (int k = 0; k < 4; k++) strong_small_line(s0 + k * al, ac, tc);
        } else {
            const int side = (beta + (beta >> 1)) >> 3;
            const int ext_p = (dp0 + dp3) < side && lp > 1, ext_q = (dq0 + dq3) < side && lp > 1;      // sic (rcn_df.c:1505, 2080)
            for (int k = 0; k < 4; k++) weak_line(s0 + k * al, ac, tc, ext_p, ext_q, maxv);
        }
    }

    OVD void chroma_unit(const Args &a, int comp, int e, int u)
    {
        const DbfMapsDev &m = a.m;
        const int bs2 = eb(m, DBF_BS2_C, e, u), bs1 = eb(m, comp == 1 ? DBF_BS1_CB : DBF_BS1_CR, e, u);
        if (!eb(m, DBF_EDGE_C, e, u) || !(bs2 | bs1)) return;
        int any = 0;                                              // derive_large_map_from_ngh, :1087-1105
#pragma unroll
        for (int k = 1; k <= 3; k++) any |= eb(m, DBF_EDGE_C, e - k, u) | eb(m, DBF_EDGE_C, e + k, u);
        const int large = !any;
        if (!(bs2 | (bs1 & large))) return;
        const int X = HOR ? u : e, Y = HOR ? e : u;
        const uint8_t *qpm = m.qp[comp];
        const int qp_p = HOR ? qpm[(Y - 1) * m.w_u + X] : qpm[Y * m.w_u + X - 1];
        const int qp = (qp_p + qpm[Y * m.w_u + X] + 1) >> 1, maxv = (1 << a.bitdepth) - 1;
        int beta, tc;
        limits(a, qp, 1 + bs2, beta, tc);
        if (tc == 0 || beta == 0) return;
        const ptrdiff_t ac = HOR ? a.pitch[comp] : 1, al = HOR ? 1 : a.pitch[comp];
        uint16_t *s0 = a.plane[comp] + (size_t)(Y * 2) * a.pitch[comp] + X * 2, *s1 = s0 + al;
        const int ctb_b = HOR && (e % m.n_u == 0);                // horizontal edge on a CTU boundary: 2-line P side
        int strong = 0;
        if (large) {
            const int dp0 = iabs((ctb_b ? P(s0, ac, 1) : P(s0, ac, 2)) - 2 * P(s0, ac, 1) + P(s0, ac, 0));
            const int dp3 = iabs((ctb_b ? P(s1, ac, 1) : P(s1, ac, 2)) - 2 * P(s1, ac, 1) + P(s1, ac, 0));
            const int d0 = dp0 + dq_of(s0, ac), d3 = dp3 + dq_of(s1, ac);
            strong = (d0 + d3 < beta) && 2 * d0 < (beta >> 2) && 2 * d3 < (beta >> 2);
            for (int k = 0; k < 2 && strong; k++) {
                const uint16_t *s = k ? s1 : s0;
                const int p3 = ctb_b ? P(s, ac, 1) : P(s, ac, 3);
                strong = (iabs(p3 - P(s, ac, 0)) + iabs(Q(s, ac, 3) - Q(s, ac, 0))) < (beta >> 3) && iabs(P(s, ac, 0) - Q(s, ac, 0)) < ((tc * 5 + 1) >> 1);
            }
        }
        for (int k = 0; k < 2; k++) {
            uint16_t *s = k ? s1 : s0;
            const int p0 = P(s, ac, 0), p1 = P(s, ac, 1), q0 = Q(s, ac, 0), q1 = Q(s, ac, 1);
            if (strong) {
                const int p2 = P(s, ac, 2), p3 = P(s, ac, 3), q2 = Q(s, ac, 2), q3 = Q(s, ac, 3);
                if (ctb_b) {
                    s[-1 * ac] = (uint16_t)ovd_clamp((3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
                    s[0]       = (uint16_t)ovd_clamp((2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
                } else {
                    s[-3 * ac] = (uint16_t)ovd_clamp((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
                    s[-2 * ac] = (uint16_t)ovd_clamp((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3, p1 - tc, p1 + tc);
                    s[-1 * ac] = (uint16_t)ovd_clamp((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
                    s[0]       = (uint16_t)ovd_clamp((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
                }
                s[ac]      = (uint16_t)ovd_clamp((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
                s[2 * ac]  = (uint16_t)ovd_clamp((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
            } else {
                const int delta = ovd_clamp(((q0 << 2) - (p0 << 2) + p1 - q1 + 4) >> 3, -tc, tc);
                s[-1 * ac] = (uint16_t)ovd_clamp(p0 + delta, 0, maxv);
                s[0]       = (uint16_t)ovd_clamp(q0 - delta, 0, maxv);
            }
        }
    }

    template <int PH>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st;
        const int x = bx * BX + (tid & 31), y = by * BY + (tid >> 5);
        // vertical edges: x = edge column (1..w_u-1), y = unit row; horizontal: x = unit column, y = edge row (1..h_u-1)
        if (bz == 0) {
            if (HOR) { if (x < a.m.w_u && y >= 1 && y < a.m.h_u) luma_unit(a, y, x); }
            else     { if (x >= 1 && x < a.m.w_u && y < a.m.h_u) luma_unit(a, x, y); }
        } else {
            // chroma edges lie on the 16-luma-sample grid: every 4th unit
            if (HOR) { if (x < a.m.w_u && 4 * y >= 1 && 4 * y < a.m.h_u) chroma_unit(a, bz, 4 * y, x); }
            else     { if (4 * x >= 1 && 4 * x < a.m.w_u && y < a.m.h_u) chroma_unit(a, bz, 4 * x, y); }
        }
    }
};
