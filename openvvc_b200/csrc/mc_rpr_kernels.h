// mc_rpr_kernels.h -- motion compensation from SCALED references (reference picture resampling; sm_100a, also compiled by
// tests/emu, see dev.h).
//
// Reference behaviour: rcn_mcp_rpr_l / _bi_l / _c / _bi_c (libovvc/rcn_inter.c:2025-2513), their combinations rcn_mc_rpr_b_l /
// _b_c (:2524-2731; the GPM path :3118-3146 goes through them too), rcn_mc_rpr_prof_b_l (:2594-2649), the dispatch on the
// slice's scale factors (:2749-2966) and the filter kernels put_vvc_*_rpr_* (libovvc/rcn_mc.c:535-785).
//
// A command lands here when a reference it names has a scale factor other than 1 (decided by the device-side command walk
// from the batch's scaling table, mc_host.h).  One warp owns one <=16x16 luma block (+ its chroma blocks) of the PU:
//   phases 0 / 2  horizontal pass of list 0 / 1: output column c of a scaled list reads reference position
//                 ref_pos_x + c * step (1/16 sample, x 2^14), one filter per column picked by its own fraction; the rows the
//                 vertical pass will reach go to shared memory as int16 (the reference's tmp_rpr)
//   phases 1 / 3  vertical pass: one filter per output row, 14-bit result per list in shared memory
//   phase 4       the lists are combined (rounding / average / BCW / GPM weights), clipped, forward-LMCS'ed (luma), stored
// The unscaled list of a mixed bi-prediction runs through the same two passes with a step of one sample and the regular filters
// (rcn_mcp_bidir0_l / _c: clip_mv, half-pel AMVR, 6-tap set of 4x4 blocks), refined by PROF in phase 4 when the command asks for
// it (rcn_prof_mcp_bi_l).  Plain scalar integer code: resampled pictures are rare (not part of the random-access test
// conditions) and this kernel is about being right, not about the roofline.
#pragma once
#include "dev.h"
#include "mc_kernels.h"
#include "mc_tools_kernels.h"
#include "mc_rpr_tables.h"

struct McRprK {
    static const int THREADS = 128, MIN_BLOCKS = 4, PHASES = 5, WARPS = 4;
    static const int TMP_Y_ROWS = 40, TMP_C_ROWS = 24;                    // 16 rows at step 2 + 7, 8 rows at step 2 + 3 (+ slack)
    static const int SM_TMP = TMP_Y_ROWS * 16 + 2 * TMP_C_ROWS * 8;      // int16 entries
    static const int SM_EX = 2 * (256 + 2 * 64);                         // [list][luma 16x16, Cb 8x8, Cr 8x8]
    static const int SM_WARP = (SM_TMP + SM_EX) * 2;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        McK::Args mc;
        int ref_w[2][16], ref_h[2][16];       // luma size of every reference
        uint16_t scale[2][16][2];             // inter_ctx->scale_fact_rpl0 / 1 [ref_idx][hor, ver]
        int hor_col, ver_col;                 // sps_chroma_{horizontal,vertical}_collocated_flag
        const ProfInfoDev *prof;
    };
    struct State { int dummy; };

    struct Cmd {                              // the warp's item, resolved
        int valid, x0, y0, log2_w, log2_h, ox, oy, w, h, flags, tool, bi, gpm_split, bcw, prof_dir, prof_idx;
        int list[2], ridx[2], mvx[2], mvy[2], scaled[2];
    };

    OVD Cmd load(const Args &a, int item_idx)
    {
        Cmd c;
        c.valid = item_idx < McK::nb_items(a.mc);
        if (!c.valid) return c;
        const McItemDev im = a.mc.items[item_idx];
        const McCmdDev m = a.mc.cmds[im.cmd];
        c.x0 = m.x0; c.y0 = m.y0; c.log2_w = m.log2_w; c.log2_h = m.log2_h; c.ox = im.ox; c.oy = im.oy; c.w = im.w; c.h = im.h;
        c.flags = m.flags; c.tool = m.tool;
        const int gpm = m.tool == 4;
        const int identical = m.tool == 0 && m.inter_dir == 3 && (uint16_t)a.mc.refs.poc[0][m.ref_idx0] == (uint16_t)a.mc.refs.poc[1][m.ref_idx1] &&
                              m.mv0_x == m.mv1_x && m.mv0_y == m.mv1_y;
        c.bi = gpm || (m.inter_dir == 3 && !identical);
        c.gpm_split = gpm ? (int)(m.aux & 63) : -1;
        c.bcw = (gpm || m.bcw_idx_plus1 == 0 || m.bcw_idx_plus1 == 3) ? 0 : m.bcw_idx_plus1;
        c.prof_dir = m.tool == 2 ? (int)(m.aux & 3) : 0; c.prof_idx = (int)(m.aux >> 8);
        if (c.bi) {
            c.list[0] = gpm ? (m.aux >> 8) & 1 : 0; c.list[1] = gpm ? (m.aux >> 9) & 1 : 1;
            c.ridx[0] = m.ref_idx0; c.ridx[1] = m.ref_idx1;
            c.mvx[0] = m.mv0_x; c.mvy[0] = m.mv0_y; c.mvx[1] = m.mv1_x; c.mvy[1] = m.mv1_y;
        } else {
            const int l1 = (m.inter_dir & 2) || identical;
            c.list[0] = l1; c.ridx[0] = l1 ? m.ref_idx1 : m.ref_idx0; c.mvx[0] = l1 ? m.mv1_x : m.mv0_x; c.mvy[0] = l1 ? m.mv1_y : m.mv0_y;
            c.list[1] = 0; c.ridx[1] = 0; c.mvx[1] = c.mvy[1] = 0;
        }
        for (int l = 0; l < 2; l++)
            c.scaled[l] = (l == 0 || c.bi) && (a.scale[c.list[l]][c.ridx[l]][0] != (1 << 14) || a.scale[c.list[l]][c.ridx[l]][1] != (1 << 14));
        return c;
    }

    // sampling grid of one list in one component
    struct Grid {
        int scaled, chroma, shift_mv, nt, before;
        int fih, fiv, step_x, step_y;         // scaled: filter sets, steps
        int pos_x, pos_y;                     // scaled: clipped ref_pos_x / _y; unscaled: integer position of the PU's sample (0, 0)
        int frac_x, frac_y;                   // unscaled: the MV's fractions
        int hpel, small;
        const uint16_t *plane; int pitch, w, h;
    };

    OVD int filter_set(int scale, int flag_4x4)
    {   // compute_rpr_filter_idx, rcn_inter.c:1990-2006
        int idx = flag_4x4 ? 3 : 0;
        if (scale > (1 << 14) * 7 / 4) idx += 2;
        else if (scale > (1 << 14) * 5 / 4) idx += 1;
        return idx;
    }

    OVD Grid grid(const Args &a, const Cmd &c, int l, int comp)
    {
        Grid g;
        const int list = c.list[l], ridx = c.ridx[l], sh = comp ? 1 : 0;
        g.scaled = c.scaled[l]; g.chroma = comp != 0; g.shift_mv = comp ? 5 : 4; g.nt = comp ? 4 : 8; g.before = comp ? 1 : 3;
        g.plane = a.mc.refs.plane[list][ridx][comp]; g.pitch = a.mc.refs.pitch[list][ridx][comp];
        g.w = a.ref_w[list][ridx] >> sh; g.h = a.ref_h[list][ridx] >> sh;
        g.small = c.log2_w == 2 && c.log2_h == 2;
        g.hpel = 0; g.fih = g.fiv = 0; g.step_x = g.step_y = 0; g.frac_x = g.frac_y = 0;
        const int pu_w = (1 << c.log2_w) >> sh, pu_h = (1 << c.log2_h) >> sh, px = c.x0 >> sh, py = c.y0 >> sh;
        if (!g.scaled) {
            // rcn_mcp_bidir0_l / _c: clip_mv (rcn_inter.c:96-108) with the PU's luma position and size, then the fraction split
            const int lw = 1 << c.log2_w, lh = 1 << c.log2_h;
            const int mvx = ovd_clamp(c.mvx[l], -((lw + 3 + c.x0) << 4), (a.mc.pic_w + 2 - c.x0) << 4);
            const int mvy = ovd_clamp(c.mvy[l], -((lh + 3 + c.y0) << 4), (a.mc.pic_h + 2 - c.y0) << 4);
            g.frac_x = mvx & ((1 << g.shift_mv) - 1); g.frac_y = mvy & ((1 << g.shift_mv) - 1);
            g.pos_x = px + (mvx >> g.shift_mv); g.pos_y = py + (mvy >> g.shift_mv);
            g.hpel = !comp && (c.flags & MC_HPEL_AMVR) && !g.small;
            return g;
        }
        const int sch = a.scale[list][ridx][0], scv = a.scale[list][ridx][1];
        const int shift_pos = 14 + g.shift_mv, offset = 1 << 13;
        g.fih = filter_set(sch, g.small); g.fiv = filter_set(scv, g.small);
        if (comp) { g.fih %= 3; g.fiv %= 3; }
        g.step_x = ((sch + 8) >> 4) << 4; g.step_y = ((scv + 8) >> 4) << 4;
        // 32-bit products as the reference computes them
        int rx = (int)((uint32_t)((px << g.shift_mv) + c.mvx[l]) * (uint32_t)sch);
        int ry = (int)((uint32_t)((py << g.shift_mv) + c.mvy[l]) * (uint32_t)scv);
        if (comp) {
            rx += (1 - a.hor_col) * 8 * (sch - (1 << 14)) + (1 << 8);
            ry += (1 - a.ver_col) * 8 * (scv - (1 << 14)) + (1 << 8);
        } else { rx += 1 << 7; ry += 1 << 7; }
        const int ref_x = (rx + offset) >> shift_pos, ref_y = (ry + offset) >> shift_pos;
        const int ref_pu_w = ((rx + (((pu_w - 1) * g.step_x) << g.shift_mv) + offset) >> shift_pos) - ref_x + 1;
        const int ref_pu_h = ovd_max(1, ((ry + (((pu_h - 1) * g.step_y) << g.shift_mv) + offset) >> shift_pos) - ref_y + 1);
        // clip_rpr_position, rcn_inter.c:2009-2022
        const int prec_x = rx & ((1 << shift_pos) - 1), prec_y = ry & ((1 << shift_pos) - 1);
        g.pos_x = ovd_clamp(rx, -((ref_pu_w + 4) << shift_pos) + prec_x, ((g.w + 3) << shift_pos) + prec_x);
        g.pos_y = ovd_clamp(ry, -((ref_pu_h + 4) << shift_pos) + prec_y, ((g.h + 3) << shift_pos) + prec_y);
        return g;
    }

    // integer position and taps of output column `col` / row `row` of the PU (component samples)
    OVD const int8_t *taps_of(const Grid &g, int set, int frac)
    {
        if (g.chroma) return mc_rpr_c[set][frac];
        if (g.hpel && frac == 8) return mc_hpel_l;
        return mc_rpr_l[set][frac];
    }
    OVD const int8_t *col_taps(const Grid &g, int col, int *xi)
    {
        if (!g.scaled) { *xi = g.pos_x + col; return taps_of(g, g.small && !g.chroma ? 3 : 0, g.frac_x); }
        const int p = (g.pos_x + ((col * g.step_x) << g.shift_mv) + (1 << 13)) >> 14;
        *xi = p >> g.shift_mv;
        return taps_of(g, g.fih, p & ((1 << g.shift_mv) - 1));
    }
    OVD const int8_t *row_taps(const Grid &g, int row, int *yi, int *quirk)
    {
        *quirk = 0;
        if (!g.scaled) { *yi = g.pos_y + row; return taps_of(g, g.small && !g.chroma ? 3 : 0, g.frac_y); }
        const int p = (g.pos_y + ((row * g.step_y) << g.shift_mv) + (1 << 13)) >> 14;
        *yi = p >> g.shift_mv;
        *quirk = !g.fiv && !(p & ((1 << g.shift_mv) - 1));          // rpr_v_uni[0]: put_vvc_pel_rpr_clip
        return taps_of(g, g.fiv, p & ((1 << g.shift_mv) - 1));
    }

    OVD int16_t *tmp_of(int16_t *wsm, int comp) { return wsm + (comp == 0 ? 0 : TMP_Y_ROWS * 16 + (comp - 1) * TMP_C_ROWS * 8); }
    OVD int16_t *ex_of(int16_t *wsm, int l, int comp) { return wsm + SM_TMP + l * 384 + (comp == 0 ? 0 : 256 + (comp - 1) * 64); }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31;
        const Cmd c = load(a, bx * WARPS + warp);
        if (!c.valid) return;
        int16_t *wsm = (int16_t *)(smem + warp * SM_WARP);
        const int bd = a.mc.bitdepth;
        if (P < 4) {
            const int l = P >> 1;
            if (l == 1 && !c.bi) return;
            for (int comp = 0; comp < 3; comp++) {
                if (!(c.flags & (comp ? MC_CHROMA : MC_LUMA))) continue;
                const int sh = comp ? 1 : 0, w = c.w >> sh, h = c.h >> sh, ox = c.ox >> sh, oy = c.oy >> sh;
                const int lw = ovd_floor_log2((unsigned)w), cap = comp ? TMP_C_ROWS : TMP_Y_ROWS, ts = comp ? 8 : 16;
                const Grid g = grid(a, c, l, comp);
                int y_first, y_last, q;
                row_taps(g, oy, &y_first, &q); row_taps(g, oy + h - 1, &y_last, &q);
                y_first -= g.before;
                const int nrows = ovd_min(y_last + (g.nt - 1 - g.before) - y_first + 1, cap);
                int16_t *tmp = tmp_of(wsm, comp);
                if ((P & 1) == 0) {
                    // horizontal pass: put_vvc_{q,e}pel_rpr_h (put_vvc_pel_rpr == the identity row of set 0)
                    for (int i = lane; i < (nrows << lw); i += 32) {
                        const int r = i >> lw, col = i & (w - 1);
                        int xi;
                        const int8_t *f = col_taps(g, ox + col, &xi);
                        const uint16_t *row = g.plane + (size_t)ovd_clamp(y_first + r, 0, g.h - 1) * g.pitch;
                        int acc = 0;
                        for (int k = 0; k < g.nt; k++) acc += f[k] * (int)row[ovd_clamp(xi + k - g.before, 0, g.w - 1)];
                        tmp[r * ts + col] = (int16_t)(acc >> (bd - 8));
                    }
                } else {
                    // vertical pass: put_vvc_*_rpr_bi_v / _clip_v before the final rounding
                    int16_t *ex = ex_of(wsm, l, comp);
                    for (int i = lane; i < (h << lw); i += 32) {
                        const int y = i >> lw, x = i & (w - 1);
                        int yi, quirk;
                        const int8_t *f = row_taps(g, oy + y, &yi, &quirk);
                        const int rb = ovd_clamp(yi - g.before - y_first, 0, cap - g.nt);
                        int acc = 0;
                        for (int k = 0; k < g.nt; k++) acc += f[k] * (int)tmp[(rb + k) * ts + x];
                        int v = (int16_t)(acc >> 6);
                        // uni-prediction, integer row position, regular filter set: the reference reads the intermediate as UNSIGNED
                        // (put_vvc_pel_rpr_clip, rcn_mc.c:623-642), so a negative one ends up at the maximum sample value
                        if (!c.bi && quirk && v < 0) v = 32767;
                        ex[y * ts + x] = (int16_t)v;
                    }
                }
            }
            return;
        }
        // ---- phase 4: combination
        const int maxv = (1 << bd) - 1;
        GpmGeom gg;
        if (c.gpm_split >= 0) gg = gpm_geom(c.gpm_split, c.log2_w, c.log2_h);
        const int w1 = c.bcw ? mc_bcw_w(c.bcw) : 0;
        for (int comp = 0; comp < 3; comp++) {
            if (!(c.flags & (comp ? MC_CHROMA : MC_LUMA))) continue;
            const int sh = comp ? 1 : 0, w = c.w >> sh, h = c.h >> sh, lw = ovd_floor_log2((unsigned)w), ts = comp ? 8 : 16;
            const int16_t *e0 = ex_of(wsm, 0, comp), *e1 = ex_of(wsm, 1, comp);
            uint16_t *d = a.mc.dst[comp] + (size_t)((c.y0 + c.oy) >> sh) * a.mc.dst_pitch[comp] + ((c.x0 + c.ox) >> sh);
            for (int i = lane; i < (h << lw); i += 32) {
                const int y = i >> lw, x = i & (w - 1);
                int p0 = e0[y * ts + x], p1 = c.bi ? e1[y * ts + x] : 0, v;
                if (c.bi && c.prof_dir && comp == 0) {      // rcn_mc_rpr_prof_b_l: PROF on the unscaled list (4x4 luma block)
                    if (!c.scaled[0] && (c.prof_dir & 1)) p0 = prof_refine(a, c, 0, e0, x, y);
                    if (!c.scaled[1] && (c.prof_dir & 2)) p1 = prof_refine(a, c, 1, e1, x, y);
                }
                if (!c.bi) v = (p0 + (1 << (13 - bd))) >> (14 - bd);
                else if (c.gpm_split >= 0) {
                    const int g0 = gpm_weight(gg, (c.ox >> sh << sh) + (x << sh), (c.oy >> sh << sh) + (y << sh));
                    v = (p1 * (8 - g0) + p0 * g0 + (1 << (16 - bd))) >> (17 - bd);
                } else if (!c.bcw) v = (p0 + p1 + (1 << (14 - bd))) >> (15 - bd);
                else v = (p1 * w1 + p0 * (8 - w1) + (1 << (16 - bd))) >> (17 - bd);
                v = ovd_clamp(v, 0, maxv);
                if (comp == 0 && a.mc.lmcs_fwd) v = a.mc.lmcs_fwd[v];
                d[(size_t)y * a.mc.dst_pitch[comp] + x] = (uint16_t)v;
            }
        }
    }

    OVD int mc_bcw_w(int bcw_idx_plus1)
    {   // bcw_weights, rcn_inter.c:89
        return bcw_idx_plus1 == 1 ? -2 : bcw_idx_plus1 == 2 ? 3 : bcw_idx_plus1 == 3 ? 4 : bcw_idx_plus1 == 4 ? 5 : 10;
    }

    // rcn_prof_mcp_bi_l (rcn_inter.c:1729-1820) on the 4x4 block whose 14-bit prediction is e[]: gradients from the block padded with
    // the integer reference samples next to it (rcn_prof_bdof.c:153-171), offset clipped to 14 bits, int16 sum
    OVD int prof_refine(const Args &a, const Cmd &c, int l, const int16_t *e, int x, int y)
    {
        const Grid g = grid(a, c, l, 0);
        const int exx = (g.hpel && g.frac_x == 8 ? 16 : g.frac_x) >> 3 != 0, eyy = (g.hpel && g.frac_y == 8 ? 16 : g.frac_y) >> 3 != 0;
        const int sh = 14 - a.mc.bitdepth;
        int P[4];                                 // left, right, above, below
        const int nx[4] = { x - 1, x + 1, x, x }, ny[4] = { y, y, y - 1, y + 1 };
        for (int k = 0; k < 4; k++) {
            if (nx[k] >= 0 && nx[k] < 4 && ny[k] >= 0 && ny[k] < 4) P[k] = e[ny[k] * 16 + nx[k]];
            else {
                const int rx = ovd_clamp(g.pos_x + nx[k] + exx, 0, g.w - 1), ry = ovd_clamp(g.pos_y + ny[k] + eyy, 0, g.h - 1);
                P[k] = (int16_t)((int)g.plane[(size_t)ry * g.pitch + rx] << sh);
            }
        }
        const int gx = (int16_t)(((P[1] - 8192) >> 6) - ((P[0] - 8192) >> 6)), gy = (int16_t)(((P[3] - 8192) >> 6) - ((P[2] - 8192) >> 6));
        const ProfInfoDev *pi = &a.prof[c.prof_idx];
        const int16_t *dh = pi->dmv[2 * l], *dv = pi->dmv[2 * l + 1];
        return (int16_t)(e[y * 16 + x] + ovd_clamp(dh[4 * y + x] * gx + dv[4 * y + x] * gy, -8192, 8191));
    }
};
