// mc_tools_kernels.h -- BDOF and PROF motion compensation (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour:
//   BDOF  rcn_bdof_mcp_l (libovvc/rcn_inter.c:1137-1250): 14-bit predictions of both lists, a one-sample ring taken
//         from the nearest integer reference samples (extend_bdof_buff, libovvc/rcn_prof_bdof.c:295-345), gradients
//         (compute_prof_grad :153-171), replicated padding of gradients and predictions (extend_bdof_grad,
//         rcn_inter.c:840-870), per 4x4 sub-block weights from the 6x6 neighbourhood (derive_bdof_weights :347-422) and
//         the corrected average (rcn_apply_bdof_subblock :60-99, int16 cast before the clip), forward LMCS.
//   PROF  rcn_prof_mcp_b_l -> rcn_prof_motion_compensation_b_l / rcn_prof_mcp_l (rcn_inter.c:2865-2919, 1253-1388,
//         1632-1727): 4x4 affine sub-block with the 6-tap filters, integer-sample ring (extend_prof_buff :173-221),
//         gradients, per-sample refinement clip(dh*gx + dv*gy, -8192, 8191) (rcn_prof :223-282), average / BCW / uni
//         rounding, forward LMCS.
//
// BdofK: one warp per <=16x16 sub-block.  ProfK: one warp per FOUR 4x4 sub-blocks (lanes = sub-block x list x column
// while filtering).  Both reuse the column-per-lane interpolation of mc_kernels.h.
#pragma once
#include "mc_kernels.h"

#define MC_TOOL_BDOF 1
#define MC_TOOL_PROF 2

struct ProfInfoDev { int16_t dmv[4][16]; };       // == OVGPUProfInfo: dmv_scale_h_0, v_0, h_1, v_1

// ---- shared helpers ----------------------------------------------------------------------------------------------
struct McToolCommon {
    // clip_mv (rcn_inter.c:96-108) of one MV for a block at (x0, y0) of size w x h
    OVD void clip_mv(int pic_w, int pic_h, int x0, int y0, int w, int h, int &mvx, int &mvy)
    {
        mvx = ovd_clamp(mvx, -((w + 3 + x0) << 4), (pic_w + 2 - x0) << 4);
        mvy = ovd_clamp(mvy, -((h + 3 + y0) << 4), (pic_h + 2 - y0) << 4);
    }
    // luma filter rows of mc_cw_l for a fraction (half-pel AMVR row 16, 6-tap set for 4x4 blocks)
    OVD int filter_row(int frac, int hpel, int small)
    {
        if (hpel && !small && frac == 8) frac = 16;
        return small ? frac + 17 : frac;
    }
};

// ---- BDOF on two 14-bit predictions held in shared memory (ex[list][16][16]) ------------------------------------------
struct BdofCore {
    // where the integer reference samples of the one-sample ring live: sample at padded position (x, y), x in -1..w,
    // is the 16-bit sample (col0 + x, row0 + y) of a window stored as packed pairs with ww words per row
    struct Ring { const uint32_t *win; int ww, col0, row0; };

    OVD int floor_log2(unsigned v) { return ovd_floor_log2(v); }
    OVD int ring_sample(const Ring &g, int x, int y, int sh)
    {
        const int col = g.col0 + x;
        const uint32_t wd = g.win[(g.row0 + y) * g.ww + (col >> 1)];
        return (int)(int16_t)(((col & 1) ? wd >> 16 : wd & 0xffff) << sh);
    }
    OVD int pint(const int16_t *ex, const Ring &g, int r, int x, int y, int iw, int ih, int sh)
    {
        return (x >= 0 && x < iw && y >= 0 && y < ih) ? (int)ex[r * 256 + y * 16 + x] : ring_sample(g, x, y, sh);
    }
    // gradients of the interior (compute_prof_grad, rcn_prof_bdof.c:153-171): G = [gx0, gy0, gx1, gy1][16][16]
    OVD void grad(const int16_t *ex, int16_t *G, int iw, int ih, const Ring *ring, int sh, int lane)
    {
        const int lw = iw == 16 ? 4 : 3;                                   // sub-blocks are 8 or 16 samples wide: no integer division
        for (int i = lane; i < iw * ih; i += 32) {
            const int x = i & (iw - 1), y = i >> lw;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int gx = ((pint(ex, ring[r], r, x + 1, y, iw, ih, sh) - 8192) >> 6) - ((pint(ex, ring[r], r, x - 1, y, iw, ih, sh) - 8192) >> 6);
                const int gy = ((pint(ex, ring[r], r, x, y + 1, iw, ih, sh) - 8192) >> 6) - ((pint(ex, ring[r], r, x, y - 1, iw, ih, sh) - 8192) >> 6);
                G[(2 * r) * 256 + y * 16 + x] = (int16_t)gx;
                G[(2 * r + 1) * 256 + y * 16 + x] = (int16_t)gy;
            }
        }
    }
    // partial sums of derive_bdof_weights (:347-422): lane = (sub-block, half of the 6 rows); the replicated padding of
    // gradients and predictions (extend_bdof_grad) = clamped coordinates
    OVD void partial(const int16_t *ex, const int16_t *G, int *part, int iw, int ih, int lane)
    {
        const int sb = lane >> 1, half = lane & 1, nsx = iw >> 2, nsb = nsx * (ih >> 2);
        if (sb >= nsb) return;
        const int sx = (sb & (nsx - 1)) * 4, sy = (sb >> (nsx >> 1)) * 4;    // nsx is 2 or 4
        int s_ax = 0, s_ay = 0, s_dx = 0, s_dy = 0, s_sign = 0;
        for (int i = 3 * half; i < 3 * half + 3; i++) {
            const int y = ovd_clamp(sy + i - 1, 0, ih - 1);
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int x = ovd_clamp(sx + j - 1, 0, iw - 1), o = y * 16 + x;
                const int ax = (G[o] + G[2 * 256 + o]) >> 1, ay = (G[256 + o] + G[3 * 256 + o]) >> 1;
                const int dr = ((ex[256 + o] - 8192) >> 4) - ((ex[o] - 8192) >> 4);
                s_ax += ax < 0 ? -ax : ax; s_ay += ay < 0 ? -ay : ay;
                s_sign += ay < 0 ? -ax : (ay == 0 ? 0 : ax);
                s_dx += ax < 0 ? -dr : (ax == 0 ? 0 : dr);
                s_dy += ay < 0 ? -dr : (ay == 0 ? 0 : dr);
            }
        }
        int *p = part + (sb * 2 + half) * 5;
        p[0] = s_ax; p[1] = s_ay; p[2] = s_dx; p[3] = s_dy; p[4] = s_sign;
    }
    // weights of the two sub-blocks a lane's 8 samples fall in, corrected average (rcn_apply_bdof_subblock :60-99), LMCS, store
    OVD void apply(const McK::Args &a, const int16_t *ex, const int16_t *G, const int *part, int iw, int ih, int ibx, int iby, int lane)
    {
        const int y = lane >> 1, x0 = (lane & 1) * 8;
        if (y >= ih || x0 >= iw) return;
        const int nsx = iw >> 2, shift = 15 - a.bitdepth, off = 1 << (shift - 1), maxv = (1 << a.bitdepth) - 1;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int sb = (y >> 2) * nsx + (x0 >> 2) + q;
            const int *p0 = part + sb * 10, *p1 = p0 + 5;
            const int s_ax = p0[0] + p1[0], s_ay = p0[1] + p1[1], s_dx = p0[2] + p1[2], s_dy = p0[3] + p1[3], s_sign = p0[4] + p1[4];
            int wx = 0, wy = 0;
            if (s_ax) wx = ovd_clamp((s_dx << 2) >> floor_log2((unsigned)s_ax), -15, 15);
            if (s_ay) {
                int xo = 0;
                if (wx) { const int hi = s_sign >> 12, lo = s_sign & 4095; xo = (((wx * hi) << 12) + (wx * lo)) >> 1; }
                wy = ovd_clamp(((s_dy << 2) - xo) >> floor_log2((unsigned)s_ay), -15, 15);
            }
            int v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = y * 16 + x0 + 4 * q + k;
                const int b = wx * (G[i] - G[2 * 256 + i]) + wy * (G[256 + i] - G[3 * 256 + i]);
                v[k] = ovd_clamp((int)(int16_t)((ex[i] + ex[256 + i] + b + off) >> shift), 0, maxv);
                if (a.lmcs_fwd) v[k] = a.lmcs_fwd[v[k]];
            }
            o[2 * q] = (uint32_t)v[0] | ((uint32_t)v[1] << 16); o[2 * q + 1] = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
        }
        uint16_t *d = a.dst[0] + (size_t)(iby + y) * a.dst_pitch[0] + ibx + x0;
        ovd_u2 r0, r1;                                                     // block positions are multiples of 4 samples
        r0.x = o[0]; r0.y = o[1]; ovd_st8(d, r0);
        r1.x = o[2]; r1.y = o[3]; ovd_st8(d + 4, r1);
    }
};

// =================================================================================================================
struct BdofK {
    static const int THREADS = 128, MIN_BLOCKS = 6, PHASES = 5;
    static const int WARPS = 4;
    static const int WW = McK::WIN_Y_WORDS, WR = McK::WIN_Y_ROWS;
    static const int SM_ITEM = 16, SM_WIN = 2 * WR * WW, SM_EX = 2 * 16 * 16 / 2, SM_G = 4 * 16 * 16 / 2, SM_PART = 16 * 2 * 5;
    static const int SM_WARP = (SM_ITEM + SM_WIN + SM_EX + SM_G + SM_PART) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;
    typedef McK::Args Args;
    struct State { int dummy; };

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t *sm = (uint32_t *)(smem + warp * SM_WARP);
        int *sit = (int *)sm;
        uint32_t *win = sm + SM_ITEM;
        int16_t *ex = (int16_t *)(win + SM_WIN);             // [list][16][16] 14-bit predictions
        int16_t *G = ex + 2 * 256;                           // [gx0, gy0, gx1, gy1][16][16]
        int *part = (int *)(G + 4 * 256);                    // [sub-block][half][5]
        const int item = bx * WARPS + warp;

        if (P == 0) {
            if (lane == 0) sit[0] = item < McK::nb_items(a);
            if (item >= McK::nb_items(a)) return;
            const McCmdDev c = a.cmds[a.items[item].cmd];
            const int w = 1 << c.log2_w, h = 1 << c.log2_h, hpel = (c.flags & MC_HPEL_AMVR) != 0;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                int mvx = r ? c.mv1_x : c.mv0_x, mvy = r ? c.mv1_y : c.mv0_y;
                McToolCommon::clip_mv(a.pic_w, a.pic_h, c.x0, c.y0, w, h, mvx, mvy);
                const int ridx = r ? c.ref_idx1 : c.ref_idx0;
                const uint16_t *pl = a.refs.plane[r][ridx][0];
                const int pitch = a.refs.pitch[r][ridx][0], nrows = h + 7;
                const int xs = c.x0 + (mvx >> 4) - 3, ys = c.y0 + (mvy >> 4) - 3, x_al = xs & ~3;
                uint32_t *wv = win + r * WR * WW;
                if (lane == 0) {
                    const int px = mvx & 15, py = mvy & 15;
                    sit[4 + 4 * r] = xs - x_al;
                    sit[5 + 4 * r] = McToolCommon::filter_row(px, hpel, 0) | (McToolCommon::filter_row(py, hpel, 0) << 8);
                    sit[6 + 4 * r] = ((px >> 3) != 0) | (((py >> 3) != 0) << 1);          // ring offsets ext_x / ext_y
                }
                if (x_al >= 0 && x_al + 2 * WW <= a.pic_w && ys >= 0 && ys + nrows <= a.pic_h) {
                    const int cy = lane & 7, ry = lane >> 3;
                    ovd_u2 v[6];
                    if (cy < 7) {
#pragma unroll
                        for (int k = 0; k < 6; k++)
                            if (ry + 4 * k < nrows) v[k] = ovd_ld8(pl + (size_t)(ys + ry + 4 * k) * pitch + x_al + 4 * cy);
#pragma unroll
                        for (int k = 0; k < 6; k++)
                            if (ry + 4 * k < nrows) { wv[(ry + 4 * k) * WW + 2 * cy] = v[k].x; wv[(ry + 4 * k) * WW + 2 * cy + 1] = v[k].y; }
                    }
                } else {
                    for (int i = lane; i < nrows * WW; i += 32) {
                        const int row = i / WW, wd = i - row * WW;
                        wv[i] = McK::load_pair(pl, pitch, a.pic_w, a.pic_h, x_al + 2 * wd, ys + row);
                    }
                }
            }
            if (lane == 0) { sit[1] = c.x0; sit[2] = c.y0; sit[3] = w | (h << 8); }
            return;
        }
        if (!sit[0]) return;
        const int ibx = sit[1], iby = sit[2], iw = sit[3] & 0xff, ih = sit[3] >> 8;

        if (P == 1) {
            // 14-bit predictions of both lists: lanes 0-15 list 0, 16-31 list 1, one column each
            const int r = lane >> 4, x = lane & 15, sh1 = a.bitdepth - 8;
            if (x < iw) {
                const int o = sit[4 + 4 * r] + x, frl = sit[5 + 4 * r];
                uint32_t Bh[5], Bv0[4], Bv1[5];
#pragma unroll
                for (int j = 0; j < 5; j++) { Bh[j] = mc_cw_l[frl & 0xff][o & 1][j]; Bv1[j] = mc_cw_l[frl >> 8][1][j]; }
#pragma unroll
                for (int j = 0; j < 4; j++) Bv0[j] = mc_cw_l[frl >> 8][0][j];
                const uint32_t *wv = win + r * WR * WW + (o >> 1);
                int16_t *e = ex + r * 256 + x;
                if (ih == 16) McK::column<16, 8, 5>(wv, WW, Bh, Bv0, Bv1, sh1, e, 16);
                else          McK::column<8, 8, 5>(wv, WW, Bh, Bv0, Bv1, sh1, e, 16);
            }
            return;
        }

        // ring = integer reference samples of the staged window (extend_bdof_buff)
        BdofCore::Ring ring[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int ext = sit[6 + 4 * r];
            ring[r].win = win + r * WR * WW; ring[r].ww = WW;
            ring[r].col0 = sit[4 + 4 * r] + 3 + (ext & 1); ring[r].row0 = 3 + (ext >> 1);
        }
        if (P == 2) { BdofCore::grad(ex, G, iw, ih, ring, 14 - a.bitdepth, lane); return; }
        if (P == 3) { BdofCore::partial(ex, G, part, iw, ih, lane); return; }
        if (P == 4) BdofCore::apply(a, ex, G, part, iw, ih, ibx, iby, lane);
    }
};

// =================================================================================================================
#ifndef PROF_MIN_BLOCKS
#define PROF_MIN_BLOCKS 4
#endif
struct ProfK {
    static const int THREADS = 128, MIN_BLOCKS = PROF_MIN_BLOCKS, PHASES = 4;
    static const int WARPS = 4, SBS = 4;                     // 4 sub-blocks (commands) per warp
    static const int WW = 8, WR = 11;                        // 16 samples x 11 rows per sub-block and list
    static const int SM_ITEM = SBS * 16, SM_WIN = SBS * 2 * WR * WW, SM_EX = SBS * 2 * 16 / 2;
    static const int SM_WARP = (SM_ITEM + SM_WIN + 2 * SM_EX) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        McK::Args mc;
        const ProfInfoDev *prof;
    };
    struct State { int dummy; };

    template <int P>
    OVD void phase(const Args &aa, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const McK::Args &a = aa.mc;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t *sm = (uint32_t *)(smem + warp * SM_WARP);
        int *sit = (int *)sm;                                // [sub-block][16]
        uint32_t *win = sm + SM_ITEM;                        // [sub-block][list][row][word]
        int16_t *ex = (int16_t *)(win + SM_WIN);             // [sub-block][list][4][4] 14-bit predictions
        int16_t *ex2 = ex + SBS * 2 * 16;                    // refined predictions
        const int sb = lane >> 3, item = (bx * WARPS + warp) * SBS + sb;
        int *si = sit + sb * 16;

        if (P == 0) {
            const int r = (lane >> 2) & 1, cc = lane & 3;
            if ((lane & 7) == 0) si[0] = item < McK::nb_items(a);
            if (item >= McK::nb_items(a)) return;
            const McCmdDev c = a.cmds[a.items[item].cmd];
            const int used = (c.inter_dir >> r) & 1;
            int mvx = r ? c.mv1_x : c.mv0_x, mvy = r ? c.mv1_y : c.mv0_y;
            McToolCommon::clip_mv(a.pic_w, a.pic_h, c.x0, c.y0, 4, 4, mvx, mvy);
            const int xs = c.x0 + (mvx >> 4) - 3, ys = c.y0 + (mvy >> 4) - 3, x_al = xs & ~3;
            if (cc == 0) {
                const int px = mvx & 15, py = mvy & 15;
                si[4 + 4 * r] = xs - x_al;
                si[5 + 4 * r] = McToolCommon::filter_row(px, 0, 1) | (McToolCommon::filter_row(py, 0, 1) << 8);
                si[6 + 4 * r] = ((px >> 3) != 0) | (((py >> 3) != 0) << 1);
                // bit 2: a plain (tool 0) 4x4 luma command routed here by mc_route: same prediction, no refinement
                if (r == 0) { si[1] = c.x0; si[2] = c.y0; si[3] = c.tool == 0 ? (c.inter_dir | 4 | (c.bcw_idx_plus1 << 8))
                                                                              : (c.inter_dir | ((c.aux & 3) << 4) | (c.bcw_idx_plus1 << 8) | ((c.aux >> 8) << 12)); }
            }
            if (!used) return;
            const int ridx = r ? c.ref_idx1 : c.ref_idx0;
            const uint16_t *pl = a.refs.plane[r][ridx][0];
            const int pitch = a.refs.pitch[r][ridx][0];
            uint32_t *wv = win + (sb * 2 + r) * WR * WW;
            if (x_al >= 0 && x_al + 2 * WW <= a.pic_w && ys >= 0 && ys + WR <= a.pic_h) {
                ovd_u2 v[WR];
#pragma unroll
                for (int k = 0; k < WR; k++) v[k] = ovd_ld8(pl + (size_t)(ys + k) * pitch + x_al + 4 * cc);
#pragma unroll
                for (int k = 0; k < WR; k++) { wv[k * WW + 2 * cc] = v[k].x; wv[k * WW + 2 * cc + 1] = v[k].y; }
            } else {
                for (int i = cc; i < WR * WW; i += 4) {
                    const int row = i / WW, wd = i - row * WW;
                    wv[i] = McK::load_pair(pl, pitch, a.pic_w, a.pic_h, x_al + 2 * wd, ys + row);
                }
            }
            return;
        }
        if (!si[0]) return;
        const int info = si[3], inter_dir = info & 3, prof_dir = (info >> 4) & 3, bcw_raw = (info >> 8) & 15, tab = info >> 12;

        if (P == 1) {
            // 4x4 prediction with the 6-tap set: lane = (sub-block, list, column)
            const int r = (lane >> 2) & 1, x = lane & 3, sh1 = a.bitdepth - 8;
            if (!((inter_dir >> r) & 1)) return;
            const int o = si[4 + 4 * r] + x, frl = si[5 + 4 * r];
            uint32_t Bh[5], Bv0[4], Bv1[5];
#pragma unroll
            for (int j = 0; j < 5; j++) { Bh[j] = mc_cw_l[frl & 0xff][o & 1][j]; Bv1[j] = mc_cw_l[frl >> 8][1][j]; }
#pragma unroll
            for (int j = 0; j < 4; j++) Bv0[j] = mc_cw_l[frl >> 8][0][j];
            McK::column<4, 8, 5>(win + (sb * 2 + r) * WR * WW + (o >> 1), WW, Bh, Bv0, Bv1, sh1, ex + (sb * 2 + r) * 16 + x, 4);
            return;
        }
        if (P == 2) {
            // gradients + refinement: lane = (sub-block, list, row), 4 samples each
            const int r = (lane >> 2) & 1, y = lane & 3;
            if (!((inter_dir >> r) & 1)) return;
            const int apply = (info & 4) ? 0 : (inter_dir == 3 ? (prof_dir >> r) & 1 : 1);       // uni-prediction always refines (rcn_prof_mcp_l)
            const int16_t *e = ex + (sb * 2 + r) * 16;
            const int16_t *dh = aa.prof[tab].dmv[2 * r], *dv = aa.prof[tab].dmv[2 * r + 1];
            const int ext = si[6 + 4 * r], c0 = si[4 + 4 * r] + 3 + (ext & 1), r0 = 3 + (ext >> 1), sh = 14 - a.bitdepth;
            const uint32_t *wv = win + (sb * 2 + r) * WR * WW;
#define PROF_RING(x_, y_) ((int)(int16_t)((((c0 + (x_)) & 1 ? wv[(r0 + (y_)) * WW + ((c0 + (x_)) >> 1)] >> 16 : wv[(r0 + (y_)) * WW + ((c0 + (x_)) >> 1)] & 0xffff)) << sh))
#define PROF_PINT(x_, y_) (((x_) >= 0 && (x_) < 4 && (y_) >= 0 && (y_) < 4) ? (int)e[(y_) * 4 + (x_)] : PROF_RING(x_, y_))
#pragma unroll
            for (int x = 0; x < 4; x++) {
                int v = e[y * 4 + x];
                if (apply) {
                    const int gx = (int16_t)(((PROF_PINT(x + 1, y) - 8192) >> 6) - ((PROF_PINT(x - 1, y) - 8192) >> 6));
                    const int gy = (int16_t)(((PROF_PINT(x, y + 1) - 8192) >> 6) - ((PROF_PINT(x, y - 1) - 8192) >> 6));
                    v = (int16_t)(v + ovd_clamp(dh[4 * y + x] * gx + dv[4 * y + x] * gy, -8192, 8191));
                }
                ex2[(sb * 2 + r) * 16 + y * 4 + x] = (int16_t)v;
            }
#undef PROF_PINT
#undef PROF_RING
            return;
        }
        if (P == 3) {
            // combine + LMCS + store: lane = (sub-block, row, half): 2 samples
            const int y = (lane >> 1) & 3, x0 = (lane & 1) * 2, maxv = (1 << a.bitdepth) - 1;
            const int bcw = (bcw_raw == 0 || bcw_raw == 3) ? 0 : bcw_raw;
            const int w1 = bcw == 1 ? -2 : bcw == 2 ? 3 : bcw == 4 ? 5 : 10;
            const int16_t *e0 = ex2 + (sb * 2 + (inter_dir == 2 ? 1 : 0)) * 16, *e1 = ex2 + (sb * 2 + 1) * 16;
            int v[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int i = y * 4 + x0 + k;
                v[k] = McK::combine(e0[i], inter_dir == 3 ? e1[i] : 0, inter_dir == 3 ? 2 : 1, bcw, w1, a.bitdepth, maxv);
                if (a.lmcs_fwd) v[k] = a.lmcs_fwd[v[k]];
            }
            *(uint32_t *)(a.dst[0] + (size_t)(si[2] + y) * a.dst_pitch[0] + si[1] + x0) = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
        }
    }
};

// =================================================================================================================
// Small chroma blocks: plain (tool 0) chroma-only commands of 8x8 luma samples -- the Cb / Cr blocks of affine coding
// units (one per 2x2 luma sub-blocks with the averaged MV, drv_affine_mvp.c:3380-3410; rcn_mcp_b_c, rcn_inter.c:2921).  A
// whole McK warp per 4x4 chroma block wastes three quarters of its lanes; here a warp takes TWO commands:
// lane = (command, component, list, column).  Same arithmetic as McK's chroma path (4-tap, rcn_mc.c:904-1100).
#define MC_ROUTE_SMALL_C 5
struct McSmallCK {
    static const int THREADS = 128, MIN_BLOCKS = 6, PHASES = 3;
    static const int WARPS = 4, SBS = 2;                     // 2 commands per warp
    static const int WW = 6, WR = 7;                         // 12 samples x 7 rows per (command, component, list)
    static const int SM_ITEM = SBS * 16, SM_WIN = SBS * 4 * WR * WW, SM_EX = SBS * 4 * 16 / 2;
    static const int SM_WARP = (SM_ITEM + SM_WIN + SM_EX) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;
    typedef McK::Args Args;
    struct State { int dummy; };

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t *sm = (uint32_t *)(smem + warp * SM_WARP);
        int *sit = (int *)sm;                                // [command][16]
        uint32_t *win = sm + SM_ITEM;                        // [command][component][list][row][word]
        int16_t *ex = (int16_t *)(win + SM_WIN);             // [command][component][list][4][4] 14-bit predictions
        const int sb = lane >> 4, item = (bx * WARPS + warp) * SBS + sb;
        int *si = sit + sb * 16;
        const int cw = a.pic_w >> 1, ch = a.pic_h >> 1;

        if (P == 0) {
            const int comp = (lane >> 3) & 1, r = (lane >> 2) & 1, cc = lane & 3;
            if ((lane & 15) == 0) si[0] = item < McK::nb_items(a);
            if (item >= McK::nb_items(a)) return;
            const McCmdDev c = a.cmds[a.items[item].cmd];
            const int used = (c.inter_dir >> r) & 1;
            int mvx = r ? c.mv1_x : c.mv0_x, mvy = r ? c.mv1_y : c.mv0_y;
            McToolCommon::clip_mv(a.pic_w, a.pic_h, c.x0, c.y0, 8, 8, mvx, mvy);
            const int xc = (c.x0 >> 1) + (mvx >> 5) - 1, yc = (c.y0 >> 1) + (mvy >> 5) - 1, xc_al = xc & ~3;
            if (cc == 0 && comp == 0) {
                si[4 + 2 * r] = xc - xc_al;
                si[5 + 2 * r] = (mvx & 31) | ((mvy & 31) << 8);
                if (r == 0) { si[1] = c.x0 >> 1; si[2] = c.y0 >> 1; si[3] = c.inter_dir | (c.bcw_idx_plus1 << 8); }
            }
            if (!used || cc >= 3) return;
            const int ridx = r ? c.ref_idx1 : c.ref_idx0;
            const uint16_t *pl = a.refs.plane[r][ridx][1 + comp];
            const int pitch = a.refs.pitch[r][ridx][1 + comp];
            uint32_t *wv = win + ((sb * 2 + comp) * 2 + r) * WR * WW;
            if (xc_al >= 0 && xc_al + 2 * WW <= cw && yc >= 0 && yc + WR <= ch) {
                ovd_u2 v[WR];
#pragma unroll
                for (int k = 0; k < WR; k++) v[k] = ovd_ld8(pl + (size_t)(yc + k) * pitch + xc_al + 4 * cc);
#pragma unroll
                for (int k = 0; k < WR; k++) { wv[k * WW + 2 * cc] = v[k].x; wv[k * WW + 2 * cc + 1] = v[k].y; }
            } else {
                for (int i = cc; i < WR * WW; i += 3) {
                    const int row = i / WW, wd = i - row * WW;
                    wv[i] = McK::load_pair(pl, pitch, cw, ch, xc_al + 2 * wd, yc + row);
                }
            }
            return;
        }
        if (!si[0]) return;
        const int info = si[3], inter_dir = info & 3, bcw_raw = (info >> 8) & 15;
        if (P == 1) {
            const int comp = (lane >> 3) & 1, r = (lane >> 2) & 1, x = lane & 3, sh1 = a.bitdepth - 8;
            if (!((inter_dir >> r) & 1)) return;
            const int o = si[4 + 2 * r] + x, frc = si[5 + 2 * r];
            uint32_t Bh[3], Bv0[2], Bv1[3];
#pragma unroll
            for (int j = 0; j < 3; j++) { Bh[j] = mc_cw_c[frc & 0xff][o & 1][j]; Bv1[j] = mc_cw_c[frc >> 8][1][j]; }
#pragma unroll
            for (int j = 0; j < 2; j++) Bv0[j] = mc_cw_c[frc >> 8][0][j];
            McK::column<4, 4, 3>(win + ((sb * 2 + comp) * 2 + r) * WR * WW + (o >> 1), WW, Bh, Bv0, Bv1, sh1, ex + ((sb * 2 + comp) * 2 + r) * 16 + x, 4);
            return;
        }
        // combine + store: lane = (command, component, row, half): 2 samples
        const int comp = (lane >> 3) & 1, y = (lane >> 1) & 3, x0 = (lane & 1) * 2, maxv = (1 << a.bitdepth) - 1;
        const int bcw = (bcw_raw == 0 || bcw_raw == 3) ? 0 : bcw_raw;
        const int w1 = bcw == 1 ? -2 : bcw == 2 ? 3 : bcw == 4 ? 5 : 10;
        const int16_t *e0 = ex + ((sb * 2 + comp) * 2 + (inter_dir == 2 ? 1 : 0)) * 16, *e1 = ex + ((sb * 2 + comp) * 2 + 1) * 16;
        int v[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = y * 4 + x0 + k;
            v[k] = McK::combine(e0[i], inter_dir == 3 ? e1[i] : 0, inter_dir == 3 ? 2 : 1, bcw, w1, a.bitdepth, maxv);
        }
        uint16_t *d = (comp ? a.dst[2] : a.dst[1]) + (size_t)(si[2] + y) * (comp ? a.dst_pitch[2] : a.dst_pitch[1]) + si[1] + x0;
        *(uint32_t *)d = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
    }
};

// =================================================================================================================
// DMVR: rcn_dmvr_mv_refine (libovvc/rcn_inter.c:873-1126).  One warp per <=16x16 sub-block:
//   0  padded reference windows (luma 27x27 max, chroma 15x15 max) of both lists around the CLIPPED integer MV: picture
//      borders replicated (derive_dmvr_ref_buf_y/_c :432-518) and the window itself extended by replication (padd_dmvr
//      :326-378) -- both are coordinate clamps here
//   1  bilinear interpolation of (w + 4) x (h + 4) samples per list (rcn_mc.c:789-897), fraction of the UNCLIPPED MV
//   2  25 SADs on every other row, mirrored offsets (:615-755)
//   3  decision (one lane): early-out, integer minimum with the centre tie-break, parametric sub-sample refinement
//      (:758-833), refined MVs written back for the CPU's TMVP store
//   4  final 8-tap / 4-tap prediction from the padded windows at the refined MV
//   5-7 average, or BDOF on the refined predictions (when enabled and min_cost >= 2wh); chroma average; LMCS; stores
#define MC_TOOL_DMVR 3
struct DmvrMvDev { int32_t mv0_x, mv0_y, mv1_x, mv1_y; };      // == OVGPUDmvrMv

struct DmvrK {
    static const int THREADS = 128, MIN_BLOCKS = 3, PHASES = 8;
    static const int WARPS = 4;
    static const int PWW = 14, PWR = 27, PCW = 8, PCR = 15, BLS = 20;      // PCW must stay 8 (phase 0 splits indices with shifts)
    static const int SM_ITEM = 32, SM_PW = 2 * PWR * PWW, SM_PWC = 2 * 2 * PCR * PCW, SM_BL = 2 * BLS * BLS / 2, SM_SAD = 32,
                     SM_EX = 256, SM_EXC = 128, SM_G = 512, SM_PART = 160;
    static const int SM_WARP = (SM_ITEM + SM_PW + SM_PWC + SM_BL + SM_SAD + SM_EX + SM_EXC + SM_G + SM_PART) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        McK::Args mc;
        DmvrMvDev *out;           // refined MVs, one entry per work item
    };
    struct State { int dummy; };

    OVD int div_for_maxq7(int num, int den)
    {
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

    template <int P>
    OVD void phase(const Args &aa, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const McK::Args &a = aa.mc;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t *sm = (uint32_t *)(smem + warp * SM_WARP);
        int *sit = (int *)sm;
        uint32_t *pw = sm + SM_ITEM;                          // [list][27][14] luma, padded window
        uint32_t *pwc = pw + SM_PW;                           // [list][comp][15][8] chroma
        uint16_t *bl = (uint16_t *)(pwc + SM_PWC);            // [list][20][20] bilinear samples
        int *sadv = (int *)(pwc + SM_PWC + SM_BL);
        int16_t *ex = (int16_t *)(sadv + SM_SAD);
        int16_t *exc = ex + 2 * SM_EX;
        int16_t *G = exc + 2 * SM_EXC;
        int *part = (int *)(G + 2 * SM_G);
        const int item = bx * WARPS + warp;
        const int cw = a.pic_w >> 1, ch = a.pic_h >> 1;

        if (P == 0) {
            if (lane == 0) sit[0] = item < McK::nb_items(a);
            if (item >= McK::nb_items(a)) return;
            const McCmdDev c = a.cmds[a.items[item].cmd];
            const int w = 1 << c.log2_w, h = 1 << c.log2_h, wc = w >> 1, hc = h >> 1;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int omx = r ? c.mv1_x : c.mv0_x, omy = r ? c.mv1_y : c.mv0_y;
                int mvx = omx, mvy = omy;
                McToolCommon::clip_mv(a.pic_w, a.pic_h, c.x0, c.y0, w, h, mvx, mvy);
                const int ridx = r ? c.ref_idx1 : c.ref_idx0;
                const int bxl = c.x0 + (mvx >> 4), byl = c.y0 + (mvy >> 4);
                const uint16_t *pl = a.refs.plane[r][ridx][0];
                const int pitch = a.refs.pitch[r][ridx][0];
                uint32_t *pv = pw + r * PWR * PWW;
                for (int i = lane; i < (h + 11) * PWW; i += 32) {
                    const int row = i / PWW, wd = i - row * PWW;
                    const int y = ovd_clamp(byl + ovd_clamp(row - 5, -3, h + 3), 0, a.pic_h - 1);
                    const int x0s = ovd_clamp(bxl + ovd_clamp(2 * wd - 5, -3, w + 3), 0, a.pic_w - 1);
                    const int x1s = ovd_clamp(bxl + ovd_clamp(2 * wd - 4, -3, w + 3), 0, a.pic_w - 1);
                    const uint16_t *rowp = pl + (size_t)y * pitch;
                    pv[i] = (uint32_t)rowp[x0s] | ((uint32_t)rowp[x1s] << 16);
                }
                const int bxc = (c.x0 >> 1) + (mvx >> 5), byc = (c.y0 >> 1) + (mvy >> 5);
                for (int i = lane; i < 2 * (hc + 7) * PCW; i += 32) {
                    const int nc1 = (hc + 7) * PCW, comp = i >= nc1, k = i - comp * nc1, row = k >> 3, wd = k & 7;          // PCW == 8: no division
                    const uint16_t *plc = a.refs.plane[r][ridx][1 + comp];
                    const int pitch_c = a.refs.pitch[r][ridx][1 + comp];
                    const int y = ovd_clamp(byc + ovd_clamp(row - 3, -1, hc + 1), 0, ch - 1);
                    const int x0s = ovd_clamp(bxc + ovd_clamp(2 * wd - 3, -1, wc + 1), 0, cw - 1);
                    const int x1s = ovd_clamp(bxc + ovd_clamp(2 * wd - 2, -1, wc + 1), 0, cw - 1);
                    const uint16_t *rowp = plc + (size_t)y * pitch_c;
                    pwc[((r * 2 + comp) * PCR + row) * PCW + wd] = (uint32_t)rowp[x0s] | ((uint32_t)rowp[x1s] << 16);
                }
                if (lane == 0) { sit[8 + 2 * r] = omx; sit[9 + 2 * r] = omy; }
            }
            if (lane == 0) { sit[1] = c.x0; sit[2] = c.y0; sit[3] = w | (h << 8); sit[4] = (c.flags & MC_HPEL_AMVR) != 0; sit[5] = c.aux & 1; }
            return;
        }
        if (!sit[0]) return;
        const int ibx = sit[1], iby = sit[2], iw = sit[3] & 0xff, ih = sit[3] >> 8;
        const uint16_t *pw16 = (const uint16_t *)pw;

        if (P == 1) {
            // bilinear samples at window positions (j - 2, i - 2); sample (x, y) of the padded window sits at column x + 5
            const int n = (ih + 4) * (iw + 4);
            for (int o = lane; o < 2 * n; o += 32) {
                // iw + 4 is 12 or 20: a warp-uniform choice between two constant divisors instead of a run-time division
                const int r = o >= n, k = o - r * n, i = iw == 16 ? k / 20 : k / 12, j = k - i * (iw + 4);
                const int fx = sit[8 + 2 * r] & 15, fy = sit[9 + 2 * r] & 15;
                const uint16_t *s = pw16 + (r * PWR + i + 3) * (2 * PWW) + j + 3;
                int v;
                if (!fx && !fy) v = s[0] << (10 - a.bitdepth);
                else if (!fy) v = ((16 - fx) * s[0] + fx * s[1] + 8) >> 4;
                else if (!fx) v = ((16 - fy) * s[0] + fy * s[2 * PWW] + 8) >> 4;
                else {
                    const int t0 = ((16 - fx) * s[0] + fx * s[1] + 8) >> 4, t1 = ((16 - fx) * s[2 * PWW] + fx * s[2 * PWW + 1] + 8) >> 4;
                    v = ((16 - fy) * t0 + fy * t1 + 8) >> 4;
                }
                bl[(r * BLS + i) * BLS + j] = (uint16_t)v;
            }
            return;
        }
        if (P == 2) {
            if (lane < 25) {
                const int dx = lane % 5 - 2, dy = lane / 5 - 2;
                int s = 0;
                for (int i = 0; i < ih; i += 2) {
                    const uint16_t *r0 = bl + (i + 2 + dy) * BLS + 2 + dx, *r1 = bl + (BLS + i + 2 - dy) * BLS + 2 - dx;
                    for (int j = 0; j < iw; j++) s = ovd_sad((int)(int16_t)r0[j], (int)(int16_t)r1[j], s);
                }
                sadv[lane] = s;
            }
            return;
        }
        if (P == 3) {
            if (lane != 0) return;
            int mv[4] = { sit[8], sit[9], sit[10], sit[11] };
            int min_cost = sadv[12] - (sadv[12] >> 2);
            if (min_cost >= iw * ih) {
                sadv[12] = min_cost;
                int best = 12;
                unsigned mc = 0xffffffffu;
                for (int idx = 0; idx < 25; idx++)
                    if ((unsigned)sadv[idx] < mc || (idx == 12 && (unsigned)sadv[idx] <= mc)) { mc = (unsigned)sadv[idx]; best = idx; }
                int dh = (best % 5 - 2) << 4, dv = (best / 5 - 2) << 4;
                min_cost = sadv[best];
                if (dh != 32 && dh != -32 && dv != 32 && dv != -32) {
                    const int *sb = sadv + best;
                    const int den_h = sb[-1] + sb[1] - (sb[0] << 1), den_v = sb[-5] + sb[5] - (sb[0] << 1);
                    if (den_h) {
                        if (sb[-1] != sb[0] && sb[1] != sb[0]) dh += div_for_maxq7((sb[-1] << 4) - (sb[1] << 4), den_h);
                        else dh += sb[-1] == sb[0] ? -8 : 8;
                    }
                    if (den_v) {
                        if (sb[-5] != sb[0] && sb[5] != sb[0]) dv += div_for_maxq7((sb[-5] - sb[5]) << 4, den_v);
                        else dv += sb[-5] == sb[0] ? -8 : 8;
                    }
                }
                mv[0] = ovd_clamp(mv[0] + dh, -(1 << 17), (1 << 17) - 1); mv[1] = ovd_clamp(mv[1] + dv, -(1 << 17), (1 << 17) - 1);
                mv[2] = ovd_clamp(mv[2] - dh, -(1 << 17), (1 << 17) - 1); mv[3] = ovd_clamp(mv[3] - dv, -(1 << 17), (1 << 17) - 1);
            }
            DmvrMvDev o = { mv[0], mv[1], mv[2], mv[3] };
            aa.out[item] = o;
            sit[6] = sit[5] ? (min_cost < 2 * iw * ih) : 1;                    // disable_bdof
            const int hpel = sit[4];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int nx = mv[2 * r], ny = mv[2 * r + 1], ox = sit[8 + 2 * r], oy = sit[9 + 2 * r];
                sit[12 + 6 * r] = (nx >> 4) - (ox >> 4); sit[13 + 6 * r] = (ny >> 4) - (oy >> 4);
                sit[14 + 6 * r] = McToolCommon::filter_row(nx & 15, hpel, 0) | (McToolCommon::filter_row(ny & 15, hpel, 0) << 8)
                                  | ((((nx & 15) >> 3) != 0) << 16) | ((((ny & 15) >> 3) != 0) << 17);
                sit[15 + 6 * r] = (nx >> 5) - (ox >> 5); sit[16 + 6 * r] = (ny >> 5) - (oy >> 5);
                sit[17 + 6 * r] = (nx & 31) | ((ny & 31) << 8);
            }
            return;
        }
        const int disable_bdof = sit[6];
        if (P == 4) {
            const int r = lane >> 4, sh1 = a.bitdepth - 8;
            {
                const int x = lane & 15;
                if (x < iw) {
                    const int c0 = x + sit[12 + 6 * r] + 2, row0 = sit[13 + 6 * r] + 2, frl = sit[14 + 6 * r];
                    uint32_t Bh[5], Bv0[4], Bv1[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) { Bh[j] = mc_cw_l[frl & 0xff][c0 & 1][j]; Bv1[j] = mc_cw_l[(frl >> 8) & 0xff][1][j]; }
#pragma unroll
                    for (int j = 0; j < 4; j++) Bv0[j] = mc_cw_l[(frl >> 8) & 0xff][0][j];
                    const uint32_t *wv = pw + (r * PWR + row0) * PWW + (c0 >> 1);
                    if (ih == 16) McK::column<16, 8, 5>(wv, PWW, Bh, Bv0, Bv1, sh1, ex + r * 256 + x, 16);
                    else          McK::column<8, 8, 5>(wv, PWW, Bh, Bv0, Bv1, sh1, ex + r * 256 + x, 16);
                }
            }
            {
                const int x = lane & 7, comp = (lane >> 3) & 1;
                if (x < (iw >> 1)) {
                    const int c0 = x + sit[15 + 6 * r] + 2, row0 = sit[16 + 6 * r] + 2, frc = sit[17 + 6 * r];
                    uint32_t Bh[3], Bv0[2], Bv1[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) { Bh[j] = mc_cw_c[frc & 0xff][c0 & 1][j]; Bv1[j] = mc_cw_c[frc >> 8][1][j]; }
#pragma unroll
                    for (int j = 0; j < 2; j++) Bv0[j] = mc_cw_c[frc >> 8][0][j];
                    const uint32_t *wv = pwc + ((r * 2 + comp) * PCR + row0) * PCW + (c0 >> 1);
                    int16_t *e = exc + ((r * 2 + comp) * 8) * 8 + x;
                    if (ih == 16) McK::column<8, 4, 3>(wv, PCW, Bh, Bv0, Bv1, sh1, e, 8);
                    else          McK::column<4, 4, 3>(wv, PCW, Bh, Bv0, Bv1, sh1, e, 8);
                }
            }
            return;
        }
        BdofCore::Ring ring[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int frl = sit[14 + 6 * r];
            ring[r].win = pw + r * PWR * PWW; ring[r].ww = PWW;
            ring[r].col0 = 5 + sit[12 + 6 * r] + ((frl >> 16) & 1); ring[r].row0 = 5 + sit[13 + 6 * r] + ((frl >> 17) & 1);
        }
        if (P == 5) {
            const int maxv = (1 << a.bitdepth) - 1;
            {   // chroma average: lanes 0-15 = (component, row)
                const int comp = (lane >> 3) & 1, y = lane & 7;
                if (lane < 16 && y < (ih >> 1)) {
                    uint16_t *d = (comp ? a.dst[2] : a.dst[1]) + (size_t)((iby >> 1) + y) * (comp ? a.dst_pitch[2] : a.dst_pitch[1]) + (ibx >> 1);
                    McK::store_row8(a, 1, exc + (comp * 8 + y) * 8, exc + ((2 + comp) * 8 + y) * 8, d, iw >> 1, 2, 0, 0, maxv, false);
                }
            }
            if (disable_bdof) {
                const int y = lane >> 1, x = (lane & 1) * 8;
                if (y < ih && x < iw)
                    McK::store_row8(a, 0, ex + y * 16 + x, ex + (16 + y) * 16 + x, a.dst[0] + (size_t)(iby + y) * a.dst_pitch[0] + ibx + x,
                                    iw - x, 2, 0, 0, maxv, a.lmcs_fwd != 0);
            } else BdofCore::grad(ex, G, iw, ih, ring, 14 - a.bitdepth, lane);
            return;
        }
        if (disable_bdof) return;
        if (P == 6) { BdofCore::partial(ex, G, part, iw, ih, lane); return; }
        if (P == 7) BdofCore::apply(a, ex, G, part, iw, ih, ibx, iby, lane);
    }
};
