// mc_strip_kernels.h -- translational motion compensation, one THREAD per 4-sample-wide strip (sm_100a; also compiled by
// tests/emu, see dev.h).  Replaces the warp-per-block McK / McSmallCK launches for plain and GPM commands.
//
// Reference behaviour: as mc_kernels.h (rcn_mcp_b & co, libovvc/rcn_inter.c:2750-2966; interpolation libovvc/rcn_mc.c:381-1608;
// GPM blend rcn_inter.c:3080-3115; forward LMCS of the luma prediction rcn_inter.c:599).
//
// Why threads, not warps: a VVC picture is mostly SMALL prediction blocks (4x4 affine / 8x8 SbTMVP sub-blocks, 8x8 .. 16x16
// coding units).  A warp per block spends its instructions on staging, barriers and addressing; a thread that owns a 4 x H strip
// of ONE component keeps everything in registers and never synchronises:
//   * it walks its reference rows top to bottom: 7 (luma) / 5 (chroma) 32-bit loads per row straight from L1/L2 -- neighbouring
//     threads own neighbouring strips, so a warp's loads of one row are one or a few contiguous segments; a 16-bit funnel shift
//     puts the first tap on a word boundary, so the two tap packings of mc_tables.h (first tap on an even / odd sample) are
//     chosen statically per output: 4 outputs = 18 IDP.2A for 8 taps;
//   * the 14-bit horizontal results are kept as VERTICAL pairs (row 2j, row 2j+1) per column in a ring of 5 (3) registers per
//     column; an even output row is 4 IDP.2A on the even-phase packing, an odd row 5 IDP.2A on the shifted packing over the
//     same pairs -- no transposition, no shared-memory exchange;
//   * bi-prediction / GPM: the first list's strip goes to the thread's private shared-memory slot (no barrier: nobody else
//     reads it), the second list's rows are combined with it as they are produced, clipped, forward-LMCS'ed (luma) and stored
//     as one 8-byte write per row.
// Strips are bucketed by height on the host (mc_host.h) so the threads of a warp run the same trip count.
#pragma once
#include "mc_kernels.h"

struct McStripDev {              // one 4 x h strip of component comp of command cmd; ox / oy in samples of that component
    uint32_t cmd;
    uint8_t ox, oy, h, comp;
};

struct McStripArgs {
    McK::Args mc;                    // items / nb_items unused
    const McStripDev *strips;
    int nb_strips;
};

// One kernel per (taps, strip height) class -- NT = 8 luma / 4 chroma, H = 16 / 8 / 4 / 2 rows: each launch runs one fully
// unrolled loop and gets its own register allocation; the host hands every class its contiguous range of the bucketed strip list.
template <int NT, int H>
struct McStripK {
    static const int THREADS = 128, MIN_BLOCKS = (H >= 16 ? 5 : 6), PHASES = 1;
    static const int MAX_H = H;
    static const int L1_HEAVY = 40;      // per cent of the L1 / shared-memory array given to shared memory: the reference rows live in L1
    static const int SMEM_BYTES = THREADS * MAX_H * 2 * 4;       // [row][2 words][thread]

    typedef McStripArgs Args;
    struct State { int dummy; };

    OVD uint32_t pack2(int a, int b) { return (uint32_t)(uint16_t)(int16_t)a | ((uint32_t)b << 16); }

    // one reference row of the strip: 4 horizontal results (14 bit), first tap at sample xs of row y
    OVD void hrow(const uint16_t *plane, int pitch, int pw, int ph, int xw, int par, int y, bool fast, const uint32_t *A, const uint32_t *B, int sh1, int t[4])
    {
        const int NA = NT / 2, NL = NT / 2 + 3;
        uint32_t w[NL];
        if (fast) {
            const uint32_t *rp = (const uint32_t *)(plane + (size_t)y * pitch + xw);
#pragma unroll
            for (int k = 0; k < NL; k++) w[k] = rp[k];
        } else {
#pragma unroll
            for (int k = 0; k < NL; k++) w[k] = McK::load_pair(plane, pitch, pw, ph, xw + 2 * k, y);
        }
        const int sh = par << 4;
#pragma unroll
        for (int k = 0; k < NL - 1; k++) w[k] = ovd_funnel_r(w[k], w[k + 1], sh);
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int j = 0; j < NA; j++) { a0 = ovd_dp2a_lo(w[j], A[j], a0); a2 = ovd_dp2a_lo(w[j + 1], A[j], a2); }
#pragma unroll
        for (int j = 0; j < NA + 1; j++) { a1 = ovd_dp2a_lo(w[j], B[j], a1); a3 = ovd_dp2a_lo(w[j + 1], B[j], a3); }
        t[0] = a0 >> sh1; t[1] = a1 >> sh1; t[2] = a2 >> sh1; t[3] = a3 >> sh1;
    }

    // One prediction direction of a 4 x H strip.  sink(y, v0..v3) receives the 14-bit results row by row.
    template <class Sink>
    OVD void run_dir(const uint16_t *plane, int pitch, int pw, int ph, int xs, int ys, const uint32_t *Ah, const uint32_t *Bh,
                     const uint32_t *Av, const uint32_t *Bv, int sh1, Sink &sink)
    {
        const int NA = NT / 2, NB = NT / 2 + 1, NR = H + NT - 1;
        const int par = xs & 1, xw = xs - par;
        // window inside the picture: plain loads; else clamped coordinates (emulate_block_border, rcn_inter.c:148-225)
        const bool fast = xw >= 0 && xw + 2 * (NT / 2 + 3) <= pw && ys >= 0 && ys + NR <= ph;
        uint32_t P[NB][4];
        int t0[4], t1[4];
#pragma unroll
        for (int j = 0; j < NA; j++) {
            hrow(plane, pitch, pw, ph, xw, par, ys + 2 * j, fast, Ah, Bh, sh1, t0);
            hrow(plane, pitch, pw, ph, xw, par, ys + 2 * j + 1, fast, Ah, Bh, sh1, t1);
#pragma unroll
            for (int c = 0; c < 4; c++) P[j][c] = pack2(t0[c], t1[c]);
        }
#pragma unroll
        for (int m = 0; m < H / 2; m++) {
            const int r0 = 2 * m + NT, r1 = r0 + 1;
            if (r0 < NR) hrow(plane, pitch, pw, ph, xw, par, ys + r0, fast, Ah, Bh, sh1, t0); else { t0[0] = t0[1] = t0[2] = t0[3] = 0; }
            if (r1 < NR) hrow(plane, pitch, pw, ph, xw, par, ys + r1, fast, Ah, Bh, sh1, t1); else { t1[0] = t1[1] = t1[2] = t1[3] = 0; }
#pragma unroll
            for (int c = 0; c < 4; c++) P[NA][c] = pack2(t0[c], t1[c]);
            int e[4], o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int ae = 0, ao = 0;
#pragma unroll
                for (int j = 0; j < NA; j++) ae = ovd_dp2a_lo(P[j][c], Av[j], ae);
#pragma unroll
                for (int j = 0; j < NB; j++) ao = ovd_dp2a_lo(P[j][c], Bv[j], ao);
                e[c] = (int16_t)(ae >> 6); o[c] = (int16_t)(ao >> 6);
            }
            sink(2 * m, e);
            sink(2 * m + 1, o);
#pragma unroll
            for (int j = 0; j < NA; j++)
#pragma unroll
                for (int c = 0; c < 4; c++) P[j][c] = P[j + 1][c];
        }
    }

    struct ToPicture {                   // last direction: combine (rcn_mc.c put_*_uni / put_*_bi / put_weighted_*), clip, LMCS, store
        uint32_t *slot;                  // the thread's private shared-memory slot
        int first, bi;                   // first: this is the first of two directions -> keep it in the slot; bi: combine with the slot
        uint16_t *dst; int pitch, n;     // n = samples to store (4, or 2 for a two-sample-wide chroma block)
        int bd, maxv, bcw_w1, gpm, gx, gy, gstep, lw, lh;
        const uint16_t *lut;
        OVDM void operator()(int y, const int v[4]) const
        {
            int r[4];
            if (first) {                 // no synchronisation: nobody else reads the slot
                slot[(2 * y) * THREADS] = (uint32_t)(uint16_t)(int16_t)v[0] | ((uint32_t)v[1] << 16);
                slot[(2 * y + 1) * THREADS] = (uint32_t)(uint16_t)(int16_t)v[2] | ((uint32_t)v[3] << 16);
                return;
            }
            if (!bi) {
#pragma unroll
                for (int c = 0; c < 4; c++) r[c] = (v[c] + (1 << (13 - bd))) >> (14 - bd);
            } else {
                const uint32_t q0 = slot[(2 * y) * THREADS], q1 = slot[(2 * y + 1) * THREADS];
                const int p0[4] = { (int16_t)(q0 & 0xffff), (int16_t)(q0 >> 16), (int16_t)(q1 & 0xffff), (int16_t)(q1 >> 16) };
                if (gpm) {
                    const GpmGeom gg = gpm_geom(gpm - 1, lw, lh);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int w = gpm_weight(gg, gx + c * gstep, gy + y * gstep);
                        r[c] = (v[c] * (8 - w) + p0[c] * w + (1 << (16 - bd))) >> (17 - bd);
                    }
                } else if (bcw_w1 == 4) {
#pragma unroll
                    for (int c = 0; c < 4; c++) r[c] = (p0[c] + v[c] + (1 << (14 - bd))) >> (15 - bd);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) r[c] = (v[c] * bcw_w1 + p0[c] * (8 - bcw_w1) + (1 << (16 - bd))) >> (17 - bd);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) { r[c] = ovd_clamp(r[c], 0, maxv); if (lut) r[c] = lut[r[c]]; }
            uint16_t *d = dst + (size_t)y * pitch;
            if (n == 4) { ovd_u2 o; o.x = (uint32_t)r[0] | ((uint32_t)r[1] << 16); o.y = (uint32_t)r[2] | ((uint32_t)r[3] << 16); ovd_st8(d, o); }
            else *(uint32_t *)d = (uint32_t)r[0] | ((uint32_t)r[1] << 16);
        }
    };

    OVD void strip(const Args &a, const McK::Item &it, const McStripDev &s, const McCmdDev &c, uint32_t *slot)
    {
        const McK::Args &m = a.mc;
        const int chroma = s.comp != 0;
        const int pw = chroma ? m.pic_w >> 1 : m.pic_w, ph = chroma ? m.pic_h >> 1 : m.pic_h;
        const int bx = (chroma ? c.x0 >> 1 : c.x0) + s.ox, by = (chroma ? c.y0 >> 1 : c.y0) + s.oy;
        const int sh1 = m.bitdepth - 8;
        const int bw = (1 << c.log2_w) >> chroma;
        ToPicture out;
        out.slot = slot; out.bi = it.nref == 2;
        out.dst = m.dst[s.comp] + (size_t)by * m.dst_pitch[s.comp] + bx; out.pitch = m.dst_pitch[s.comp];
        out.n = bw - s.ox < 4 ? bw - s.ox : 4;
        out.bd = m.bitdepth; out.maxv = (1 << m.bitdepth) - 1;
        out.bcw_w1 = it.bcw == 0 ? 4 : (it.bcw == 1 ? -2 : it.bcw == 2 ? 3 : it.bcw == 4 ? 5 : 10);      // rcn_inter.c:89
        out.gpm = it.gpm ? (it.gpm & 0xff) : 0; out.lw = c.log2_w; out.lh = c.log2_h;
        out.gstep = chroma ? 2 : 1; out.gx = s.ox * out.gstep; out.gy = s.oy * out.gstep;
        out.lut = chroma ? NULL : m.lmcs_fwd;
#pragma unroll 1
        for (int r = 0; r < it.nref; r++) {
            const int list = r ? it.list[1] : it.list[0], ridx = r ? it.ridx[1] : it.ridx[0];      // (no dynamic register indexing)
            const int mvx = r ? it.mvx[1] : it.mvx[0], mvy = r ? it.mvy[1] : it.mvy[0];
            const uint16_t *plane = m.refs.plane[list][ridx][s.comp];
            const int pitch = m.refs.pitch[list][ridx][s.comp];
            uint32_t Ah[NT / 2], Bh[NT / 2 + 1], Av[NT / 2], Bv[NT / 2 + 1];
            int xs, ys;
            if (NT == 8) {
                int px = mvx & 15, py = mvy & 15;
                if ((it.flags & MC_HPEL_AMVR) && !it.pu_small) { if (px == 8) px = 16; if (py == 8) py = 16; }       // rcn_inter.c:572-577
                if (it.pu_small) { px += 17; py += 17; }                                                               // 6-tap set, rcn_mc.c:457
#pragma unroll
                for (int j = 0; j < 4; j++) { Ah[j] = mc_cw_l[px][0][j]; Av[j] = mc_cw_l[py][0][j]; }
#pragma unroll
                for (int j = 0; j < 5; j++) { Bh[j] = mc_cw_l[px][1][j]; Bv[j] = mc_cw_l[py][1][j]; }
                xs = bx + (mvx >> 4) - 3; ys = by + (mvy >> 4) - 3;
            } else {
                const int px = mvx & 31, py = mvy & 31;
#pragma unroll
                for (int j = 0; j < 2; j++) { Ah[j] = mc_cw_c[px][0][j]; Av[j] = mc_cw_c[py][0][j]; }
#pragma unroll
                for (int j = 0; j < 3; j++) { Bh[j] = mc_cw_c[px][1][j]; Bv[j] = mc_cw_c[py][1][j]; }
                xs = bx + (mvx >> 5) - 1; ys = by + (mvy >> 5) - 1;
            }
            out.first = r + 1 < it.nref;
            run_dir(plane, pitch, pw, ph, xs, ys, Ah, Bh, Av, Bv, sh1, out);
        }
    }

    // the command's motion data (McK::load_item without the item geometry)
    OVD McK::Item setup(const Args &a, const McStripDev &s, const McCmdDev &c)
    {
        McK::Item it;
        const McK::Args &m = a.mc;
        const int pu_w = 1 << c.log2_w, pu_h = 1 << c.log2_h;
        it.valid = 1; it.bx = it.by = it.w = it.h = 0;
        it.pu_small = (pu_w == 4 && pu_h == 4);
        it.flags = c.flags;
        const int identical = c.inter_dir == 3 && (uint16_t)m.refs.poc[0][c.ref_idx0] == (uint16_t)m.refs.poc[1][c.ref_idx1] &&
                              c.mv0_x == c.mv1_x && c.mv0_y == c.mv1_y;                       // check_identical_motion, rcn_inter.c:256-267
        const int gpm = c.tool == 4;
        const int bi = (c.inter_dir == 3 && !identical) || gpm;
        it.gpm = gpm ? (int)(1 + (c.aux & 63)) : 0;
        if (gpm) {
            it.nref = 2;
            it.list[0] = (c.aux >> 8) & 1; it.ridx[0] = c.ref_idx0; it.mvx[0] = c.mv0_x; it.mvy[0] = c.mv0_y;
            it.list[1] = (c.aux >> 9) & 1; it.ridx[1] = c.ref_idx1; it.mvx[1] = c.mv1_x; it.mvy[1] = c.mv1_y;
        } else if (bi) {
            it.nref = 2;
            it.list[0] = 0; it.ridx[0] = c.ref_idx0; it.mvx[0] = c.mv0_x; it.mvy[0] = c.mv0_y;
            it.list[1] = 1; it.ridx[1] = c.ref_idx1; it.mvx[1] = c.mv1_x; it.mvy[1] = c.mv1_y;
        } else {
            const int l1 = (c.inter_dir & 2) || identical;
            it.nref = 1;
            it.list[0] = l1; it.ridx[0] = l1 ? c.ref_idx1 : c.ref_idx0;
            it.mvx[0] = l1 ? c.mv1_x : c.mv0_x; it.mvy[0] = l1 ? c.mv1_y : c.mv0_y;
            it.list[1] = 0; it.ridx[1] = 0; it.mvx[1] = 0; it.mvy[1] = 0;
        }
        it.bcw = (gpm || c.bcw_idx_plus1 == 0 || c.bcw_idx_plus1 == 3) ? 0 : c.bcw_idx_plus1;
        const int x_max = (m.pic_w + 2 - c.x0) << 4, y_max = (m.pic_h + 2 - c.y0) << 4;            // clip_mv, rcn_inter.c:96-108
        const int x_min = -((pu_w + 3 + c.x0) << 4), y_min = -((pu_h + 3 + c.y0) << 4);
#pragma unroll
        for (int r = 0; r < 2; r++) { it.mvx[r] = ovd_clamp(it.mvx[r], x_min, x_max); it.mvy[r] = ovd_clamp(it.mvy[r], y_min, y_max); }
        (void)s;
        return it;
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int i = bx * THREADS + tid;
        if (i >= a.nb_strips) return;
        const McStripDev s = a.strips[i];
        const McCmdDev c = a.mc.cmds[s.cmd];
        const McK::Item it = setup(a, s, c);
        uint32_t *slot = (uint32_t *)smem + tid;
        strip(a, it, s, c, slot);
    }
};
