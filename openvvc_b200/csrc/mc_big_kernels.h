// mc_big_kernels.h -- translational motion compensation of LARGE prediction units (width >= 32, height >= 16): most of
// the area of a picture (sm_100a; also compiled by tests/emu, see dev.h).  Same reference behaviour as mc_kernels.h
// (rcn_mcp_b & co, libovvc/rcn_inter.c:2750-2966; interpolation libovvc/rcn_mc.c:381-1608); GPM stays in McK.
//
// One warp owns a 32 x 32 (or 32 x 16) luma tile of one PU and its two 16 x 16 (16 x 8) chroma tiles:
//   phase 0  the reference windows of both lists are copied global -> shared with 16-byte cp.async (LDGSTS), no registers
//            involved; windows that cross the picture border take clamped per-pair loads instead;
//   phase 1  cp.async.wait_all (the CTA barrier that follows publishes the windows to the warp);
//   phase 2  every lane owns TWO adjacent output columns of a 16-row (8-row) group: each window word it loads feeds both
//            columns (IDP.2A.LO with the first column's taps, IDP.2A.HI with the second's, packed in one tap word), the
//            14-bit intermediates stay packed in registers for the vertical filter, list 0's predictions wait in
//            registers while list 1 is filtered, and the combined (average / BCW) clipped samples of the two columns
//            leave as one 32-bit store per row -- no exchange buffer, no separate store phase.
// Against McK's 16x16 items this amortises the window halo ((32+7)x(32+7) for 1024 samples instead of 4 x 23x23), halves
// the shared-memory loads per output and drops the shared-memory round trip before the store.
//
// Shared-memory rows are swizzled so that the two row groups (16 or 8 rows apart) of a warp read disjoint banks:
// luma row r starts at word r*24 + (r >> log2(RH))*16, chroma row r of a component at r*16 + (r >> log2(RHC))*8 with the
// two components 336 words (= 16 mod 32) apart.
#pragma once
#include "mc_kernels.h"

#ifndef MCBIG_MIN_BLOCKS
#define MCBIG_MIN_BLOCKS 4
#endif
struct McBigK {
    static const int THREADS = 128, MIN_BLOCKS = MCBIG_MIN_BLOCKS, PHASES = 3;
    static const int WARPS = 4;
    static const int WY_STRIDE = 24, WY_WORDS = 39 * 24 + 2 * 16;       // 48 samples x (32 + 7) rows per list
    static const int WC_STRIDE = 16, WC_COMP = 336;                     // 32 samples x (16 + 3) rows per list and component
    static const int SM_ITEM = 32;
    static const int SM_WIN_Y = 2 * WY_WORDS, SM_WIN_C = 2 * 2 * WC_COMP;
    static const int SM_WARP = (SM_ITEM + SM_WIN_Y + SM_WIN_C) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;
    typedef McK::Args Args;
    struct State { int dummy; };

    OVD int rb_y(int r, int lrh) { return r * WY_STRIDE + ((r >> lrh) << 4); }
    OVD int rb_c(int r, int lrhc) { return r * WC_STRIDE + ((r >> lrhc) << 3); }

    // Predictions of one list for the lane's two columns (A at window sample o, B at o + 1) and NOUT rows:
    //   w      first window word of the lane (row 0 of its row group, word o >> 1); row k is ROWS(k) words further
    //   TH[j]  horizontal taps of window word j: column A in the low half (IDP.2A.LO), column B in the high half (.HI)
    //   TV[j]  vertical taps: even output rows in the low half, odd output rows in the high half
    //   PA/PB  results, rows packed in pairs
    // NT = 8 (luma) or 4 (chroma); STRIDE / SWZ: row k of the group sits at k * STRIDE + (k >> LOG2(NOUT)) * SWZ.
    template <int NOUT, int LOG2_NOUT, int NT, int STRIDE, int SWZ>
    OVD void predict2(const uint32_t *w, const uint32_t *TH, const uint32_t *TV, int sh1, uint32_t *PA, uint32_t *PB)
    {
        const int NW = NT / 2 + 1, NR = NOUT + NT - 1, NP = (NR + 1) / 2;
        uint32_t TA[NP], TB[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
            const uint32_t *r0 = w + (2 * k) * STRIDE + (((2 * k) >> LOG2_NOUT) * SWZ);
#pragma unroll
            for (int j = 0; j < NW; j++) { const uint32_t v = r0[j]; a0 = ovd_dp2a_lo(v, TH[j], a0); b0 = ovd_dp2a_hi(v, TH[j], b0); }
            if (2 * k + 1 < NR) {
                const uint32_t *r1 = w + (2 * k + 1) * STRIDE + (((2 * k + 1) >> LOG2_NOUT) * SWZ);
#pragma unroll
                for (int j = 0; j < NW; j++) { const uint32_t v = r1[j]; a1 = ovd_dp2a_lo(v, TH[j], a1); b1 = ovd_dp2a_hi(v, TH[j], b1); }
            }
            TA[k] = ovd_byte_perm((uint32_t)(a0 >> sh1), (uint32_t)(a1 >> sh1), 0x5410);
            TB[k] = ovd_byte_perm((uint32_t)(b0 >> sh1), (uint32_t)(b1 >> sh1), 0x5410);
        }
#pragma unroll
        for (int y = 0; y < NOUT; y += 2) {
            const int k = y >> 1;
            int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
#pragma unroll
            for (int j = 0; j < NT / 2; j++) { a0 = ovd_dp2a_lo(TA[k + j], TV[j], a0); b0 = ovd_dp2a_lo(TB[k + j], TV[j], b0); }
#pragma unroll
            for (int j = 0; j < NT / 2 + 1; j++) { a1 = ovd_dp2a_hi(TA[k + j], TV[j], a1); b1 = ovd_dp2a_hi(TB[k + j], TV[j], b1); }
            PA[k] = ovd_byte_perm((uint32_t)(a0 >> 6), (uint32_t)(a1 >> 6), 0x5410);
            PB[k] = ovd_byte_perm((uint32_t)(b0 >> 6), (uint32_t)(b1 >> 6), 0x5410);
        }
    }

    // sign-extended halves of a packed pair: one PRMT (selector nibble 9 = sign of byte 1) / one shift
    OVD int lo16(uint32_t v) { return (int)ovd_prmt(v, 0, 0x9910); }
    OVD int hi16(uint32_t v) { return (int)v >> 16; }

    // combine + store NOUT rows of the lane's column pair; d = first row's address of the pair.  One branch-free formula for
    // uni-prediction, the plain average and BCW (rcn_mc.c put_*_uni / put_*_bi / put_weighted_*): (q*wq + p*wp + rnd) >> sh with
    // q = the first list's prediction (0 when there is one list).
    template <int NOUT, bool LUT>
    OVD void store2(const uint32_t *QA, const uint32_t *QB, const uint32_t *PA, const uint32_t *PB, int wq, int wp, int rnd, int sh, int maxv,
                    const uint16_t *lut, uint16_t *d, int pitch)
    {
#pragma unroll
        for (int k = 0; k < NOUT / 2; k++) {
            int a0 = ovd_clamp((lo16(QA[k]) * wq + lo16(PA[k]) * wp + rnd) >> sh, 0, maxv), b0 = ovd_clamp((lo16(QB[k]) * wq + lo16(PB[k]) * wp + rnd) >> sh, 0, maxv);
            int a1 = ovd_clamp((hi16(QA[k]) * wq + hi16(PA[k]) * wp + rnd) >> sh, 0, maxv), b1 = ovd_clamp((hi16(QB[k]) * wq + hi16(PB[k]) * wp + rnd) >> sh, 0, maxv);
            if (LUT) { a0 = lut[a0]; b0 = lut[b0]; a1 = lut[a1]; b1 = lut[b1]; }
            *(uint32_t *)(d + (size_t)(2 * k) * pitch) = (uint32_t)a0 | ((uint32_t)b0 << 16);
            *(uint32_t *)(d + (size_t)(2 * k + 1) * pitch) = (uint32_t)a1 | ((uint32_t)b1 << 16);
        }
    }

    // luma (NT = 8) or chroma (NT = 4) of the lane's column pair: the lists one after the other through the same code (the first
    // list's predictions move to Q), combination, stores
    template <int NOUT, int LOG2_NOUT, int NT, int STRIDE, int SWZ>
    OVD void pair(const uint32_t *win, int list_words, const int *sit, int chroma, int nref, int bcw, int bd,
                  const uint16_t *lut, uint16_t *d, int pitch, int col)
    {
        const int NW = NT / 2 + 1;
        uint32_t QA[NOUT / 2], QB[NOUT / 2], PA[NOUT / 2], PB[NOUT / 2];
#pragma unroll
        for (int k = 0; k < NOUT / 2; k++) QA[k] = QB[k] = PA[k] = PB[k] = 0;
#pragma unroll 1
        for (int r = 0; r < nref; r++) {
            if (r) {
#pragma unroll
                for (int k = 0; k < NOUT / 2; k++) { QA[k] = PA[k]; QB[k] = PB[k]; }
            }
            const int offs = chroma ? (sit[8 + 4 * r] >> 8) & 0xff : sit[8 + 4 * r] & 0xff;
            const int fr = chroma ? sit[10 + 4 * r] : sit[9 + 4 * r], fx = fr & 0xff, fy = fr >> 8;
            const int o = offs + col, par = o & 1;
            uint32_t TH[NW], TV[NW];
#pragma unroll
            for (int j = 0; j < NW; j++) {
                uint32_t ta, tb, v0, v1;
                if constexpr (NT == 8) {
                    // column A starts on sample o (parity par), column B on o + 1: on an even o, B uses the shifted packing of the
                    // same words; on an odd o, B starts one word later with the plain packing
                    ta = mc_cw_l[fx][par][j];
                    tb = par ? (j ? mc_cw_l[fx][0][j - 1] : 0u) : mc_cw_l[fx][1][j];
                    v0 = mc_cw_l[fy][0][j]; v1 = mc_cw_l[fy][1][j];
                } else {
                    ta = mc_cw_c[fx][par][j];
                    tb = par ? (j ? mc_cw_c[fx][0][j - 1] : 0u) : mc_cw_c[fx][1][j];
                    v0 = mc_cw_c[fy][0][j]; v1 = mc_cw_c[fy][1][j];
                }
                TH[j] = ovd_byte_perm(ta, tb, 0x5410);
                TV[j] = ovd_byte_perm(v0, v1, 0x5410);
            }
            predict2<NOUT, LOG2_NOUT, NT, STRIDE, SWZ>(win + r * list_words + (o >> 1), TH, TV, bd - 8, PA, PB);
        }
        // uni: (p + rnd) >> (14 - bd); average: (q + p + rnd) >> (15 - bd); BCW: (p * w1 + q * (8 - w1) + rnd) >> (17 - bd)
        const int w1 = bcw == 1 ? -2 : bcw == 2 ? 3 : bcw == 4 ? 5 : 10;           // rcn_inter.c:89
        const int sh = nref == 1 ? 14 - bd : (bcw ? 17 - bd : 15 - bd), wp = (nref == 2 && bcw) ? w1 : 1, wq = nref == 1 ? 0 : (bcw ? 8 - w1 : 1);
        if (lut) store2<NOUT, true>(QA, QB, PA, PB, wq, wp, 1 << (sh - 1), sh, (1 << bd) - 1, lut, d, pitch);
        else     store2<NOUT, false>(QA, QB, PA, PB, wq, wp, 1 << (sh - 1), sh, (1 << bd) - 1, lut, d, pitch);
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t *sm = (uint32_t *)(smem + warp * SM_WARP);
        int *sit = (int *)sm;
        uint32_t *win_y = sm + SM_ITEM;
        uint32_t *win_c = win_y + SM_WIN_Y;
        const int cw = a.pic_w >> 1, ch = a.pic_h >> 1;

        if (P == 0) {
            const McK::Item it = McK::load_item(a, bx * WARPS + warp);
            if (lane == 0) sit[0] = it.valid;
            if (!it.valid) return;
            const int do_y = it.flags & MC_LUMA, do_c = it.flags & MC_CHROMA;
            const int lrh = it.h == 32 ? 4 : 3;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (r >= it.nref) break;
                const int xs = it.bx + (it.mvx[r] >> 4) - 3, ys = it.by + (it.mvy[r] >> 4) - 3, x_al = xs & ~7;
                const int xc = (it.bx >> 1) + (it.mvx[r] >> 5) - 1, yc = (it.by >> 1) + (it.mvy[r] >> 5) - 1, xc_al = xc & ~7;
                if (lane == 0) {
                    int px = it.mvx[r] & 15, py = it.mvy[r] & 15;
                    if (it.flags & MC_HPEL_AMVR) { if (px == 8) px = 16; if (py == 8) py = 16; }                  // rcn_inter.c:572-577
                    sit[8 + 4 * r] = (xs - x_al) | ((xc - xc_al) << 8);
                    sit[9 + 4 * r] = px | (py << 8);
                    sit[10 + 4 * r] = (it.mvx[r] & 31) | ((it.mvy[r] & 31) << 8);
                }
                const uint16_t *pl = a.refs.plane[it.list[r]][it.ridx[r]][0];
                const int pitch = a.refs.pitch[it.list[r]][it.ridx[r]][0], nrows = it.h + 7;
                uint32_t *w = win_y + r * WY_WORDS;
                if (do_y) {
                    if (x_al >= 0 && x_al + 2 * WY_STRIDE <= a.pic_w && ys >= 0 && ys + nrows <= a.pic_h) {
                        // 6 chunks of 16 bytes per row: lanes 0-29 take 5 rows per step
                        const int lr = (lane * 43) >> 8, c16 = lane - 6 * lr;            // lane / 6, lane % 6
                        const uint16_t *src = pl + (size_t)(ys + lr) * pitch + x_al + 8 * c16;
                        if (lane < 30)
                            for (int row = lr; row < nrows; row += 5, src += 5 * pitch) ovd_cp_async16(w + rb_y(row, lrh) + 4 * c16, src);
                    } else {                                                         // window crosses the picture border
                        for (int i = lane; i < nrows * WY_STRIDE; i += 32) {
                            const int row = i / WY_STRIDE, wd = i - row * WY_STRIDE;
                            w[rb_y(row, lrh) + wd] = McK::load_pair(pl, pitch, a.pic_w, a.pic_h, x_al + 2 * wd, ys + row);
                        }
                    }
                }
                if (do_c) {
                    const int nrows_c = (it.h >> 1) + 3, lrhc = lrh - 1;
                    const int fast_c = xc_al >= 0 && xc_al + 2 * WC_STRIDE <= cw && yc >= 0 && yc + nrows_c <= ch;
                    // 4 chunks of 16 bytes per row: the half-warps take one component each, 4 rows per step
                    const int comp = lane >> 4, lr = (lane >> 2) & 3, c16 = lane & 3;
                    const uint16_t *plc = a.refs.plane[it.list[r]][it.ridx[r]][1 + comp];
                    const int pitch_c = a.refs.pitch[it.list[r]][it.ridx[r]][1 + comp];
                    uint32_t *wc = win_c + (r * 2 + comp) * WC_COMP + 4 * c16;
                    if (fast_c) {
                        const uint16_t *src = plc + (size_t)(yc + lr) * pitch_c + xc_al + 8 * c16;
                        for (int row = lr; row < nrows_c; row += 4, src += 4 * pitch_c) ovd_cp_async16(wc + rb_c(row, lrhc), src);
                    } else {
                        for (int row = lr; row < nrows_c; row += 4) {
#pragma unroll
                            for (int q = 0; q < 4; q++) wc[rb_c(row, lrhc) + q] = McK::load_pair(plc, pitch_c, cw, ch, xc_al + 8 * c16 + 2 * q, yc + row);
                        }
                    }
                }
            }
            if (lane == 0) { sit[1] = it.bx; sit[2] = it.by; sit[3] = it.w; sit[4] = it.h; sit[5] = it.nref; sit[6] = it.flags; sit[7] = it.bcw; }
            return;
        }
        if (P == 1) { ovd_cp_async_wait_all(); return; }

        if (!sit[0]) return;
        const int ibx = sit[1], iby = sit[2], ih = sit[4], nref = sit[5], flags = sit[6], bcw = sit[7];
        if (flags & MC_LUMA) {
            const int cp = lane & 15, g = lane >> 4;
            if (ih == 32) {
                uint16_t *d = a.dst[0] + (size_t)(iby + 16 * g) * a.dst_pitch[0] + ibx + 2 * cp;
                pair<16, 4, 8, WY_STRIDE, 16>(win_y + g * (16 * WY_STRIDE + 16), WY_WORDS, sit, 0, nref, bcw, a.bitdepth, a.lmcs_fwd, d, a.dst_pitch[0], 2 * cp);
            } else {
                uint16_t *d = a.dst[0] + (size_t)(iby + 8 * g) * a.dst_pitch[0] + ibx + 2 * cp;
                pair<8, 3, 8, WY_STRIDE, 16>(win_y + g * (8 * WY_STRIDE + 16), WY_WORDS, sit, 0, nref, bcw, a.bitdepth, a.lmcs_fwd, d, a.dst_pitch[0], 2 * cp);
            }
        }
        if (flags & MC_CHROMA) {
            const int cp = lane & 7, comp = (lane >> 3) & 1, g = lane >> 4;
            const uint32_t *w0 = win_c + comp * WC_COMP;
            uint16_t *dp = comp ? a.dst[2] : a.dst[1];
            const int dpitch = comp ? a.dst_pitch[2] : a.dst_pitch[1];
            if (ih == 32) {
                uint16_t *d = dp + (size_t)((iby >> 1) + 8 * g) * dpitch + (ibx >> 1) + 2 * cp;
                pair<8, 3, 4, WC_STRIDE, 8>(w0 + g * (8 * WC_STRIDE + 8), 2 * WC_COMP, sit, 1, nref, bcw, a.bitdepth, 0, d, dpitch, 2 * cp);
            } else {
                uint16_t *d = dp + (size_t)((iby >> 1) + 4 * g) * dpitch + (ibx >> 1) + 2 * cp;
                pair<4, 2, 4, WC_STRIDE, 8>(w0 + g * (4 * WC_STRIDE + 8), 2 * WC_COMP, sit, 1, nref, bcw, a.bitdepth, 0, d, dpitch, 2 * cp);
            }
        }
    }
};
