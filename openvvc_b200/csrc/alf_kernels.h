// alf_kernels.h -- whole-picture ALF / CC-ALF kernels (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: libovvc/rcn_alf.c (classification :283-738, luma 7x7 :1157-1272, chroma 5x5
// :961-1068, CC-ALF :806-872, per-CTU driver :1285-1433) working on the padded "filter region" of
// libovvc/rcn_ctu.c:316-510.  Here one CTA owns one tile of one CTU: it stages the tile + 3-sample halo of
// the pre-ALF picture in shared memory (coordinates clamped = the reference's replicated borders),
// classifies its 4x4 blocks and filters them, so the picture is read once and written once.
//
// Luma arithmetic runs on packed pairs of 16-bit samples: VIADDMNMX.S16x2 / VIMNMX.S16x2 do the clipped
// differences of two horizontally adjacent samples per instruction and IDP.2A does the two multiply-adds.
#pragma once
#include "dev.h"
#include "tile_grid.h"

// ---- device-side parameter formats -----------------------------------------------------------------------
#define ALF_TAB_ENTRY_WORDS 24          // per (transpose, class): 12 coef operands, 12 clips (the kernel negates the clip itself)
#define ALF_TAB_SET_WORDS   (100 * ALF_TAB_ENTRY_WORDS)

struct AlfCtuDev {                      // 8 bytes per CTU
    uint8_t flag, idx, cb_alt, cr_alt, cc_cb, cc_cr, pad0, pad1;
};

struct AlfChromaTab {                   // passed by value in the kernel arguments
    int16_t coeff[8][6];
    int16_t clip[8][6];
    int16_t cc[2][4][8];
};

OVD uint32_t alf_odd_pair(uint32_t w0, uint32_t w1) { return ovd_byte_perm(w0, w1, 0x5432); }

// =========================================================================================================
// Luma: classification + 7x7 diamond
// =========================================================================================================
template <bool WIDE>
struct AlfLumaK {
#ifndef ALF_LUMA_MIN_BLOCKS
#define ALF_LUMA_MIN_BLOCKS 3
#endif
    static const int THREADS = 256, MIN_BLOCKS = ALF_LUMA_MIN_BLOCKS, PHASES = 3;          // 80 registers: 3 CTAs / SM (was 128 -> 2)
    static const int TW = 128, TH = 32;
    static const int HALO_L = 8;                     // samples staged left of the tile: 16 bytes, the granularity of cp.async
    static const int PIX_STRIDE = TW + 16;           // samples; 8 left / right of the tile
    static const int PIX_ROWS = TH + 6;
    static const int CELL_COLS = TW / 2 + 2, CELL_ROWS = TH / 2 + 2;
    static const int SMEM_PIX = PIX_ROWS * PIX_STRIDE * 2;
    static const int SMEM_CELL = CELL_ROWS * CELL_COLS * 8;
    static const int SMEM_TAB = ALF_TAB_SET_WORDS * 4;
    static const int SMEM_BYTES = SMEM_PIX + SMEM_CELL + SMEM_TAB;

    struct Args {
        const uint16_t *src;
        uint16_t *dst;
        int src_pitch, dst_pitch;        // samples
        int pic_w, pic_h, log2_ctu, bitdepth;
        int tw, th;                      // tile size = min(TW/TH, ctu size)
        int tiles_x_per_ctu, tiles_y_per_ctu, nb_ctu_w;
        const uint32_t *tab;             // [24][100][24]
        const AlfCtuDev *ctu;
        const TileGridDev *tiles;        // device memory, or NULL (one tile): tile edges are padded like picture borders
    };
    struct State { int active; };

    struct Tile {
        int x0, y0, tw, th, ctu_y0, ctu_h, vb, enabled, set;
        int bx0, bx1, by0, by1;          // the rectangle samples are fetched from: the CTU's tile (the picture without tiles)
    };

    OVD Tile tile_of(const Args &a, int bx, int by)
    {
        Tile t;
        const int ctu_s = 1 << a.log2_ctu;
        // tiles per CTU are 1, 2 or 4 each way (CTU 32..128, tile 128 x 32): n >> 1 is log2(n); no integer division
        const int sx = a.tiles_x_per_ctu >> 1, sy = a.tiles_y_per_ctu >> 1;
        int cx = bx >> sx, cy = by >> sy;
        t.x0 = cx * ctu_s + (bx & (a.tiles_x_per_ctu - 1)) * a.tw;
        t.ctu_y0 = cy * ctu_s;
        t.y0 = t.ctu_y0 + (by & (a.tiles_y_per_ctu - 1)) * a.th;
        t.ctu_h = ovd_min(ctu_s, a.pic_h - t.ctu_y0);
        t.tw = ovd_min(a.tw, a.pic_w - t.x0);
        t.th = ovd_min(a.th, t.ctu_y0 + t.ctu_h - t.y0);
        // virtual boundary, CTU relative (rcn_alf.c:724 and :1353)
        t.vb = (t.ctu_h < ctu_s) ? a.pic_h : t.ctu_h - 4;
        t.enabled = 0; t.set = 0;
        t.bx0 = 0; t.bx1 = a.pic_w; t.by0 = 0; t.by1 = a.pic_h;
        if (t.tw > 0 && t.th > 0) {
            AlfCtuDev c = a.ctu[cy * a.nb_ctu_w + cx];
            t.enabled = (c.flag >> 2) & 1;
            t.set = c.idx;
            if (a.tiles) {
                int c0, c1;
                tile_span(a.tiles->col_bd, a.tiles->nb_cols, cx, a.nb_ctu_w, c0, c1);
                t.bx0 = c0 << a.log2_ctu; t.bx1 = ovd_min(c1 << a.log2_ctu, a.pic_w);
                tile_span(a.tiles->row_bd, a.tiles->nb_rows, cy, (a.pic_h + ctu_s - 1) >> a.log2_ctu, c0, c1);
                t.by0 = c0 << a.log2_ctu; t.by1 = ovd_min(c1 << a.log2_ctu, a.pic_h);
            }
        }
        return t;
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)bz;
        const Tile t = tile_of(a, bx, by);
        if (t.tw <= 0 || t.th <= 0) return;
        uint16_t *pix = (uint16_t *)smem;
        uint32_t *cell = (uint32_t *)(smem + SMEM_PIX);
        uint32_t *tab = (uint32_t *)(smem + SMEM_PIX + SMEM_CELL);
        if (P == 0) phase_load(a, t, pix, tab, tid);
        else if (!t.enabled) return;
#ifdef ALF_EXP_NO_CELLS              /* timing experiments (tools/build_variant.sh), never part of the product build */
        else if (P == 1) return;
#else
        else if (P == 1) phase_cells(a, t, pix, cell, tid);
#endif
#ifdef ALF_EXP_NO_FILTER
        else return;
#else
        else phase_filter(a, t, pix, cell, tab, tid);
#endif
        (void)st;
    }

    // ---- phase 0: stage tile + halo (or plain copy when luma ALF is off for this CTU) -----------------------
    OVD void phase_load(const Args &a, const Tile &t, uint16_t *pix, uint32_t *tab, int tid)
    {
        if (!t.enabled) {
            // dst = src for this tile: 16-byte chunks
            const int chunks = t.tw >> 3;
            for (int i = tid; i < chunks * t.th; i += THREADS) {
                int r = i / chunks, c = i - r * chunks;
                size_t so = (size_t)(t.y0 + r) * a.src_pitch + t.x0 + 8 * c;
                size_t d_o = (size_t)(t.y0 + r) * a.dst_pitch + t.x0 + 8 * c;
                ovd_st16(a.dst + d_o, ovd_ld16(a.src + so));
            }
            return;
        }
        const int nrows = t.th + 6;
        // tiles away from the picture's left / right border: the rows go to shared memory as 16-byte asynchronous copies (LDGSTS),
        // one lane per chunk and one warp per row -- no register staging, no per-chunk border tests (rows are clamped per row)
        const bool fast = t.x0 >= t.bx0 + HALO_L && t.x0 + t.tw + 8 <= t.bx1 && !(t.tw & 7) && !(a.src_pitch & 7);
        if (fast) {
            const int cpr = (t.tw + 16) >> 3, lane = tid & 31;        // <= 18 chunks per row
            if (lane < cpr)
                for (int r = tid >> 5; r < nrows; r += THREADS / 32) {
                    const int y = ovd_clamp(t.y0 - 3 + r, t.by0, t.by1 - 1);
                    ovd_cp_async16(pix + r * PIX_STRIDE + 8 * lane, a.src + (size_t)y * a.src_pitch + t.x0 - HALO_L + 8 * lane);
                }
        }
        const int nchunks = (t.tw + 16) >> 2;         // 4-sample (8-byte) chunks, <= 36
        const int ch = tid & 63;
        if (!fast && ch < nchunks) {
            int x = t.x0 - HALO_L + 4 * ch;
            for (int r = tid >> 6; r < nrows; r += 4) {
                int y = ovd_clamp(t.y0 - 3 + r, t.by0, t.by1 - 1);
                const uint16_t *row = a.src + (size_t)y * a.src_pitch;
                ovd_u2 v;
                if (x < t.bx0)          { uint32_t e = row[t.bx0];     v.x = v.y = e * 0x10001u; }
                else if (x >= t.bx1)    { uint32_t e = row[t.bx1 - 1]; v.x = v.y = e * 0x10001u; }
                else v = ovd_ld8(row + x);
                *(ovd_u2 *)(pix + r * PIX_STRIDE + 4 * ch) = v;
            }
        }
        // (reading each thread's table entry straight from global memory instead of staging the set was measured: no gain)
        const ovd_u4 *g = (const ovd_u4 *)(a.tab + (size_t)t.set * ALF_TAB_SET_WORDS);
        for (int i = tid; i < ALF_TAB_SET_WORDS / 4; i += THREADS) ((ovd_u4 *)tab)[i] = g[i];
        if (fast) ovd_cp_async_wait_all();
    }

    // ---- phase 1: 1-D Laplacians summed per 2x2 cell (rcn_alf.c:602-657 / :366-509) -------------------------
    OVD void phase_cells(const Args &a, const Tile &t, const uint16_t *pix, uint32_t *cell, int tid)
    {
        const int ncr = (t.th >> 1) + 2, ncc = (t.tw >> 1) + 2;
        const int vb_abs = t.ctu_y0 + t.vb;
        const uint32_t *pw = (const uint32_t *)pix;
        const int ws = PIX_STRIDE / 2;                 // words per row
        for (int idx = tid; idx < ncr * CELL_COLS; idx += THREADS) {
            int k = idx / CELL_COLS, m = idx - k * CELL_COLS;
            if (m >= ncc) continue;
            // cell rows: tile rows r0 = 2k-2, r0+1; shared-memory row index = tile row + 3
            int y_r0 = t.y0 + 2 * k - 2;
            const uint32_t *rb = pw + (2 * k + 1) * ws + m + (HALO_L - 4) / 2;
            const uint32_t *ra = (y_r0 == vb_abs) ? rb : rb - ws;              // "src0 = src1", rcn_alf.c:466
            const uint32_t *rc = rb + ws;
            const uint32_t *rd = (y_r0 + 1 == vb_abs - 1) ? rc : rc + ws;      // "src3 = src2", rcn_alf.c:464
            uint32_t a0 = ra[0], a1 = ra[1];
            uint32_t b0 = rb[0], b1 = rb[1], b2 = rb[2];
            uint32_t c0 = rc[0], c1 = rc[1], c2 = rc[2];
            uint32_t d1 = rd[1], d2 = rd[2];
            int A1 = a0 >> 16, A2 = a1 & 0xffff, A3 = a1 >> 16;
            int B1 = b0 >> 16, B2 = b1 & 0xffff, B3 = b1 >> 16, B4 = b2 & 0xffff;
            int C1 = c0 >> 16, C2 = c1 & 0xffff, C3 = c1 >> 16, C4 = c2 & 0xffff;
            int D2 = d1 & 0xffff, D3 = d1 >> 16, D4 = d2 & 0xffff;
            int p0 = B2 << 1, p1 = C3 << 1;
            int v  = ovd_sad(p0, A2 + C2, ovd_sad(p1, B3 + D3, 0));
            int h  = ovd_sad(p0, B1 + B3, ovd_sad(p1, C2 + C4, 0));
            int g0 = ovd_sad(p0, A1 + C3, ovd_sad(p1, B2 + D4, 0));
            int g1 = ovd_sad(p0, A3 + C1, ovd_sad(p1, B4 + D2, 0));
            cell[2 * idx]     = (uint32_t)v | ((uint32_t)h << 16);
            cell[2 * idx + 1] = (uint32_t)g0 | ((uint32_t)g1 << 16);
        }
    }

    // ---- phase 2: class of the thread's 4x4 block, then filter it -----------------------------------------------
    OVD void tap(int &acc0, int &acc1, uint32_t na, uint32_t nb, uint32_t neg_p, const uint32_t *e, int k)
    {
        const uint32_t ncl = ovd_vsub_16x2(0u, e[12 + k]);          // -clip, recomputed: 12 fewer live registers than a third table row
        uint32_t t1 = ovd_vmax_s16x2(ovd_viaddmin_s16x2(na, neg_p, e[12 + k]), ncl);
        uint32_t t2 = ovd_vmax_s16x2(ovd_viaddmin_s16x2(nb, neg_p, e[12 + k]), ncl);
        uint32_t s = ovd_vadd_16x2(t1, t2);
        if (WIDE) {
            int c = (int)e[k];
            acc0 += c * (int)(int16_t)(s & 0xffff);
            acc1 += c * ((int)s >> 16);
        } else {
            acc0 = ovd_dp2a_lo(s, e[k], acc0);
            acc1 = ovd_dp2a_hi(s, e[k], acc1);
        }
    }

    // one output row: 7 row windows of 6 words (12 samples, the thread's own at words 2,3)
    OVD ovd_u2 filter_row(const uint32_t *U3, const uint32_t *U2, const uint32_t *U1, const uint32_t *C,
                          const uint32_t *D1, const uint32_t *D2, const uint32_t *D3,
                          const uint32_t *e, int coarse, int maxv)
    {
        uint32_t out[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const uint32_t P = C[2 + p];
            const uint32_t neg_p = ovd_vsub_16x2(0u, P);
            int acc0 = 0, acc1 = 0;
            tap(acc0, acc1, D3[2 + p], U3[2 + p], neg_p, e, 0);
            tap(acc0, acc1, alf_odd_pair(D2[2 + p], D2[3 + p]), alf_odd_pair(U2[1 + p], U2[2 + p]), neg_p, e, 1);
            tap(acc0, acc1, D2[2 + p], U2[2 + p], neg_p, e, 2);
            tap(acc0, acc1, alf_odd_pair(D2[1 + p], D2[2 + p]), alf_odd_pair(U2[2 + p], U2[3 + p]), neg_p, e, 3);
            tap(acc0, acc1, D1[3 + p], U1[1 + p], neg_p, e, 4);
            tap(acc0, acc1, alf_odd_pair(D1[2 + p], D1[3 + p]), alf_odd_pair(U1[1 + p], U1[2 + p]), neg_p, e, 5);
            tap(acc0, acc1, D1[2 + p], U1[2 + p], neg_p, e, 6);
            tap(acc0, acc1, alf_odd_pair(D1[1 + p], D1[2 + p]), alf_odd_pair(U1[2 + p], U1[3 + p]), neg_p, e, 7);
            tap(acc0, acc1, D1[1 + p], U1[3 + p], neg_p, e, 8);
            tap(acc0, acc1, alf_odd_pair(C[3 + p], C[4 + p]), alf_odd_pair(C[0 + p], C[1 + p]), neg_p, e, 9);
            tap(acc0, acc1, C[3 + p], C[1 + p], neg_p, e, 10);
            tap(acc0, acc1, alf_odd_pair(C[2 + p], C[3 + p]), alf_odd_pair(C[1 + p], C[2 + p]), neg_p, e, 11);
            int r0 = coarse ? (acc0 + 512) >> 10 : (acc0 + 64) >> 7;
            int r1 = coarse ? (acc1 + 512) >> 10 : (acc1 + 64) >> 7;
            r0 = ovd_clamp(r0 + (int)(P & 0xffff), 0, maxv);
            r1 = ovd_clamp(r1 + (int)(P >> 16), 0, maxv);
            out[p] = (uint32_t)r0 | ((uint32_t)r1 << 16);
        }
        ovd_u2 o; o.x = out[0]; o.y = out[1];
        return o;
    }

    // alf_derive_filter_idx, rcn_alf.c:283-345
    OVD int derive_class(uint32_t sv, uint32_t sh, uint32_t sd, uint32_t sb, int shift, int near_vb)
    {
        const uint32_t th_lut = 0x43333333u, th_lut_lo = 0x32222210u;   // th[0..7] / th[8..15], 4 bits each
        uint32_t scale = near_vb ? 96u : 64u;
        uint32_t act = ovd_umin(((sh + sv) * scale) >> shift, 15u);
        int cls = (int)(((act < 8 ? th_lut_lo : th_lut) >> (4 * (act & 7))) & 0xf);
        uint32_t max_hv, min_hv, max_db, min_db, max_dir, min_dir;
        int dir_hv, dir_db, main_dir, sec_dir;
        if (sv > sh) { max_hv = sv; min_hv = sh; dir_hv = 1; } else { max_hv = sh; min_hv = sv; dir_hv = 3; }
        if (sd > sb) { max_db = sd; min_db = sb; dir_db = 0; } else { max_db = sb; min_db = sd; dir_db = 2; }
        if (max_db * min_hv > max_hv * min_db) { max_dir = max_db; min_dir = min_db; main_dir = dir_db; sec_dir = dir_hv; }
        else                                   { max_dir = max_hv; min_dir = min_hv; main_dir = dir_hv; sec_dir = dir_db; }
        if (max_dir * 2 > 9 * min_dir)   cls += (((main_dir & 1) << 1) + 2) * 5;
        else if (max_dir > 2 * min_dir)  cls += (((main_dir & 1) << 1) + 1) * 5;
        const uint32_t tr_lut = 0x31322010u;                            // tr_lut[0..7], 4 bits each
        int tr = (int)((tr_lut >> (4 * ((main_dir << 1) + (sec_dir >> 1)))) & 0xf);
        return tr * 25 + cls;
    }

    OVD void phase_filter(const Args &a, const Tile &t, const uint16_t *pix, const uint32_t *cell,
                          const uint32_t *tab, int tid)
    {
        const int bxi = tid & 31, byi = tid >> 5;
        if (4 * bxi >= t.tw || 4 * byi >= t.th) return;
        const int ctu_s = 1 << a.log2_ctu;
        const int maxv = (1 << a.bitdepth) - 1;

        // ---- classification of this 4x4 block
        const int y_rel = t.y0 + 4 * byi - t.ctu_y0;
        const int blk_y = y_rel & ~31;
        const int blk_h = ovd_min(32, t.ctu_h - blk_y);
        const int vb_blk = (blk_y + blk_h >= t.vb);                // rcn_alf.c:695
        int k0 = 0, k1 = 3, near_vb = 0;
        if (vb_blk && y_rel == t.vb - 4) { k1 = 2; near_vb = 1; }  // rcn_alf.c:517
        if (vb_blk && y_rel == t.vb)     { k0 = 1; near_vb = 1; }  // rcn_alf.c:548
        uint32_t svh = 0, sdd = 0;                                 // V|H<<16 and D0|D1<<16 as packed u16 pairs
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < k0 || k > k1) continue;
            const ovd_u4 *cp = (const ovd_u4 *)(cell + 2 * ((2 * byi + k) * CELL_COLS + 2 * bxi));
            ovd_u4 q0 = cp[0], q1 = cp[1];
            svh += q0.x + q0.z + q1.x + q1.z;       // each half <= 16 * 4092 < 65536: no carry between halves
            sdd += q0.y + q0.w + q1.y + q1.w;
        }
        const int entry = derive_class(svh & 0xffff, svh >> 16, sdd & 0xffff, sdd >> 16, a.bitdepth + 4, near_vb);

        // ---- coefficients of (transpose, class)
        uint32_t e[24];                                             // coefficient operands + clips (the -clip row of the table is not read)
        {
            const ovd_u4 *tp = (const ovd_u4 *)(tab + entry * ALF_TAB_ENTRY_WORDS);
#pragma unroll
            for (int i = 0; i < 24 / 4; i++) {
                ovd_u4 q = tp[i];
                e[4 * i] = q.x; e[4 * i + 1] = q.y; e[4 * i + 2] = q.z; e[4 * i + 3] = q.w;
            }
        }

        // ---- 10 rows x 12 samples around the block; row r is first needed by output row r - 6, so rows 7..9 are loaded
        // inside the row loop: seven rows (42 registers) are live at a time instead of ten
        uint32_t R[10][6];
        const ovd_u2 *pp = (const ovd_u2 *)(pix + (4 * byi) * PIX_STRIDE + 4 * bxi + (HALO_L - 4));
#define ALF_LOAD_ROW(r_)                                                                   \
        _Pragma("unroll")                                                                   \
        for (int q = 0; q < 3; q++) {                                                       \
            ovd_u2 v = pp[(r_) * (PIX_STRIDE / 4) + q];                                     \
            R[r_][2 * q] = v.x; R[r_][2 * q + 1] = v.y;                                     \
        }
#pragma unroll
        for (int r = 0; r < 7; r++) { ALF_LOAD_ROW(r) }

        uint16_t *drow = a.dst + (size_t)(t.y0 + 4 * byi) * a.dst_pitch + t.x0 + 4 * bxi;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j > 0) { ALF_LOAD_ROW(j + 6) }
            const int y_vb = (t.y0 + 4 * byi + j) & (ctu_s - 1);
            // rows near the virtual boundary do not reach across it (rcn_alf.c:1200-1221)
            int reach = 3;
            if (y_vb < t.vb && y_vb >= t.vb - 4)       reach = t.vb - 1 - y_vb;
            else if (y_vb >= t.vb && y_vb <= t.vb + 3) reach = y_vb - t.vb;
            const int coarse = (y_vb == t.vb - 1) || (y_vb == t.vb);
            ovd_u2 o;
            if (reach >= 3) {
                o = filter_row(R[j], R[j + 1], R[j + 2], R[j + 3], R[j + 4], R[j + 5], R[j + 6], e, coarse, maxv);
            } else {
                uint32_t U1[6], U2[6], U3[6], D1[6], D2[6], D3[6];
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    U1[q] = reach >= 1 ? R[j + 2][q] : R[j + 3][q];
                    D1[q] = reach >= 1 ? R[j + 4][q] : R[j + 3][q];
                    U2[q] = reach >= 2 ? R[j + 1][q] : U1[q];
                    D2[q] = reach >= 2 ? R[j + 5][q] : D1[q];
                    U3[q] = U2[q];
                    D3[q] = D2[q];
                }
                o = filter_row(U3, U2, U1, R[j + 3], D1, D2, D3, e, coarse, maxv);
            }
            ovd_st8(drow + (size_t)j * a.dst_pitch, o);
        }
#undef ALF_LOAD_ROW
    }
};

// =========================================================================================================
// Chroma 5x5 + CC-ALF (one CTA = one 64x32 chroma tile of one component)
// =========================================================================================================
#ifndef ALF_CHROMA_MIN_BLOCKS
#define ALF_CHROMA_MIN_BLOCKS 5            /* measured 4 / 5 / 6: 30.0 / 29.1 / 28.8 us per 4K picture; 6 (40 registers) spills since the tile rectangle joined the CTA state */
#endif
struct AlfChromaK {
    static const int THREADS = 256, MIN_BLOCKS = ALF_CHROMA_MIN_BLOCKS, PHASES = 2;
    static const int TW = 64, TH = 32;
    static const int HALO_L = 8;                     // samples staged left of the tile (16 bytes: cp.async granularity)
    static const int PIX_STRIDE = TW + 16, PIX_ROWS = TH + 4;
    static const int SMEM_BYTES = PIX_ROWS * PIX_STRIDE * 2;

    struct Args {
        const uint16_t *src[2];          // cb, cr (pre-ALF)
        uint16_t *dst[2];
        const uint16_t *luma;            // pre-ALF luma for CC-ALF
        int src_pitch, dst_pitch, luma_pitch;
        int pic_w, pic_h;                // LUMA picture size
        int log2_ctu, bitdepth;
        int tw, th, tiles_x_per_ctu, tiles_y_per_ctu, nb_ctu_w;
        int cc_enabled[2];
        const AlfCtuDev *ctu;
        AlfChromaTab tab;
        const TileGridDev *tiles;        // device memory, or NULL (one tile)
    };
    struct State { int dummy; };

    // bx0 .. by1: the rectangle samples are fetched from, in CHROMA samples (luma: twice that): the CTU's tile, or the picture
    struct Tile { int x0, y0, tw, th, partial_v, do_alf, alt, cc_idx, bx0, bx1, by0, by1; };

    OVD Tile tile_of(const Args &a, int bx, int by, int comp)
    {
        Tile t;
        const int ctu_c = (1 << a.log2_ctu) >> 1;
        const int w = a.pic_w >> 1, h = a.pic_h >> 1;
        const int sx = a.tiles_x_per_ctu >> 1, sy = a.tiles_y_per_ctu >> 1;       // 1, 2 or 4 tiles per CTU each way
        int cx = bx >> sx, cy = by >> sy;
        t.x0 = cx * ctu_c + (bx & (a.tiles_x_per_ctu - 1)) * a.tw;
        t.y0 = cy * ctu_c + (by & (a.tiles_y_per_ctu - 1)) * a.th;
        int ctu_h = ovd_min(ctu_c, h - cy * ctu_c);
        t.tw = ovd_min(a.tw, w - t.x0);
        t.th = ovd_min(a.th, cy * ctu_c + ctu_h - t.y0);
        t.partial_v = ((cy + 1) << a.log2_ctu) > a.pic_h;
        t.do_alf = 0; t.alt = 0; t.cc_idx = 0;
        t.bx0 = 0; t.bx1 = w; t.by0 = 0; t.by1 = h;
        if (t.tw > 0 && t.th > 0 && a.tiles) {
            int c0, c1;
            tile_span(a.tiles->col_bd, a.tiles->nb_cols, cx, a.nb_ctu_w, c0, c1);
            t.bx0 = c0 * ctu_c; t.bx1 = ovd_min(c1 * ctu_c, w);
            tile_span(a.tiles->row_bd, a.tiles->nb_rows, cy, (h + ctu_c - 1) / ctu_c, c0, c1);
            t.by0 = c0 * ctu_c; t.by1 = ovd_min(c1 * ctu_c, h);
        }
        if (t.tw > 0 && t.th > 0) {
            AlfCtuDev c = a.ctu[cy * a.nb_ctu_w + cx];
            t.do_alf = comp == 0 ? (c.flag >> 1) & 1 : c.flag & 1;
            t.alt = comp == 0 ? c.cb_alt : c.cr_alt;
            t.cc_idx = a.cc_enabled[comp] ? (comp == 0 ? c.cc_cb : c.cc_cr) : 0;
        }
        return t;
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st;
        const int comp = bz;
        const Tile t = tile_of(a, bx, by, comp);
        if (t.tw <= 0 || t.th <= 0) return;
        uint16_t *pix = (uint16_t *)smem;
        const int w = a.pic_w >> 1, h = a.pic_h >> 1;
        (void)w; (void)h;
        const uint16_t *src = a.src[comp];
        if (P == 0) {
            // stage tile + halo 2 (8 horizontally for alignment): tiles away from the left / right border as 16-byte asynchronous
            // copies (LDGSTS), one lane per chunk and one warp per row; border tiles through the clamped 8-byte path
            const int nrows = t.th + 4, nchunks = (t.tw + 16) >> 2;   // <= 20
            const bool fast = t.x0 >= t.bx0 + HALO_L && t.x0 + t.tw + 8 <= t.bx1 && !(t.tw & 7) && !(a.src_pitch & 7);
            const int ch = tid & 31;
            if (fast) {
                if (ch < ((t.tw + 16) >> 3))
                    for (int r = tid >> 5; r < nrows; r += 8) {
                        const int y = ovd_clamp(t.y0 - 2 + r, t.by0, t.by1 - 1);
                        ovd_cp_async16(pix + r * PIX_STRIDE + 8 * ch, src + (size_t)y * a.src_pitch + t.x0 - HALO_L + 8 * ch);
                    }
                ovd_cp_async_wait_all();
                return;
            }
            if (ch < nchunks) {
                int x = t.x0 - HALO_L + 4 * ch;
                for (int r = tid >> 5; r < nrows; r += 8) {
                    int y = ovd_clamp(t.y0 - 2 + r, t.by0, t.by1 - 1);
                    const uint16_t *row = src + (size_t)y * a.src_pitch;
                    ovd_u2 v;
                    if (x < t.bx0)        { uint32_t e = row[t.bx0];     v.x = v.y = e * 0x10001u; }
                    else if (x >= t.bx1)  { uint32_t e = row[t.bx1 - 1]; v.x = v.y = e * 0x10001u; }
                    else v = ovd_ld8(row + x);
                    *(ovd_u2 *)(pix + r * PIX_STRIDE + 4 * ch) = v;
                }
            }
            return;
        }
        // thread = 4 samples x 2 rows
        const int tx = tid & 15, ty = tid >> 4;
        if (4 * tx >= t.tw || 2 * ty >= t.th) return;
        const int maxv = (1 << a.bitdepth) - 1;
        const int ctu_s = 1 << a.log2_ctu, ctu_c = ctu_s >> 1;
        const int vb = t.partial_v ? a.pic_h / 2 : (ctu_s - 4) / 2;        // rcn_alf.c:1383
        const int vb_cc = t.partial_v ? a.pic_h / 2 : ctu_s - 4;            // rcn_alf.c:1417 (sic)
        const int16_t *cf = a.tab.coeff[t.alt], *cl = a.tab.clip[t.alt];
        const int16_t *cc = a.tab.cc[comp][t.cc_idx ? t.cc_idx - 1 : 0];
        const int half = 1 << (a.bitdepth - 1);
        uint16_t *dst = a.dst[comp];
        int wide = 0;                                   // a coefficient outside int8 (IDP.2A's multiplier width): scalar path
#pragma unroll
        for (int k = 0; k < 6; k++) wide |= cf[k] < -128 || cf[k] > 127;
        int cc_wide = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) cc_wide |= cc[k] < -128 || cc[k] > 127;
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int ly = 2 * ty + rr;                      // row inside tile
            const int y = t.y0 + ly;
            const int y_vb = y & (ctu_c - 1);
            int reach = 2;
            if (y_vb < vb && y_vb >= vb - 2)       reach = vb - 1 - y_vb;
            else if (y_vb >= vb && y_vb <= vb + 1) reach = y_vb - vb;
            const int coarse = (y_vb == vb - 1) || (y_vb == vb);
            const int d1 = ovd_min(1, reach), d2 = ovd_min(2, reach);
            const uint16_t *c0 = pix + (ly + 2) * PIX_STRIDE + 4 * tx + HALO_L;
            const uint16_t *u1 = c0 - d1 * PIX_STRIDE, *u2 = c0 - d2 * PIX_STRIDE;
            const uint16_t *e1 = c0 + d1 * PIX_STRIDE, *e2 = c0 + d2 * PIX_STRIDE;
            // CC-ALF row offsets (rcn_alf.c:831-839)
            const int pos = (y << 1) & (ctu_s - 1);
            int o1 = 1, o2 = -1, o3 = 2;
            if (pos == vb_cc - 2 || pos == vb_cc + 1) o3 = o1;
            else if (pos == vb_cc - 1 || pos == vb_cc) o1 = o2 = o3 = 0;
            const int ly2 = y << 1;
            const uint16_t *l0 = a.luma + (size_t)ly2 * a.luma_pitch;
            // (luma rows / columns of CC-ALF are clamped to the luma rectangle of the tile, like the padded filter region)
            const int ly_lo = 2 * t.by0, ly_hi = ovd_min(2 * t.by1, a.pic_h) - 1, lx_lo = 2 * t.bx0, lx_hi = ovd_min(2 * t.bx1, a.pic_w) - 1;
            const uint16_t *l1 = a.luma + (size_t)ovd_clamp(ly2 + o1, ly_lo, ly_hi) * a.luma_pitch;
            const uint16_t *l2 = a.luma + (size_t)ovd_clamp(ly2 + o2, ly_lo, ly_hi) * a.luma_pitch;
            const uint16_t *l3 = a.luma + (size_t)ovd_clamp(ly2 + o3, ly_lo, ly_hi) * a.luma_pitch;
            uint32_t outv[4];
            int alfv[4] = { c0[0], c0[1], c0[2], c0[3] };
            if (t.do_alf && !wide) {
                // 5x5 diamond on packed sample pairs (VIADDMNMX.S16x2 clip, IDP.2A accumulate), as the luma kernel does: the
                // thread's four samples are words 1 and 2 of the four-word windows below (samples -2 .. 5)
                const uint32_t *Cw = (const uint32_t *)(c0 - 2), *U1w = (const uint32_t *)(u1 - 2), *E1w = (const uint32_t *)(e1 - 2);
                const uint32_t *U2w = (const uint32_t *)u2, *E2w = (const uint32_t *)e2;
                const uint32_t cw0 = Cw[0], cw1 = Cw[1], cw2 = Cw[2], cw3 = Cw[3];
                const uint32_t uw0 = U1w[0], uw1 = U1w[1], uw2 = U1w[2], uw3 = U1w[3];
                const uint32_t ew0 = E1w[0], ew1 = E1w[1], ew2 = E1w[2], ew3 = E1w[3];
                const uint32_t cwv[4] = { cw0, cw1, cw2, cw3 }, uwv[4] = { uw0, uw1, uw2, uw3 }, ewv[4] = { ew0, ew1, ew2, ew3 };
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint32_t pp = cwv[1 + p], neg_p = ovd_vsub_16x2(0u, pp);
                    int acc0 = 0, acc1 = 0;
#define CP2(k, A, B)                                                                                                          \
                    do {                                                                                                      \
                        const uint32_t cl2 = (uint32_t)(uint16_t)cl[k] * 0x10001u, ncl2 = ovd_vsub_16x2(0u, cl2);                 \
                        const uint32_t s2 = ovd_vadd_16x2(ovd_vmax_s16x2(ovd_viaddmin_s16x2(A, neg_p, cl2), ncl2),                \
                                                          ovd_vmax_s16x2(ovd_viaddmin_s16x2(B, neg_p, cl2), ncl2));               \
                        const uint32_t cf2 = (uint32_t)(uint8_t)cf[k] * 0x01000001u;                                              \
                        acc0 = ovd_dp2a_lo(s2, cf2, acc0); acc1 = ovd_dp2a_hi(s2, cf2, acc1);                                     \
                    } while (0)
                    CP2(0, E2w[p], U2w[p]);
                    CP2(1, alf_odd_pair(ewv[1 + p], ewv[2 + p]), alf_odd_pair(uwv[p], uwv[1 + p]));
                    CP2(2, ewv[1 + p], uwv[1 + p]);
                    CP2(3, alf_odd_pair(ewv[p], ewv[1 + p]), alf_odd_pair(uwv[1 + p], uwv[2 + p]));
                    CP2(4, cwv[2 + p], cwv[p]);
                    CP2(5, alf_odd_pair(cwv[1 + p], cwv[2 + p]), alf_odd_pair(cwv[p], cwv[1 + p]));
#undef CP2
                    acc0 = coarse ? (acc0 + 512) >> 10 : (acc0 + 64) >> 7;
                    acc1 = coarse ? (acc1 + 512) >> 10 : (acc1 + 64) >> 7;
                    alfv[2 * p] = ovd_clamp(acc0 + (int)(pp & 0xffff), 0, maxv);
                    alfv[2 * p + 1] = ovd_clamp(acc1 + (int)(pp >> 16), 0, maxv);
                }
            }
            int ccv[4] = { 0, 0, 0, 0 };
            if (t.cc_idx && !cc_wide) {
                // CC-ALF (rcn_alf.c:826-870) for the four chroma samples at once: their eight luma columns come in as four 8-byte
                // loads per row (+ the column to the left), the seven differences to the collocated luma sample are formed on
                // packed pairs (even / odd luma columns picked with PRMT) and accumulated with IDP.2A
                const int lx0 = (t.x0 + 4 * tx) << 1, lxm = ovd_max(lx0 - 1, lx_lo);
                const ovd_u2 r0a = ovd_ld8(l0 + lx0), r0b = ovd_ld8(l0 + lx0 + 4), r1a = ovd_ld8(l1 + lx0), r1b = ovd_ld8(l1 + lx0 + 4);
                const ovd_u2 r2a = ovd_ld8(l2 + lx0), r2b = ovd_ld8(l2 + lx0 + 4), r3a = ovd_ld8(l3 + lx0), r3b = ovd_ld8(l3 + lx0 + 4);
                const uint32_t w0[5] = { (uint32_t)l0[lxm] << 16, r0a.x, r0a.y, r0b.x, r0b.y };
                const uint32_t w1[5] = { (uint32_t)l1[lxm] << 16, r1a.x, r1a.y, r1b.x, r1b.y };
                const uint32_t w2[4] = { r2a.x, r2a.y, r2b.x, r2b.y }, w3[4] = { r3a.x, r3a.y, r3b.x, r3b.y };
                uint32_t ccw[7];
#pragma unroll
                for (int k = 0; k < 7; k++) ccw[k] = (uint32_t)(uint8_t)cc[k] * 0x01000001u;
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint32_t ctr = ovd_byte_perm(w0[1 + 2 * p], w0[2 + 2 * p], 0x5410);       // l0[lx], l0[lx + 2]
                    int acc0 = 0, acc1 = 0;
#define CC2(k, X) do { const uint32_t d2 = ovd_vsub_16x2(X, ctr); acc0 = ovd_dp2a_lo(d2, ccw[k], acc0); acc1 = ovd_dp2a_hi(d2, ccw[k], acc1); } while (0)
                    CC2(0, ovd_byte_perm(w2[2 * p], w2[1 + 2 * p], 0x5410));
                    CC2(1, ovd_byte_perm(w0[2 * p], w0[1 + 2 * p], 0x7632));
                    CC2(2, ovd_byte_perm(w0[1 + 2 * p], w0[2 + 2 * p], 0x7632));
                    CC2(3, ovd_byte_perm(w1[2 * p], w1[1 + 2 * p], 0x7632));
                    CC2(4, ovd_byte_perm(w1[1 + 2 * p], w1[2 + 2 * p], 0x5410));
                    CC2(5, ovd_byte_perm(w1[1 + 2 * p], w1[2 + 2 * p], 0x7632));
                    CC2(6, ovd_byte_perm(w3[2 * p], w3[1 + 2 * p], 0x5410));
#undef CC2
                    ccv[2 * p] = acc0; ccv[2 * p + 1] = acc1;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int cur = c0[i];
                int v = alfv[i];
                if (t.do_alf && wide) {

                    int sum = 0;
#define CP(k, A, B) sum += cf[k] * (ovd_clamp((int)(A) - cur, -cl[k], cl[k]) + ovd_clamp((int)(B) - cur, -cl[k], cl[k]))
                    CP(0, e2[i], u2[i]);
                    CP(1, e1[i + 1], u1[i - 1]);
                    CP(2, e1[i], u1[i]);
                    CP(3, e1[i - 1], u1[i + 1]);
                    CP(4, c0[i + 2], c0[i - 2]);
                    CP(5, c0[i + 1], c0[i - 1]);
#undef CP
                    sum = coarse ? (sum + 512) >> 10 : (sum + 64) >> 7;
                    v = ovd_clamp(sum + cur, 0, maxv);
                }
                if (t.cc_idx) {
                    int sum = ccv[i];
                    if (cc_wide) {
                        const int lx = (t.x0 + 4 * tx + i) << 1;
                        const int lxm = ovd_max(lx - 1, lx_lo), lxp = ovd_min(lx + 1, lx_hi);
                        const int lc = l0[lx];
                        sum = cc[0] * (l2[lx] - lc) + cc[1] * (l0[lxm] - lc) + cc[2] * (l0[lxp] - lc)
                            + cc[3] * (l1[lxm] - lc) + cc[4] * (l1[lx] - lc) + cc[5] * (l1[lxp] - lc)
                            + cc[6] * (l3[lx] - lc);
                    }
                    sum = (sum + 64) >> 7;
                    sum = ovd_clamp(sum + half, 0, maxv);
                    v = ovd_clamp(sum + v - half, 0, maxv);
                }
                outv[i] = (uint32_t)v;
            }
            ovd_u2 o;
            o.x = outv[0] | (outv[1] << 16);
            o.y = outv[2] | (outv[3] << 16);
            ovd_st8(dst + (size_t)y * a.dst_pitch + t.x0 + 4 * tx, o);
        }
    }
};
