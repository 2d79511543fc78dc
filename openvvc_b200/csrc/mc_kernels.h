// mc_kernels.h -- translational motion compensation (uni / bi / BCW, luma 8-tap + chroma 4-tap) for a whole
// picture's PU list (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: dispatch rcn_mcp_b & co (libovvc/rcn_inter.c:2750-2966), MV clip + fraction split +
// addressing (rcn_inter.c:96-108, 521-602, 1392-1554, 1822-1900), border replication = clamped coordinates
// (emulate_block_border, rcn_inter.c:148-225), interpolation kernels libovvc/rcn_mc.c:381-1608, forward LMCS
// reshape of the luma prediction (rcn_inter.c:599).
//
// One warp owns one work item = a <=16x16 luma block (+ its 8x8 Cb/Cr blocks) of one PU.  The reference
// window of each used list is staged in shared memory as packed pairs of samples; the separable filter runs
// as IDP.2A dot products: an output at an even window position takes 4 (luma) / 2 (chroma) two-tap products on
// aligned pairs, an output at an odd position 5 / 3 products with a shifted coefficient packing, so no
// realignment is needed.  Zero fractions use the identity filter {..,64,..}, which reproduces the reference's
// pel / h / v / hv variants bit for bit (64*s >> (BD-8) == s << (14-BD); (64*t) >> 6 == t).
#pragma once
#include "dev.h"

#define MC_LUMA 1
#define MC_CHROMA 2
#define MC_HPEL_AMVR 4

struct McCmdDev {            // == OVGPUMcCmd
    int16_t x0, y0;
    uint8_t log2_w, log2_h, inter_dir, flags;
    int8_t ref_idx0, ref_idx1;
    uint8_t bcw_idx_plus1, tool;
    int32_t mv0_x, mv0_y, mv1_x, mv1_y;
    uint32_t aux;
};

struct McItemDev {           // one <=16x16 luma block of command `cmd`
    uint32_t cmd;
    uint8_t ox, oy, w, h;    // offset inside the PU and size, luma samples
};

struct McRefsDev {
    const uint16_t *plane[2][16][3];
    int pitch[2][16][3];     // samples
    int poc[2][16];
};

// filter coefficient banks (constant memory on the device)
OVD_CONST int8_t mc_filt_l_tab[17][8] = {
    {  0, 0,   0, 64,  0,   0,  0,  0 },
    {  0, 1,  -3, 63,  4,  -2,  1,  0 }, { -1, 2,  -5, 62,  8,  -3,  1,  0 }, { -1, 3,  -8, 60, 13,  -4,  1,  0 },
    { -1, 4, -10, 58, 17,  -5,  1,  0 }, { -1, 4, -11, 52, 26,  -8,  3, -1 }, { -1, 3,  -9, 47, 31, -10,  4, -1 },
    { -1, 4, -11, 45, 34, -10,  4, -1 }, { -1, 4, -11, 40, 40, -11,  4, -1 }, { -1, 4, -10, 34, 45, -11,  4, -1 },
    { -1, 4, -10, 31, 47,  -9,  3, -1 }, { -1, 3,  -8, 26, 52, -11,  4, -1 }, {  0, 1,  -5, 17, 58, -10,  4, -1 },
    {  0, 1,  -4, 13, 60,  -8,  3, -1 }, {  0, 1,  -3,  8, 62,  -5,  2, -1 }, {  0, 1,  -2,  4, 63,  -3,  1,  0 },
    {  0, 3,   9, 20, 20,   9,  3,  0 } };
OVD_CONST int8_t mc_filt_l4_tab[16][8] = {
    { 0, 0,   0, 64,  0,   0, 0, 0 },
    { 0, 1,  -3, 63,  4,  -2, 1, 0 }, { 0, 1,  -5, 62,  8,  -3, 1, 0 }, { 0, 2,  -8, 60, 13,  -4, 1, 0 },
    { 0, 3, -10, 58, 17,  -5, 1, 0 }, { 0, 3, -11, 52, 26,  -8, 2, 0 }, { 0, 2,  -9, 47, 31, -10, 3, 0 },
    { 0, 3, -11, 45, 34, -10, 3, 0 }, { 0, 3, -11, 40, 40, -11, 3, 0 }, { 0, 3, -10, 34, 45, -11, 3, 0 },
    { 0, 3, -10, 31, 47,  -9, 2, 0 }, { 0, 2,  -8, 26, 52, -11, 3, 0 }, { 0, 1,  -5, 17, 58, -10, 3, 0 },
    { 0, 1,  -4, 13, 60,  -8, 2, 0 }, { 0, 1,  -3,  8, 62,  -5, 1, 0 }, { 0, 1,  -2,  4, 63,  -3, 1, 0 } };
OVD_CONST int8_t mc_filt_c_tab[32][4] = {
    {  0, 64,  0,  0 },
    { -1, 63,  2,  0 }, { -2, 62,  4,  0 }, { -2, 60,  7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
    { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 },
    { -6, 49, 24, -3 }, { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 },
    { -4, 36, 36, -4 }, { -4, 33, 39, -4 }, { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 },
    { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 }, { -2, 16, 54, -4 }, { -2, 15, 55, -4 },
    { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1,  7, 60, -2 }, {  0,  4, 62, -2 },
    {  0,  2, 63, -1 } };

struct McK {
    static const int THREADS = 128, MIN_BLOCKS = 4, PHASES = 3;
    static const int WARPS = 4;
    // per-warp shared memory (bytes)
    static const int WIN_Y_WORDS = 12, WIN_Y_ROWS = 23;
    static const int WIN_C_WORDS = 6, WIN_C_ROWS = 11;
    static const int TMP_Y_COL = 26;       // int16 per column (23 rows, padded: 52 B keeps columns on distinct banks)
    static const int TMP_C_COL = 12;
    static const int SM_WIN_Y = 2 * WIN_Y_ROWS * WIN_Y_WORDS * 4;        // [ref][row][word]
    static const int SM_WIN_C = 2 * 2 * WIN_C_ROWS * WIN_C_WORDS * 4;    // [ref][comp][row][word]
    static const int SM_TMP_Y = 2 * 16 * TMP_Y_COL * 2;                  // [ref][col][row]
    static const int SM_TMP_C = 2 * 2 * 8 * TMP_C_COL * 2;               // [ref][comp][col][row]
    static const int SM_WARP = SM_WIN_Y + SM_WIN_C + SM_TMP_Y + SM_TMP_C;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        uint16_t *dst[3];
        int dst_pitch[3];                // samples
        int pic_w, pic_h, bitdepth;
        const McCmdDev *cmds;
        const McItemDev *items;
        int nb_items;
        const uint16_t *lmcs_fwd;        // device LUT or NULL
        McRefsDev refs;
    };
    struct State { int dummy; };

    // everything a phase needs to know about the warp's item
    struct Item {
        int valid;
        int bx, by, w, h;                // luma block position / size in the picture
        int pu_small;                    // 4x4 PU -> 6-tap table
        int nref;                        // 1 uni, 2 bi
        int list[2], ridx[2];
        int mvx[2], mvy[2];              // clipped MVs
        int flags, bcw;
    };

    OVD Item load_item(const Args &a, int item_idx)
    {
        Item it;
        it.valid = item_idx < a.nb_items;
        if (!it.valid) return it;
        const McItemDev im = a.items[item_idx];
        const McCmdDev c = a.cmds[im.cmd];
        it.bx = c.x0 + im.ox; it.by = c.y0 + im.oy; it.w = im.w; it.h = im.h;
        const int pu_w = 1 << c.log2_w, pu_h = 1 << c.log2_h;
        it.pu_small = (pu_w == 4 && pu_h == 4);
        it.flags = c.flags;
        // check_identical_motion (rcn_inter.c:256-267) and the dispatch of rcn_mcp_b (:2785-2812)
        const int identical = c.inter_dir == 3 &&
                              (uint16_t)a.refs.poc[0][c.ref_idx0] == (uint16_t)a.refs.poc[1][c.ref_idx1] &&
                              c.mv0_x == c.mv1_x && c.mv0_y == c.mv1_y;
        const int bi = c.inter_dir == 3 && !identical;
        if (bi) {
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
        it.bcw = (c.bcw_idx_plus1 == 0 || c.bcw_idx_plus1 == 3) ? 0 : c.bcw_idx_plus1;
        // clip_mv (rcn_inter.c:96-108) with the PU's position and size
        const int x_max = (a.pic_w + 2 - c.x0) << 4, y_max = (a.pic_h + 2 - c.y0) << 4;
        const int x_min = -((pu_w + 3 + c.x0) << 4), y_min = -((pu_h + 3 + c.y0) << 4);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            it.mvx[r] = ovd_clamp(it.mvx[r], x_min, x_max);
            it.mvy[r] = ovd_clamp(it.mvy[r], y_min, y_max);
        }
        return it;
    }

    // one packed pair of horizontally adjacent samples at (xa, xa+1), xa even, coordinates clamped to the plane
    OVD uint32_t load_pair(const uint16_t *plane, int pitch, int w, int h, int xa, int y)
    {
        const uint16_t *row = plane + (size_t)ovd_clamp(y, 0, h - 1) * pitch;
        if (xa >= 0 && xa + 1 < w) return *(const uint32_t *)(row + xa);
        return (uint32_t)row[ovd_clamp(xa, 0, w - 1)] | ((uint32_t)row[ovd_clamp(xa + 1, 0, w - 1)] << 16);
    }

    // coefficient packing: word j = bytes [c(2j), c(2j+1) | c(2j-1), c(2j)] (see file header)
    OVD uint32_t coef_word(const int8_t *f, int ntaps, int j)
    {
        int c0 = 2 * j < ntaps ? f[2 * j] : 0, c1 = 2 * j + 1 < ntaps ? f[2 * j + 1] : 0;
        int cm = (2 * j - 1 >= 0 && 2 * j - 1 < ntaps) ? f[2 * j - 1] : 0;
        return (uint32_t)(uint8_t)c0 | ((uint32_t)(uint8_t)c1 << 8) | ((uint32_t)(uint8_t)cm << 16) | ((uint32_t)(uint8_t)c0 << 24);
    }

    // 8-tap (NW = 5 coefficient words) or 4-tap (NW = 3) filter at window sample index i over packed pairs W[]
    template <int NW>
    OVD int fir(const uint32_t *W, int i, const uint32_t *B)
    {
        const int q = i >> 1;
        int acc = 0;
        if (i & 1) {
#pragma unroll
            for (int j = 0; j < NW; j++) acc = ovd_dp2a_hi(W[q + j], B[j], acc);
        } else {
#pragma unroll
            for (int j = 0; j < NW - 1; j++) acc = ovd_dp2a_lo(W[q + j], B[j], acc);
        }
        return acc;
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31;
        const Item it = load_item(a, bx * WARPS + warp);
        if (!it.valid) return;
        unsigned char *sm = smem + warp * SM_WARP;
        uint32_t *win_y = (uint32_t *)sm;
        uint32_t *win_c = (uint32_t *)(sm + SM_WIN_Y);
        int16_t *tmp_y = (int16_t *)(sm + SM_WIN_Y + SM_WIN_C);
        int16_t *tmp_c = (int16_t *)(sm + SM_WIN_Y + SM_WIN_C + SM_TMP_Y);
        const int do_y = it.flags & MC_LUMA, do_c = it.flags & MC_CHROMA;
        const int cw = a.pic_w >> 1, ch = a.pic_h >> 1;

        if (P == 0) {
            // ---- stage reference windows -------------------------------------------------------------------
            for (int r = 0; r < it.nref; r++) {
                if (do_y) {
                    const uint16_t *pl = a.refs.plane[it.list[r]][it.ridx[r]][0];
                    const int pitch = a.refs.pitch[it.list[r]][it.ridx[r]][0];
                    const int x_al = (it.bx + (it.mvx[r] >> 4) - 3) & ~1, y0 = it.by + (it.mvy[r] >> 4) - 3;
                    for (int i = lane; i < WIN_Y_ROWS * WIN_Y_WORDS; i += 32) {
                        int row = i / WIN_Y_WORDS, wd = i - row * WIN_Y_WORDS;
                        win_y[(r * WIN_Y_ROWS + row) * WIN_Y_WORDS + wd] = load_pair(pl, pitch, a.pic_w, a.pic_h, x_al + 2 * wd, y0 + row);
                    }
                }
                if (do_c) {
                    const int x_al = ((it.bx >> 1) + (it.mvx[r] >> 5) - 1) & ~1, y0 = (it.by >> 1) + (it.mvy[r] >> 5) - 1;
                    for (int i = lane; i < 2 * WIN_C_ROWS * WIN_C_WORDS; i += 32) {
                        int comp = i / (WIN_C_ROWS * WIN_C_WORDS), k = i - comp * (WIN_C_ROWS * WIN_C_WORDS);
                        int row = k / WIN_C_WORDS, wd = k - row * WIN_C_WORDS;
                        const uint16_t *pl = a.refs.plane[it.list[r]][it.ridx[r]][1 + comp];
                        const int pitch = a.refs.pitch[it.list[r]][it.ridx[r]][1 + comp];
                        win_c[((r * 2 + comp) * WIN_C_ROWS + row) * WIN_C_WORDS + wd] = load_pair(pl, pitch, cw, ch, x_al + 2 * wd, y0 + row);
                    }
                }
            }
            return;
        }

        const int sh1 = a.bitdepth - 8;
        if (P == 1) {
            // ---- horizontal pass: window rows -> column-major 14-bit intermediates ---------------------------
            if (do_y) {
                for (int task = lane; task < it.nref * WIN_Y_ROWS * 2; task += 32) {
                    const int r = task / (WIN_Y_ROWS * 2), k = task - r * (WIN_Y_ROWS * 2);
                    const int row = k >> 1, half = k & 1;
                    int px = it.mvx[r] & 15;
                    if ((it.flags & MC_HPEL_AMVR) && !it.pu_small && px == 8) px = 16;  // rcn_inter.c:572-577
                    const int8_t *f = it.pu_small ? mc_filt_l4_tab[px] : mc_filt_l_tab[px];
                    uint32_t B[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) B[j] = coef_word(f, 8, j);
                    const int o = (it.bx + (it.mvx[r] >> 4) - 3) & 1;
                    const uint32_t *wrow = win_y + (r * WIN_Y_ROWS + row) * WIN_Y_WORDS + 4 * half;
                    uint32_t W[9];
#pragma unroll
                    for (int j = 0; j < 9; j++) W[j] = wrow[(4 * half + j < WIN_Y_WORDS) ? j : 0];
                    int16_t *tcol = tmp_y + (r * 16 + 8 * half) * TMP_Y_COL + row;
                    // o is uniform for the warp: two fully unrolled variants keep W[] in registers
                    if (o) {
#pragma unroll
                        for (int x = 0; x < 8; x++) tcol[x * TMP_Y_COL] = (int16_t)(fir<5>(W, x + 1, B) >> sh1);
                    } else {
#pragma unroll
                        for (int x = 0; x < 8; x++) tcol[x * TMP_Y_COL] = (int16_t)(fir<5>(W, x, B) >> sh1);
                    }
                }
            }
            if (do_c) {
                for (int task = lane; task < it.nref * 2 * WIN_C_ROWS; task += 32) {
                    const int rc = task / WIN_C_ROWS, row = task - rc * WIN_C_ROWS;     // rc = ref * 2 + comp
                    const int r = rc >> 1;
                    const int8_t *f = mc_filt_c_tab[it.mvx[r] & 31];
                    uint32_t B[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) B[j] = coef_word(f, 4, j);
                    const int o = ((it.bx >> 1) + (it.mvx[r] >> 5) - 1) & 1;
                    const uint32_t *wrow = win_c + (rc * WIN_C_ROWS + row) * WIN_C_WORDS;
                    uint32_t W[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) W[j] = wrow[j];
                    int16_t *tcol = tmp_c + rc * 8 * TMP_C_COL + row;
                    if (o) {
#pragma unroll
                        for (int x = 0; x < 8; x++) tcol[x * TMP_C_COL] = (int16_t)(fir<3>(W, x + 1, B) >> sh1);
                    } else {
#pragma unroll
                        for (int x = 0; x < 8; x++) tcol[x * TMP_C_COL] = (int16_t)(fir<3>(W, x, B) >> sh1);
                    }
                }
            }
            return;
        }

        // ---- vertical pass + combination + store --------------------------------------------------------------
        const int maxv = (1 << a.bitdepth) - 1;
        const int w1 = it.bcw == 1 ? -2 : it.bcw == 2 ? 3 : it.bcw == 4 ? 5 : 10;           // rcn_inter.c:89
        if (do_y) {
            const int x = lane & 15, g = lane >> 4;
            int p[2][8];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (r < it.nref) {
                    int py = it.mvy[r] & 15;
                    if ((it.flags & MC_HPEL_AMVR) && !it.pu_small && py == 8) py = 16;
                    const int8_t *f = it.pu_small ? mc_filt_l4_tab[py] : mc_filt_l_tab[py];
                    uint32_t B[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) B[j] = coef_word(f, 8, j);
                    const uint32_t *tw = (const uint32_t *)(tmp_y + (r * 16 + x) * TMP_Y_COL) + 4 * g;
                    uint32_t W[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) W[j] = tw[j];
#pragma unroll
                    for (int k = 0; k < 8; k++) p[r][k] = fir<5>(W, k, B) >> 6;
                }
            }
            if (x < it.w) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int y = 8 * g + k;
                    if (y >= it.h) continue;
                    int v;
                    if (it.nref == 1)   v = (p[0][k] + (1 << (13 - a.bitdepth))) >> (14 - a.bitdepth);
                    else if (!it.bcw)   v = (p[0][k] + p[1][k] + (1 << (14 - a.bitdepth))) >> (15 - a.bitdepth);
                    else                v = (p[1][k] * w1 + p[0][k] * (8 - w1) + (1 << (16 - a.bitdepth))) >> (17 - a.bitdepth);
                    v = ovd_clamp(v, 0, maxv);
                    if (a.lmcs_fwd) v = a.lmcs_fwd[v];
                    a.dst[0][(size_t)(it.by + y) * a.dst_pitch[0] + it.bx + x] = (uint16_t)v;
                }
            }
        }
        if (do_c) {
            const int x = lane & 7, comp = (lane >> 3) & 1, g = lane >> 4;
            int p[2][4];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (r < it.nref) {
                    const int8_t *f = mc_filt_c_tab[it.mvy[r] & 31];
                    uint32_t B[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) B[j] = coef_word(f, 4, j);
                    const uint32_t *tw = (const uint32_t *)(tmp_c + ((r * 2 + comp) * 8 + x) * TMP_C_COL) + 2 * g;
                    uint32_t W[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) W[j] = tw[j];
#pragma unroll
                    for (int k = 0; k < 4; k++) p[r][k] = fir<3>(W, k, B) >> 6;
                }
            }
            if (x < (it.w >> 1)) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int y = 4 * g + k;
                    if (y >= (it.h >> 1)) continue;
                    int v;
                    if (it.nref == 1)   v = (p[0][k] + (1 << (13 - a.bitdepth))) >> (14 - a.bitdepth);
                    else if (!it.bcw)   v = (p[0][k] + p[1][k] + (1 << (14 - a.bitdepth))) >> (15 - a.bitdepth);
                    else                v = (p[1][k] * w1 + p[0][k] * (8 - w1) + (1 << (16 - a.bitdepth))) >> (17 - a.bitdepth);
                    a.dst[1 + comp][(size_t)((it.by >> 1) + y) * a.dst_pitch[1 + comp] + (it.bx >> 1) + x] = (uint16_t)ovd_clamp(v, 0, maxv);
                }
            }
        }
    }
};
