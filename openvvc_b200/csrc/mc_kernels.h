// mc_kernels.h -- translational motion compensation (uni / bi / BCW, luma 8-tap + chroma 4-tap) for a whole
// picture's PU list (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: dispatch rcn_mcp_b & co (libovvc/rcn_inter.c:2750-2966), MV clip + fraction split +
// addressing (rcn_inter.c:96-108, 521-602, 1392-1554, 1822-1900), border replication = clamped coordinates
// (emulate_block_border, rcn_inter.c:148-225), interpolation kernels libovvc/rcn_mc.c:381-1608, forward LMCS
// reshape of the luma prediction (rcn_inter.c:599).
//
// One warp owns one work item = a <=16x16 luma block (+ its 8x8 Cb/Cr blocks) of one PU.
//   phase 0  the reference window of each used list is staged in shared memory as packed pairs of samples with
//            8-byte loads (clamped per-sample loads only for windows that cross the picture border);
//   phase 1  every lane owns one output COLUMN of one list (bi-prediction: lanes 0-15 list 0, 16-31 list 1; uni:
//            the two half-warps split the rows): it runs the horizontal filter down the column, keeps the 14-bit
//            intermediates packed in registers and applies the vertical filter to them, all as IDP.2A dot products
//            on (sample pair) x (tap pair) -- an output whose first tap sits on an odd sample index uses the shifted
//            tap packing of mc_tables.h, so no realignment is needed;
//   phase 2  the two lists are combined (average / BCW), clipped, forward-LMCS'ed (luma) and stored.
// Zero fractions use the identity filter {..,64,..}, which reproduces the reference's pel / h / v / hv variants bit
// for bit (64*s >> (BD-8) == s << (14-BD); (64*t) >> 6 == t).
#pragma once
#include "dev.h"
#include "mc_tables.h"

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

// GPM blend weights (rcn_init_gpm_params, libovvc/rcn_gpm.c:150-204; rcn_gpm_weights_and_steps, libovvc/rcn_inter.c:3080-3115):
// the reference keeps six 112 x 112 masks and walks them mirrored for the other angles; the same value in closed form.
OVD_CONST int8_t gpm_angle2mask[32] = { 0, -1, 1, 2, 3, 4, -1, -1, 5, -1, -1, 4, 3, 2, 1, -1, 0, -1, 1, 2, 3, 4, -1, -1, 5, -1, -1, 4, 3, 2, 1, -1 };
OVD_CONST int8_t gpm_angle2mirror[32] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2 };
OVD_CONST int8_t gpm_dis[32] = { 8, 8, 8, 8, 4, 4, 2, 1, 0, -1, -2, -4, -4, -8, -8, -8, -8, -8, -8, -8, -4, -4, -2, -1, 0, 1, 2, 4, 4, 8, 8, 8 };
OVD_CONST uint8_t gpm_angle[64] = { 0, 0, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 8, 8, 11, 11, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13,
                                    14, 14, 14, 14, 16, 16, 18, 18, 18, 19, 19, 19, 20, 20, 20, 21, 21, 21, 24, 24, 27, 27, 27, 28, 28, 28, 29, 29, 29, 30, 30, 30 };
OVD_CONST uint8_t gpm_dist[64] = { 1, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 1, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3,
                                   0, 1, 2, 3, 1, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3 };
OVD_CONST int8_t gpm_base_angle[6] = { 0, 2, 3, 4, 5, 8 };

struct GpmGeom { int mirror, ox, oy, dx, dy, rho; };       // per CU: everything but the sample position
OVD GpmGeom gpm_geom(int split_dir, int lw, int lh)
{
    GpmGeom g;
    const int angle = gpm_angle[split_dir], distance = gpm_dist[split_dir], w = 1 << lw, h = 1 << lh;
    g.ox = (112 - w) >> 1; g.oy = (112 - h) >> 1;
    if (distance > 0) {
        if ((angle & 15) == 8 || ((angle & 15) != 0 && h >= w)) g.oy += angle < 16 ? (distance * h) >> 3 : -((distance * h) >> 3);
        else                                                    g.ox += angle < 16 ? (distance * w) >> 3 : -((distance * w) >> 3);
    }
    g.mirror = gpm_angle2mirror[angle];
    const int a0 = gpm_base_angle[gpm_angle2mask[angle]];
    g.dx = gpm_dis[a0]; g.dy = gpm_dis[(a0 + 8) & 31];
    g.rho = (g.dx << 7) + (g.dy << 7);
    return g;
}
// weight (0..8) of the FIRST prediction at luma position (x, y) of the CU
OVD int gpm_weight(const GpmGeom &g, int x, int y)
{
    const int mx = g.mirror == 1 ? 111 - g.ox - x : g.ox + x, my = g.mirror == 2 ? 111 - g.oy - y : g.oy + y;
    const int idx = (((mx + 8) << 1) + 1) * g.dx + (((my + 8) << 1) + 1) * g.dy - g.rho;
    return ovd_clamp((32 + idx + 4) >> 3, 0, 8);
}

struct McK {
    static const int THREADS = 128, MIN_BLOCKS = 8, PHASES = 3;
    static const int WARPS = 4;
    // per-warp shared memory, 32-bit words
    static const int WIN_Y_WORDS = 14, WIN_Y_ROWS = 23;      // 28 samples x (16 + 7) rows per list
    static const int WIN_C_WORDS = 8, WIN_C_ROWS = 11;       // 16 samples x (8 + 3) rows per list and component
    static const int SM_ITEM = 16;
    static const int SM_WIN_Y = 2 * WIN_Y_ROWS * WIN_Y_WORDS;            // [list][row][word]
    static const int SM_WIN_C = 2 * 2 * WIN_C_ROWS * WIN_C_WORDS;        // [list][comp][row][word]
    static const int SM_EX_Y = 2 * 16 * 16 / 2;                          // [list][row][col] int16
    static const int SM_EX_C = 2 * 2 * 8 * 8 / 2;                        // [list][comp][row][col] int16
    static const int SM_WARP = (SM_ITEM + SM_WIN_Y + SM_WIN_C + SM_EX_Y + SM_EX_C) * 4;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        uint16_t *dst[3];
        int dst_pitch[3];                // samples
        int pic_w, pic_h, bitdepth;
        const McCmdDev *cmds;
        const McItemDev *items;
        int nb_items;                    // items the launch was sized for (the host's count)
        const int32_t *limit;            // device: items the list really holds (the device-side walk's count; NULL: nb_items)
        const uint16_t *lmcs_fwd;        // device LUT or NULL
        McRefsDev refs;
    };
    struct State { int dummy; };

    // The lists are built on the device (mc_host.h); the host sizes the launch from its own count of the commands and the kernels
    // stop at the list's real length, which is 0 when the walk found a malformed command or the picture / references do not fit.
    OVD int nb_items(const Args &a) { return a.limit ? ovd_min(a.nb_items, *a.limit) : a.nb_items; }

    // everything a phase needs to know about the warp's item
    struct Item {
        int valid;
        int bx, by, w, h;                // luma block position / size in the picture
        int pu_small;                    // 4x4 PU -> 6-tap table
        int nref;                        // 1 uni, 2 bi
        int list[2], ridx[2];
        int mvx[2], mvy[2];              // clipped MVs
        int flags, bcw;
        int gpm;                         // 0, or 1 + split | ox << 8 | oy << 16 | log2_w << 24 | log2_h << 28 (offset of the item inside the CU)
    };

    OVD Item load_item(const Args &a, int item_idx)
    {
        Item it;
        it.valid = item_idx < nb_items(a);
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
        const int gpm = c.tool == 4;                                // OVGPU_MC_TOOL_GPM: both predictions always, lists from aux
        const int bi = (c.inter_dir == 3 && !identical) || gpm;
        it.gpm = gpm ? (int)(1 + (c.aux & 63)) | (im.ox << 8) | (im.oy << 16) | (c.log2_w << 24) | (c.log2_h << 28) : 0;
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

    // horizontal filter at window row `wrow` (already offset to the lane's first word): NW IDP.2A.LO
    template <int NW>
    OVD int hfir(const uint32_t *wrow, const uint32_t *B)
    {
        int acc = 0;
#pragma unroll
        for (int j = 0; j < NW; j++) acc = ovd_dp2a_lo(wrow[j], B[j], acc);
        return acc;
    }
    OVD uint32_t pack2(int a, int b) { return (uint32_t)(uint16_t)(int16_t)a | ((uint32_t)b << 16); }

    // One output column: horizontal filter over NOUT + NT - 1 window rows, vertical filter over the packed results.
    // NT = 8 (luma, NW = 5 words) or 4 (chroma, NW = 3 words).  ex: this column of the exchange buffer, row stride exs.
    // Fully unrolled and branch-free: NOUT is the block height (a power of two), chosen by a warp-uniform switch.
    template <int NOUT, int NT, int NW>
    OVD void column(const uint32_t *win, int row_words, const uint32_t *Bh, const uint32_t *Bv0, const uint32_t *Bv1,
                    int sh1, int16_t *ex, int exs)
    {
        const int NR = NOUT + NT - 1, NP = (NR + 1) / 2;    // window rows, packed pairs of intermediates
        uint32_t T[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            const int a = hfir<NW>(win + (2 * k) * row_words, Bh) >> sh1;
            const int b = 2 * k + 1 < NR ? hfir<NW>(win + (2 * k + 1) * row_words, Bh) >> sh1 : 0;
            T[k] = pack2(a, b);
        }
#pragma unroll
        for (int y = 0; y < NOUT; y++) {
            const int k = y >> 1;
            int acc = 0;
            if (y & 1) {
#pragma unroll
                for (int j = 0; j < NT / 2 + 1; j++) acc = ovd_dp2a_lo(T[k + j], Bv1[j], acc);
            } else {
#pragma unroll
                for (int j = 0; j < NT / 2; j++) acc = ovd_dp2a_lo(T[k + j], Bv0[j], acc);
            }
            ex[y * exs] = (int16_t)(acc >> 6);
        }
    }

    // final sample from the two 14-bit predictions (rcn_mc.c put_*_uni / put_*_bi / put_weighted_*)
    OVD int combine(int p0, int p1, int nref, int bcw, int w1, int bd, int maxv)
    {
        int v;
        if (nref == 1)  v = (p0 + (1 << (13 - bd))) >> (14 - bd);
        else if (!bcw)  v = (p0 + p1 + (1 << (14 - bd))) >> (15 - bd);
        else            v = (p1 * w1 + p0 * (8 - w1) + (1 << (16 - bd))) >> (17 - bd);
        return ovd_clamp(v, 0, maxv);
    }
    // 8 horizontally adjacent final samples: two 16-byte reads of the exchange buffer, one 16-byte (or 8-byte) store
    // gpm != 0: the two predictions are blended with the geometric weights; (gx, gy) = luma position of the first sample
    // inside the CU, gstep = 1 (luma) or 2 (chroma samples the luma mask, rcn_inter.c:3090)
    OVD void store_row8(const Args &a, int plane, const int16_t *e0, const int16_t *e1, uint16_t *d, int n, int nref, int bcw, int w1,
                        int maxv, bool lut, int gpm = 0, int gx = 0, int gy = 0, int gstep = 1)
    {
        GpmGeom gg;
        if (gpm) gg = gpm_geom((gpm & 0xff) - 1, (gpm >> 24) & 15, (gpm >> 28) & 15);
        const ovd_u4 q0 = *(const ovd_u4 *)e0;
        ovd_u4 q1; q1.x = q1.y = q1.z = q1.w = 0;
        if (nref == 2) q1 = *(const ovd_u4 *)e1;
        const uint32_t w0[4] = { q0.x, q0.y, q0.z, q0.w }, w1v[4] = { q1.x, q1.y, q1.z, q1.w };
        uint32_t o[4];
        (void)plane;
        // one warp-uniform choice of the combination, then a straight loop: the per-sample code holds only that formula
        const int bd = a.bitdepth;
#define MC_ROW(EXPR_A, EXPR_B)                                                                     \
        _Pragma("unroll")                                                                          \
        for (int k = 0; k < 4; k++) {                                                              \
            const int a0 = (int16_t)(w0[k] & 0xffff), a1 = (int16_t)(w0[k] >> 16);                \
            const int b0 = (int16_t)(w1v[k] & 0xffff), b1 = (int16_t)(w1v[k] >> 16);              \
            (void)b0; (void)b1;                                                                    \
            int va = ovd_clamp(EXPR_A, 0, maxv), vb = ovd_clamp(EXPR_B, 0, maxv);                  \
            if (lut) { va = a.lmcs_fwd[va]; vb = a.lmcs_fwd[vb]; }                                 \
            o[k] = (uint32_t)va | ((uint32_t)vb << 16);                                            \
        }
        if (gpm) {              // put_weighted_gpm_bi_pixels: the BCW formula with a per-sample weight, w1 = 8 - w0
            MC_ROW((b0 * (8 - gpm_weight(gg, gx + (2 * k) * gstep, gy)) + a0 * gpm_weight(gg, gx + (2 * k) * gstep, gy) + (1 << (16 - bd))) >> (17 - bd),
                   (b1 * (8 - gpm_weight(gg, gx + (2 * k + 1) * gstep, gy)) + a1 * gpm_weight(gg, gx + (2 * k + 1) * gstep, gy) + (1 << (16 - bd))) >> (17 - bd))
        } else if (nref == 1) {
            MC_ROW((a0 + (1 << (13 - bd))) >> (14 - bd), (a1 + (1 << (13 - bd))) >> (14 - bd))
        } else if (!bcw) {
            MC_ROW((a0 + b0 + (1 << (14 - bd))) >> (15 - bd), (a1 + b1 + (1 << (14 - bd))) >> (15 - bd))
        } else {
            MC_ROW((b0 * w1 + a0 * (8 - w1) + (1 << (16 - bd))) >> (17 - bd), (b1 * w1 + a1 * (8 - w1) + (1 << (16 - bd))) >> (17 - bd))
        }
#undef MC_ROW
        // block positions are multiples of 4 luma samples only (ternary splits): pick the widest aligned store
        const unsigned al = (unsigned)(uintptr_t)d;
        if (n >= 8 && !(al & 15)) { ovd_u4 r; r.x = o[0]; r.y = o[1]; r.z = o[2]; r.w = o[3]; ovd_st16(d, r); }
        else if (n >= 4 && !(al & 7)) {
            ovd_u2 r; r.x = o[0]; r.y = o[1]; ovd_st8(d, r);
            if (n >= 8) { r.x = o[2]; r.y = o[3]; ovd_st8(d + 4, r); }
        } else {
            uint32_t *d32 = (uint32_t *)d;
            d32[0] = o[0];
            if (n >= 4) d32[1] = o[1];
            if (n >= 8) { d32[2] = o[2]; d32[3] = o[3]; }
        }
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
        int16_t *ex_y = (int16_t *)(win_c + SM_WIN_C);
        int16_t *ex_c = (int16_t *)(win_c + SM_WIN_C + SM_EX_Y);
        const int cw = a.pic_w >> 1, ch = a.pic_h >> 1;

        if (P == 0) {
            const Item it = load_item(a, bx * WARPS + warp);
            if (lane == 0) sit[0] = it.valid;
            if (!it.valid) return;
            const int do_y = it.flags & MC_LUMA, do_c = it.flags & MC_CHROMA;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (r >= it.nref) break;
                // ---- luma window: 28 samples from a 4-sample aligned start, h + 7 rows
                const int xs = it.bx + (it.mvx[r] >> 4) - 3, ys = it.by + (it.mvy[r] >> 4) - 3, x_al = xs & ~3;
                const int xc = (it.bx >> 1) + (it.mvx[r] >> 5) - 1, yc = (it.by >> 1) + (it.mvy[r] >> 5) - 1, xc_al = xc & ~3;
                if (lane == 0) {
                    int px = it.mvx[r] & 15, py = it.mvy[r] & 15;
                    if ((it.flags & MC_HPEL_AMVR) && !it.pu_small) { if (px == 8) px = 16; if (py == 8) py = 16; }   // rcn_inter.c:572-577
                    if (it.pu_small) { px += 17; py += 17; }                                                       // 6-tap set, rcn_mc.c:457
                    sit[8 + 4 * r] = (xs - x_al) | ((xc - xc_al) << 8);
                    sit[9 + 4 * r] = px | (py << 8);
                    sit[10 + 4 * r] = (it.mvx[r] & 31) | ((it.mvy[r] & 31) << 8);
                }
                // All global loads of a list are issued before the first shared-memory store so that they overlap
                // (the windows are latency-, not bandwidth-bound).
                const uint16_t *pl = a.refs.plane[it.list[r]][it.ridx[r]][0];
                const int pitch = a.refs.pitch[it.list[r]][it.ridx[r]][0], nrows = it.h + 7;
                uint32_t *w = win_y + r * WIN_Y_ROWS * WIN_Y_WORDS;
                const int fast_y = do_y && x_al >= 0 && x_al + 2 * WIN_Y_WORDS <= a.pic_w && ys >= 0 && ys + nrows <= a.pic_h;
                const int nrows_c = (it.h >> 1) + 3, comp = lane >> 4;
                const uint16_t *plc = a.refs.plane[it.list[r]][it.ridx[r]][1 + comp];
                const int pitch_c = a.refs.pitch[it.list[r]][it.ridx[r]][1 + comp];
                uint32_t *wc = win_c + (r * 2 + comp) * WIN_C_ROWS * WIN_C_WORDS;
                const int fast_c = do_c && xc_al >= 0 && xc_al + 2 * WIN_C_WORDS <= cw && yc >= 0 && yc + nrows_c <= ch;
                ovd_u2 vy[6], vc[3];
                const int cy = lane & 7, ry = lane >> 3, cc = lane & 3, rc = (lane >> 2) & 3;
                if (fast_y && cy < 7) {
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        if (ry + 4 * k < nrows) vy[k] = ovd_ld8(pl + (size_t)(ys + ry + 4 * k) * pitch + x_al + 4 * cy);
                }
                if (fast_c) {
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        if (rc + 4 * k < nrows_c) vc[k] = ovd_ld8(plc + (size_t)(yc + rc + 4 * k) * pitch_c + xc_al + 4 * cc);
                }
                if (fast_y && cy < 7) {
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        if (ry + 4 * k < nrows) { w[(ry + 4 * k) * WIN_Y_WORDS + 2 * cy] = vy[k].x; w[(ry + 4 * k) * WIN_Y_WORDS + 2 * cy + 1] = vy[k].y; }
                }
                if (fast_c) {
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        if (rc + 4 * k < nrows_c) { wc[(rc + 4 * k) * WIN_C_WORDS + 2 * cc] = vc[k].x; wc[(rc + 4 * k) * WIN_C_WORDS + 2 * cc + 1] = vc[k].y; }
                }
                if (do_y && !fast_y) {              // window crosses the picture border: clamped coordinates
                    for (int i = lane; i < nrows * WIN_Y_WORDS; i += 32) {
                        const int row = i / WIN_Y_WORDS, wd = i - row * WIN_Y_WORDS;
                        w[i] = load_pair(pl, pitch, a.pic_w, a.pic_h, x_al + 2 * wd, ys + row);
                    }
                }
                if (do_c && !fast_c) {
                    for (int i = lane & 15; i < nrows_c * WIN_C_WORDS; i += 16) {
                        const int row = i / WIN_C_WORDS, wd = i - row * WIN_C_WORDS;
                        wc[i] = load_pair(plc, pitch_c, cw, ch, xc_al + 2 * wd, yc + row);
                    }
                }
            }
            if (lane == 0) {
                sit[1] = it.bx; sit[2] = it.by; sit[3] = it.w; sit[4] = it.h; sit[5] = it.nref; sit[6] = it.flags; sit[7] = it.bcw;
                sit[11] = it.gpm;
            }
            return;
        }

        if (!sit[0]) return;
        const int ibx = sit[1], iby = sit[2], iw = sit[3], ih = sit[4], nref = sit[5], flags = sit[6], bcw = sit[7];
        const int do_y = flags & MC_LUMA, do_c = flags & MC_CHROMA;

        if (P == 1) {
            const int sh1 = a.bitdepth - 8, hi = lane >> 4, bi = nref == 2;
            const int r = bi ? hi : 0;
            const int offs = sit[8 + 4 * r], frl = sit[9 + 4 * r], frc = sit[10 + 4 * r];
            if (do_y) {
                const int x = lane & 15, row0 = bi ? 0 : 8 * hi;
                if (x < iw && row0 < ih) {
                    const int o = (offs & 0xff) + x;
                    uint32_t Bh[5], Bv0[4], Bv1[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) { Bh[j] = mc_cw_l[frl & 0xff][o & 1][j]; Bv1[j] = mc_cw_l[frl >> 8][1][j]; }
#pragma unroll
                    for (int j = 0; j < 4; j++) Bv0[j] = mc_cw_l[frl >> 8][0][j];
                    const uint32_t *w = win_y + (r * WIN_Y_ROWS + row0) * WIN_Y_WORDS + (o >> 1);
                    int16_t *ex = ex_y + (r * 16 + row0) * 16 + x;
                    if (bi) {
                        if (ih == 16)     column<16, 8, 5>(w, WIN_Y_WORDS, Bh, Bv0, Bv1, sh1, ex, 16);
                        else if (ih == 8) column<8, 8, 5>(w, WIN_Y_WORDS, Bh, Bv0, Bv1, sh1, ex, 16);
                        else              column<4, 8, 5>(w, WIN_Y_WORDS, Bh, Bv0, Bv1, sh1, ex, 16);
                    } else {
                        if (ih >= 8)      column<8, 8, 5>(w, WIN_Y_WORDS, Bh, Bv0, Bv1, sh1, ex, 16);
                        else              column<4, 8, 5>(w, WIN_Y_WORDS, Bh, Bv0, Bv1, sh1, ex, 16);
                    }
                }
            }
            if (do_c) {
                const int x = lane & 7, comp = (lane >> 3) & 1, row0 = bi ? 0 : 4 * hi, hc = ih >> 1;
                if (x < (iw >> 1) && row0 < hc) {
                    const int o = (offs >> 8) + x;
                    uint32_t Bh[3], Bv0[2], Bv1[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) { Bh[j] = mc_cw_c[frc & 0xff][o & 1][j]; Bv1[j] = mc_cw_c[frc >> 8][1][j]; }
#pragma unroll
                    for (int j = 0; j < 2; j++) Bv0[j] = mc_cw_c[frc >> 8][0][j];
                    const uint32_t *w = win_c + ((r * 2 + comp) * WIN_C_ROWS + row0) * WIN_C_WORDS + (o >> 1);
                    int16_t *ex = ex_c + ((r * 2 + comp) * 8 + row0) * 8 + x;
                    if (bi) {
                        if (hc == 8)      column<8, 4, 3>(w, WIN_C_WORDS, Bh, Bv0, Bv1, sh1, ex, 8);
                        else if (hc == 4) column<4, 4, 3>(w, WIN_C_WORDS, Bh, Bv0, Bv1, sh1, ex, 8);
                        else              column<2, 4, 3>(w, WIN_C_WORDS, Bh, Bv0, Bv1, sh1, ex, 8);
                    } else {
                        if (hc >= 4)      column<4, 4, 3>(w, WIN_C_WORDS, Bh, Bv0, Bv1, sh1, ex, 8);
                        else              column<2, 4, 3>(w, WIN_C_WORDS, Bh, Bv0, Bv1, sh1, ex, 8);
                    }
                }
            }
            return;
        }

        // ---- combination + store ----------------------------------------------------------------------------------
        const int maxv = (1 << a.bitdepth) - 1;
        const int w1 = bcw == 1 ? -2 : bcw == 2 ? 3 : bcw == 4 ? 5 : 10;           // rcn_inter.c:89
        const int gpm = sit[11], gox = (gpm >> 8) & 0xff, goy = (gpm >> 16) & 0xff;
        if (do_y) {
            // lane -> (row, 8-sample half): 16 rows x 2 halves
            const int y = lane >> 1, x = (lane & 1) * 8;
            if (y < ih && x < iw)
                store_row8(a, 0, ex_y + y * 16 + x, ex_y + (16 + y) * 16 + x, a.dst[0] + (size_t)(iby + y) * a.dst_pitch[0] + ibx + x,
                           iw - x, nref, bcw, w1, maxv, a.lmcs_fwd != 0, gpm, gox + x, goy + y, 1);
        }
        if (do_c) {
            // lanes 0-15 -> (component, row): 2 x 8 rows of <= 8 samples
            const int comp = (lane >> 3) & 1, y = lane & 7;
            if (lane < 16 && y < (ih >> 1)) {
                uint16_t *d = (comp ? a.dst[2] : a.dst[1]) + (size_t)((iby >> 1) + y) * (comp ? a.dst_pitch[2] : a.dst_pitch[1]) + (ibx >> 1);
                store_row8(a, 1, ex_c + (comp * 8 + y) * 8, ex_c + ((2 + comp) * 8 + y) * 8, d, iw >> 1, nref, bcw, w1, maxv, false, gpm, gox, goy + 2 * y, 2);
            }
        }
    }
};
