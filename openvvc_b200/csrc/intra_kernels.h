// intra_kernels.h -- luma + chroma intra prediction of a list of coding units (sm_100a; also compiled by tests/emu).
//
// Chroma (vvc_intra_pred_chroma, libovvc/rcn_intra.c:1015-1182, planar / DC / angular, no CCLM): 2-sample availability
// units (rcn_fill_ref.c:147-230, 392-480), no reference filtering, 2-tap linear interpolation (rcn_intra_angular.c:
// 270-470), PDPC only when both sides are >= 4 samples.  Grid y = component (0 luma, 1 Cb, 2 Cr).
//
// CIIP (rcn_ciip(_b), libovvc/rcn_inter.c:2968-3067): the MC stage has written the inter prediction into the picture;
// a command with the CIIP flag computes the planar prediction and blends it over those samples in place.
//
// Luma: vvc_intra_pred (libovvc/rcn_intra.c:484-560): reference sample fill with availability-driven
// substitution (libovvc/rcn_fill_ref.c:71-145, 320-390), [1 2 1] reference filter (:47-68), planar / DC with PDPC
// (libovvc/rcn_intra_dc_planar.c:151-236), wide-angle remapping and the angular dispatch (rcn_intra.c:55-481), angular
// kernels with cubic / Gaussian 4-tap interpolation and PDPC (libovvc/rcn_intra_angular.c:102-930).
//
// Intra blocks depend on the RECONSTRUCTED samples of earlier blocks, so the host sorts the blocks of a picture into
// dependency levels (a block's level = 1 + the highest level among the blocks its reference samples come from) and
// launches one grid per level; inside a level every block is independent.  One warp per block: phase 0 builds the
// two reference arrays in shared memory, phase 1 their filtered copies, phase 2 the negative-index extension, phase 3
// predicts -- each lane computes samples straight from the arrays with closed-form per-sample expressions (the
// horizontal class is the vertical class on swapped roles), so there is no transposition pass.
//
// MIP (rcn_intra_mip, libovvc/rcn_intra_mip.c:135-403): same reference fill; phase 1 averages the boundary to 2 / 4 samples
// per side, phase 2 does the matrix product (one lane per reduced sample, weights from mip_tables.h), phase 3 evaluates
// the two linear up-sampling passes in closed form per output sample.
#pragma once
#include "dev.h"
#include "mip_tables.h"

struct IntraCmdDev {          // == OVGPUIntraCmd
    int16_t x0, y0;
    uint8_t log2_w, log2_h, mode, mode_c, flags, mrl_idx, ciip_wt, pad[5];
    uint64_t avail_cols, avail_rows;
};

OVD_CONST int16_t intra_angle_tab[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29,
                                          32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
OVD_CONST int16_t intra_inv_angle_tab[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170, 1024, 910, 819, 712, 630, 565,
                                              512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
OVD_CONST uint8_t intra_filter_thr[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };
OVD_CONST uint8_t intra_wide_shift[6] = { 0, 6, 10, 12, 14, 15 };
OVD_CONST int8_t intra_cubic[32][4] = {
    { 0, 64, 0, 0 }, { -1, 63, 2, 0 }, { -2, 62, 4, 0 }, { -2, 60, 7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
    { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 }, { -6, 49, 24, -3 },
    { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 }, { -4, 36, 36, -4 }, { -4, 33, 39, -4 },
    { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 }, { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 },
    { -2, 16, 54, -4 }, { -2, 15, 55, -4 }, { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 },
    { 0, 4, 62, -2 }, { 0, 2, 63, -1 } };

struct IntraK {
    static const int THREADS = 128, MIN_BLOCKS = 4, PHASES = 4;
    static const int WARPS = 4;
    static const int REF_OFF = 72, REF_LEN = 256;                    // 64 + 2 negative-extension slots (MRL: two more), then 2 * 64 + 1 +
                                                                     // (m + 1) samples + up to 16 * (m + 1) of wide-angle padding
    static const int SM_REFS = 4 * REF_LEN * 2;                      // abv, lft, filtered abv, filtered lft (uint16)
    static const int SM_BLK = 256;                                   // the block's derived parameters (Blk), shared by the phases
    static const int SM_WARP = SM_REFS + SM_BLK;
    static const int SMEM_BYTES = WARPS * SM_WARP;

    struct Args {
        uint16_t *plane[3];
        int pitch[3], log2_ctu, bitdepth;
        const IntraCmdDev *cmds;
        const uint32_t *order;      // command indices of this level
        int nb;
        int comp0;                  // component of grid row 0 (luma levels: 0; chroma levels: 1, grid rows = Cb, Cr)
        // ISP (dataflow kernel only): the partition to predict inside the command's coding block and the two availability
        // words of this prediction call (vvc_intra_pred_isp, rcn_intra.c:586-596)
        int isp_ox, isp_oy, isp_lpw, isp_lph;
        uint64_t isp_cols, isp_rows;
        int warp_comps;             // 0, or (dataflow kernel, recon_kernels.h) the CTA owns ONE command: its threads split into warp_comps groups,
                                    // group k predicts component comp0 + k (1: luma, 2: Cb and Cr)
        int predecoded;             // the block parameters are already in shared memory (predecode(), dataflow kernel): phase 0 does not derive them
    };
    struct State { int dummy; };

    // everything the phases derive from the command (warp-uniform)
    struct Blk {
        int mip, mip_tr, mip_idx;                    // matrix-based prediction: transposed variant, matrix index
        int isp, cb_x0, cb_y0, cb_lw, cb_lh, ox, oy; // ISP partition: the coding block it belongs to and its offset inside
        int valid, chroma, pd, ciip, mrl, cclm, coll;  // cclm: 0 or the chroma mode 67 (LM) / 68 (MDLM left) / 69 (MDLM top)                 // pd: planar / DC use their PDPC variant; ciip: blend weight or 0
        int x0, y0, lw, lh, mode;                    // in samples of the component; mode as signalled: 0 planar, 1 DC, 2..66 angular
        int vert;                                    // angular: vertical class
        int mlw, mlh;                                // block size as seen by the class (swapped for the horizontal class)
        int mode_idx, a, angle, inv, frac, gauss, pdpc_scale, use_pdpc, filt_m, filt_s, ext;
        uint64_t cols, rows;
    };

    OVD int floor_log2(unsigned v) { return ovd_floor_log2(v); }
    OVD int pdpc_w(int scale, int pos) { const int s = (pos << 1) >> scale; return s > 5 ? 0 : 32 >> s; }

    OVD Blk load(const Args &a, int idx, int comp)
    {
        Blk b;
        b.valid = idx < a.nb;
        if (!b.valid) return b;
        const IntraCmdDev &c = a.cmds[a.order[idx]];
        b.chroma = comp > 0;
        b.ciip = (c.flags & 4) ? c.ciip_wt : 0;
        b.valid = b.chroma ? ((c.flags & 2) && !(b.ciip && c.log2_w <= 2)) : (c.flags & 1);
        if (!b.valid) return b;
        const int sh = b.chroma;
        b.x0 = c.x0 >> sh; b.y0 = c.y0 >> sh; b.lw = c.log2_w - sh; b.lh = c.log2_h - sh; b.cols = c.avail_cols; b.rows = c.avail_rows;
        b.isp = !b.chroma && (c.flags & 32) != 0;
        b.cb_x0 = b.x0; b.cb_y0 = b.y0; b.cb_lw = b.lw; b.cb_lh = b.lh; b.ox = b.oy = 0;
        if (b.isp) {
            b.ox = a.isp_ox; b.oy = a.isp_oy; b.x0 += b.ox; b.y0 += b.oy; b.lw = a.isp_lpw; b.lh = a.isp_lph;
            b.cols = a.isp_cols; b.rows = a.isp_rows;
        }
        b.mrl = (b.chroma || b.ciip || b.isp) ? 0 : c.mrl_idx;
        b.pd = b.chroma ? (b.lw > 1 && b.lh > 1) : (b.isp ? b.lh > 1 : !b.mrl);     // ISP: planar / DC PDPC iff log2_pb_h > 1 (rcn_intra.c:612-631)
        int mode = b.ciip ? 0 : (b.chroma ? c.mode_c : c.mode);
        // BDPCM block (pad[4]: bit 0 luma, bit 1 its direction, bit 2 chroma, bit 3 its direction, as OVGPUTuCmd.bdpcm): the pure
        // horizontal / vertical kernel on the plain references -- no PDPC, no filtering (rcn_intra.c:511-522, 1043-1074)
        const int bd = (b.ciip || b.isp || (!b.chroma && (c.flags & 16))) ? 0 : (b.chroma ? (c.pad[4] >> 2) & 3 : c.pad[4] & 3);
        if (bd & 1) mode = (bd & 2) ? 50 : 18;
        b.mip = !b.chroma && (c.flags & 16);
        b.mip_tr = (c.mode >> 7) & 1; b.mip_idx = c.mode & 0x3f;
        if (b.mip) mode = 1;                           // the phases below branch on b.mip first; DC keeps the derivation trivial
        b.cclm = mode >= 67 ? mode : 0;
        b.coll = (c.flags & 8) != 0;
        if (b.cclm) mode = 0;
        b.mode = mode;                                 // 0 planar, 1 DC, else angular (the remapped value lives in mode_idx)
        b.vert = 1; b.mlw = b.lw; b.mlh = b.lh;
        b.mode_idx = b.a = b.angle = b.inv = b.frac = b.gauss = b.pdpc_scale = b.use_pdpc = b.filt_m = b.filt_s = b.ext = 0;
        if (mode > 1) {
            const int d = b.cb_lw - b.cb_lh, ms = intra_wide_shift[d < 0 ? -d : d];            // derive_wide_angular_mode (ISP: the coding block's shape)
            if (b.cb_lw > b.cb_lh && mode < 2 + ms) mode += 65;
            else if (b.cb_lh > b.cb_lw && mode > 66 - ms) mode -= 65;
            b.vert = mode >= 34;
            b.mode_idx = b.vert ? mode - 50 : 18 - mode;
            if (!b.vert) { b.mlw = b.lh; b.mlh = b.lw; }
            b.a = b.mode_idx < 0 ? -b.mode_idx : b.mode_idx;
            b.gauss = !b.isp && !b.chroma && (b.lw + b.lh) > 5 && b.a > intra_filter_thr[(b.lw + b.lh) >> 1];
            if (b.isp) {
                // intra_angular_cubic_{v,h} (rcn_intra.c:288-452): cubic only, unfiltered references, and every PDPC decision
                // tests log2_pb_h of the block -- also in the horizontal class
                if (b.a == 0) b.use_pdpc = b.lh > 1;
                else if (b.mode_idx == 16) { b.use_pdpc = b.lh > 1; b.pdpc_scale = ovd_min(2, b.mlh - 2); }
                else {
                    b.angle = intra_angle_tab[b.a]; b.inv = intra_inv_angle_tab[b.a];
                    b.frac = (b.angle & 31) != 0;
                    b.pdpc_scale = ovd_min(2, b.mlh - (floor_log2((unsigned)(3 * b.inv - 2)) - 8));
                    if (b.mode_idx < 0) { b.ext = 1; b.angle = -b.angle; }
                    else if (!(b.pdpc_scale < 0 || b.lh < 2)) b.use_pdpc = 1;
                }
            } else if (b.mrl) {                                                                       // vvc_intra_pred_multi_ref: no filter, no PDPC
                b.gauss = 0;
                b.angle = b.a == 16 ? 32 : intra_angle_tab[b.a]; b.inv = intra_inv_angle_tab[b.a];
                b.frac = (b.angle & 31) != 0;
                if (b.mode_idx < 0) { b.ext = 1; b.angle = -b.angle; }
            } else if (b.a == 0) {
                b.use_pdpc = b.chroma ? b.pd : (b.gauss || b.mlh > 1);
            } else if (b.mode_idx == 16) {
                b.use_pdpc = b.chroma ? b.pd : (b.gauss || b.mlh > 1);
                b.pdpc_scale = ovd_min(2, b.mlh - 2);
                b.filt_m = b.filt_s = b.gauss;
            } else {
                b.angle = intra_angle_tab[b.a]; b.inv = intra_inv_angle_tab[b.a];
                b.frac = (b.angle & 31) != 0;
                b.pdpc_scale = ovd_min(2, b.mlh - (floor_log2((unsigned)(3 * b.inv - 2)) - 8));
                const int f = b.gauss && !b.frac;                                              // integer slope: filtered references
                if (b.mode_idx < 0) { b.ext = 1; b.filt_m = b.filt_s = f; b.angle = -b.angle; }
                else if (b.pdpc_scale < 0 || (!b.gauss && b.mlh < 2) || !b.pd) { b.filt_m = f; }
                else { b.filt_m = b.filt_s = f; b.use_pdpc = 1; }
            }
        } else if (mode == 0) {
            b.filt_m = b.filt_s = !b.chroma && !b.mrl && !b.isp && (b.lw + b.lh) > 5;                    // planar: rcn_intra.c:531-538
        }
        if (bd & 1) { b.use_pdpc = 0; b.gauss = 0; b.filt_m = b.filt_s = 0; b.mrl = 0; }
        return b;
    }

    // fill_ref_left_0 / fill_ref_above_0 (offset 0): one lane per sample
    OVD void fill_ref(uint16_t *ref, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                      uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int lane, int nl, int us, int left_chroma, int off = 0)
    {
        // off (ISP, rcn_fill_ref.c:71-145 / 320-390 with offset_y / offset_x): the array covers the whole coding block and starts
        // `off` samples before the partition; the corner sample, the "nothing available" value and the padding length move with it
        // us = log2 of the availability unit in samples of this plane (2 luma, 1 chroma)
        const int nb_units = ((1 << (log2_len + 1)) >> us) + 1, ref_len = (1 << (log2_len + 1)) + 1;
        const uint64_t mask = (1ull << (nb_units + 1)) - 1, avl = (map >> (pos >> us)) & mask;
        int nb_avl = 0;
        for (uint64_t t = avl; t; t >>= 1) nb_avl++;                                           // 64 - clz
        int pad = 1 << (bitdepth - 1);
        if (!avl) {
            const int nb_units_o = ((1 << (log2_other + 1)) >> us) + 1;
            const uint64_t avl_o = (other_map >> (other_pos >> us)) & ((1ull << (nb_units_o + 1)) - 1);
            if (avl_o) pad = p[other_step + (ptrdiff_t)off * step];
        } else if (nb_avl > 1) pad = p[(size_t)((nb_avl - 1) << us) * step];
        else if (!left_chroma) pad = p[(ptrdiff_t)off * step];                                 // only the corner unit: its sample
        // (fill_ref_left_0_chroma keeps the mid value there, rcn_fill_ref.c:193-200)
        for (int i = lane; i < ref_len + 4 + off; i += nl) {
            int v = pad;
            const int k = ovd_min(i, ref_len - 1);                                             // wide-angle padding repeats the last sample
            if (avl) {
                if (k == 0) v = avl == mask ? p[0] : ((avl & 1) ? p[(ptrdiff_t)off * step] : p[step]);   // (everything available: plain copy)
                else if (1 + ((k - 1) >> us) < nb_avl) v = p[(size_t)k * step];
            }
            ref[i] = (uint16_t)v;
        }
    }

    // fill_ref_left_0_mref / fill_ref_above_0_mref (rcn_fill_ref.c:228-318, 493-590): raw[0] = the sample m + 1 away both ways,
    // corner group of m + 1 samples, then 4-sample units; left and above differ in two places (rcn_fill_ref.c mref variants)
    OVD void fill_ref_mref(uint16_t *raw, const uint16_t *p, ptrdiff_t step, ptrdiff_t other_step, int log2_len, int log2_other,
                           uint64_t map, uint64_t other_map, int pos, int other_pos, int bitdepth, int lane, int nl, int m, int is_left)
    {
        const int nb_units = ((1 << (log2_len + 1)) >> 2) + 1, ref_len = (1 << (log2_len + 1)) + 1 + (m + 1);
        const int ratio = ovd_max(1, (1 << log2_len) >> log2_other);
        const uint64_t mask = (1ull << (nb_units + 1)) - 1, avl = (map >> (pos >> 2)) & mask;
        int nb_avl = 0;
        for (uint64_t t = avl; t; t >>= 1) nb_avl++;
        const int first = (avl & 1) ? 0 : (is_left ? 1 : m + 1);                               // replicated corner-group sample
        const int grp_off = (is_left && !(avl & 1)) ? 1 : m + 1;                               // p index of the first unit sample
        int pad = 1 << (bitdepth - 1);
        if (!avl) {
            const int nb_units_o = ((1 << (log2_other + 1)) >> 2) + 1;
            const uint64_t avl_o = (other_map >> (other_pos >> 2)) & ((1ull << (nb_units_o + 1)) - 1);
            if (avl_o) pad = is_left ? p[(size_t)(1 + m) * other_step] : p[(ptrdiff_t)m * step + (ptrdiff_t)(m + 1) * other_step];
        } else if (nb_avl > 1) pad = p[(ptrdiff_t)(grp_off + 4 * (nb_avl - 1) - 1) * step];
        else if (!is_left) pad = p[(ptrdiff_t)m * step];
        for (int i = lane; i < ref_len - 1 + ratio * (m + 1); i += nl) {
            int v = pad;
            const int k = ovd_min(i, ref_len - 2);         // the reference's wide-angle padding starts on the last sample (:312-316)
            if (avl) {
                if (k < m + 1) v = (avl & 1) ? p[(ptrdiff_t)k * step] : p[(ptrdiff_t)first * step];
                else if (1 + ((k - (m + 1)) >> 2) < nb_avl) v = p[(ptrdiff_t)(grp_off + k - (m + 1)) * step];
            }
            raw[i] = (uint16_t)v;
        }
    }

    // CCLM luma down-sampling at luma (2i, 2j): 6-tap over two rows, or the 5-tap cross when chroma is vertically collocated
    // (compute_lm_subsample(_cl), rcn_intra_cclm.c:297-324, 626-657)
    OVD int lm_luma(const uint16_t *y, ptrdiff_t st, int coll, int pad_left, int pad_abv)
    {
        const int l = pad_left ? 0 : -1;
        if (coll) return (y[pad_abv ? 0 : -st] + 4 * y[0] + y[l] + y[1] + y[st] + 4) >> 3;
        return (y[l] + 2 * y[0] + y[1] + y[st + l] + 2 * y[st] + y[st + 1] + 4) >> 3;
    }
    OVD int cclm_abv(const Blk &b, int cm) { return ((b.rows >> (((b.x0 & cm) >> 1) + 1)) & ((1ull << ((1 << b.lw) >> 1)) - 1)) != 0; }
    OVD int cclm_lft(const Blk &b, int cm) { return ((b.cols >> (((b.y0 & cm) >> 1) + 1)) & ((1ull << ((1 << b.lh) >> 1)) - 1)) != 0; }
    OVD int ctz64_ones(uint64_t v) { int n = 0; while (v & 1) { v >>= 1; n++; } return n; }       // ctz(~v)

    // linear-model parameters of one chroma component (one lane): neighbour selection (intra_cclm / intra_mdlm_*,
    // :659-883), min / max pairing (:137-186), a / b / shift (:72-135)
    OVD void cclm_params(const Blk &b, const uint16_t *Y, ptrdiff_t sl, const uint16_t *C, ptrdiff_t sc, int cm, int bitdepth, int *prm)
    {
        const int w = 1 << b.lw, h = 1 << b.lh, xl = b.x0 & cm, yl = b.y0 & cm;
        const int abv = cclm_abv(b, cm), lft = cclm_lft(b, cm);
        int nb_abv = 0, nb_lft = 0, abv_step = 1, lft_step = 1;
        if (b.cclm == 67) {
            if (abv) { nb_abv = ovd_min(w, (abv + !lft) << 1); abv_step = w >> (abv + !lft); }
            if (lft) { nb_lft = ovd_min(h, (lft + !abv) << 1); lft_step = h >> (lft + !abv); }
        } else if (b.cclm == 69) {
            if (abv) {
                const int nb_units = (w + ovd_min(h, w)) >> 1;
                const int len = ctz64_ones((b.rows >> ((xl >> 1) + 1)) & ((1ull << nb_units) - 1)) << 1;
                nb_abv = ovd_min(len, 4); abv_step = len >> 2;
            }
        } else if (lft) {
            const int nb_units = (h + ovd_min(h, w)) >> 1;
            const int len = ctz64_ones((b.cols >> ((yl >> 1) + 1)) & ((1ull << nb_units) - 1)) << 1;
            nb_lft = ovd_min(len, 4); lft_step = len >> 2;
        }
        abv_step = ovd_max(abv_step, 1); lft_step = ovd_max(lft_step, 1);
        int sy[4] = { 0, 0, 0, 0 }, scv[4] = { 0, 0, 0, 0 };
        const int n = nb_abv + nb_lft;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < nb_abv) {
                const int x = (abv_step >> 1) + i * abv_step, pad_left = x == 0 && !lft;
                if (yl == 0) { const uint16_t *r = Y - sl + 2 * x; sy[i] = (r[pad_left ? 0 : -1] + 2 * r[0] + r[1] + 2) >> 2; }
                else sy[i] = lm_luma(Y - 2 * sl + 2 * x, sl, b.coll, pad_left, 0);
                scv[i] = C[-sc + x];
            } else if (i < n) {
                const int y = (lft_step >> 1) + (i - nb_abv) * lft_step;
                sy[i] = lm_luma(Y + (ptrdiff_t)(2 * y) * sl - 2, sl, b.coll, 0, y == 0 && !abv);
                scv[i] = C[(ptrdiff_t)y * sc - 1];
            }
        }
        prm[0] = 0; prm[1] = 1 << (bitdepth - 1); prm[2] = 0;
        if (!n) return;
        int min_l, max_l, min_c, max_c;
        if (n == 2) {
            const int sw = sy[0] >= sy[1];
            min_l = sw ? sy[1] : sy[0]; max_l = sw ? sy[0] : sy[1];
            min_c = sw ? scv[1] : scv[0]; max_c = sw ? scv[0] : scv[1];
        } else {
            // the reference's 4-element sorting network on (luma, chroma) pairs: {mn0, mn1} = the two smaller, {mx0, mx1} the larger
            int l0 = sy[0], l1 = sy[2], l2 = sy[1], l3 = sy[3], c0 = scv[0], c1 = scv[2], c2 = scv[1], c3 = scv[3], t;
            if (sy[0] > sy[2]) { t = l0; l0 = l1; l1 = t; t = c0; c0 = c1; c1 = t; }
            if (sy[1] > sy[3]) { t = l2; l2 = l3; l3 = t; t = c2; c2 = c3; c3 = t; }
            if (l0 > l3) { t = l0; l0 = l2; l2 = t; t = l1; l1 = l3; l3 = t; t = c0; c0 = c2; c2 = t; t = c1; c1 = c3; c3 = t; }
            if (l1 > l2) { t = l1; l1 = l2; l2 = t; t = c1; c1 = c2; c2 = t; }
            min_l = (l0 + l1 + 1) >> 1; max_l = (l2 + l3 + 1) >> 1;
            min_c = (c0 + c1 + 1) >> 1; max_c = (c2 + c3 + 1) >> 1;
        }
        prm[0] = 0; prm[1] = min_c; prm[2] = 0;
        const int range_l = max_l - min_l;
        if (!range_l) return;
        int l2r = floor_log2((unsigned)range_l);
        const int norm = ((range_l << 4) >> l2r) & 15;
        const int v = (int)((0x0111122334455670ull >> (4 * norm)) & 15) | 8;     // div_lut {0,7,6,5,5,4,4,3,3,2,2,1,1,1,1,0}
        l2r += norm != 0;
        const int range_c = max_c - min_c;
        const int l2c = range_c ? floor_log2((unsigned)(range_c < 0 ? -range_c : range_c)) + 1 : 0;
        int ac = (range_c * v + ((1 << l2c) >> 1)) >> l2c, shift = 3 + l2r - l2c;
        if (shift < 1) { shift = 1; ac = ac ? (ac < 0 ? -15 : 15) : 0; }
        prm[0] = ac; prm[1] = min_c - ((ac * min_l) >> shift); prm[2] = shift;
    }

    // one sample of the vertical-class prediction at (x along the main reference, y across)
    OVD int class_pixel(const Blk &b, const uint16_t *M, const uint16_t *S, int tl, int x, int y, int maxv)
    {
        const int W = 1 << b.mlw;
        if (b.mrl) {                                                    // intra_angular_{v,h}_cubic_mref and the nofrac kernels
            if (b.a == 0) return M[x + 1];
            const int delta = (y + 1 + b.mrl) * b.angle, di = delta >> 5, df = delta & 31;
            const uint16_t *r = M + x + di;
            if (!b.frac) return r[1];
            return ovd_clamp((r[0] * intra_cubic[df][0] + r[1] * intra_cubic[df][1] + r[2] * intra_cubic[df][2] + r[3] * intra_cubic[df][3] + 32) >> 6, 0, maxv);
        }
        if (b.a == 0) {
            if (!b.use_pdpc) return M[x + 1];
            return ovd_clamp((pdpc_w((b.lw + b.lh - 2) >> 2, x) * (S[y + 1] - tl) + (M[x + 1] << 6) + 32) >> 6, 0, maxv);
        }
        if (b.mode_idx == 16) {
            int v = M[x + y + 2];
            if (b.use_pdpc && x < ovd_min(3 << b.pdpc_scale, W))
                v = ovd_clamp(v + (((32 >> ((2 * x) >> b.pdpc_scale)) * (S[y + x + 2] - v) + 32) >> 6), 0, maxv);
            return v;
        }
        const int delta = (y + 1) * b.angle, di = delta >> 5, df = delta & 31;
        int v;
        if (!b.frac) v = M[x + di + 1];
        else {
            const uint16_t *r = M + x + di;
            if (b.chroma) v = ovd_clamp(r[1] + ((df * (r[2] - r[1]) + 16) >> 5), 0, maxv);     // intra_angular_v_c
            else if (b.gauss) v = (r[0] * (16 - (df >> 1)) + r[1] * (32 - (df >> 1)) + r[2] * (16 + (df >> 1)) + r[3] * (df >> 1) + 32) >> 6;
            else v = ovd_clamp((r[0] * intra_cubic[df][0] + r[1] * intra_cubic[df][1] + r[2] * intra_cubic[df][2] + r[3] * intra_cubic[df][3] + 32) >> 6, 0, maxv);
        }
        if (b.use_pdpc && x < ovd_min(3 << b.pdpc_scale, W)) {
            const int wL = 32 >> ((x << 1) >> b.pdpc_scale), left = S[y + ((256 + (x + 1) * b.inv) >> 9) + 1];
            v = ovd_clamp(v + ((wL * (left - v) + 32) >> 6), 0, maxv);
        }
        return v;
    }

    // The derivation alone, for callers that can run it ahead of time: the dataflow kernel derives a command's parameters BEFORE it
    // waits for the command's neighbours (they depend on the command only), which takes the derivation off the dependency chain.
    // Followed by a barrier, then the phases with Args::predecoded set.
    OVD void predecode(const Args &a, unsigned char *smem, int tid, int bx, int by)
    {
        const int nl = a.warp_comps ? THREADS / a.warp_comps : 32;
        const int warp = tid / nl, lane = tid - warp * nl;
        if (lane) return;
        const int comp = a.warp_comps ? a.comp0 + warp : a.comp0 + by;
        Blk *const sblk = (Blk *)(smem + warp * SM_WARP + SM_REFS);
        const Blk b = load(a, a.warp_comps ? bx : bx * WARPS + warp, comp);
        if (b.valid) *sblk = b; else sblk->valid = 0;
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)bz;
        // a worker = the group of `nl` threads that owns one block: a warp (one launch per level), or -- in the dataflow kernel,
        // where the CTA owns ONE command and latency is what counts -- the whole CTA (luma) / one half each for Cb and Cr
        const int nl = a.warp_comps ? THREADS / a.warp_comps : 32;
        const int warp = tid / nl, lane = tid - warp * nl;
        const int comp = a.warp_comps ? a.comp0 + warp : a.comp0 + by;
        // The parameters every phase needs are derived from the command ONCE, in phase 0, and left in shared memory for the others:
        // the derivation is a few hundred dependent scalar instructions, and in the dataflow kernel it sat four times on every
        // block's critical path (measured: 7.9 k of the 12 k cycles a block adds to the dependency chain).
        Blk *const sblk = (Blk *)(smem + warp * SM_WARP + SM_REFS);
        Blk b;
        if (P == 0 && !a.predecoded) {
            b = load(a, a.warp_comps ? bx : bx * WARPS + warp, comp);
            if (lane == 0) { if (b.valid) *sblk = b; else sblk->valid = 0; }
        } else b = *sblk;
        if (!b.valid) return;
        uint16_t *plane = comp == 0 ? a.plane[0] : (comp == 1 ? a.plane[1] : a.plane[2]);
        const int pitch = comp == 0 ? a.pitch[0] : (comp == 1 ? a.pitch[1] : a.pitch[2]);
        uint16_t *base = (uint16_t *)(smem + warp * SM_WARP);
        // ISP: the reference arrays are filled for the whole coding block (phase 0) and then read from the partition's offset on
        uint16_t *abv = base + REF_OFF + b.ox, *lft = base + REF_LEN + REF_OFF + b.oy, *fabv = base + 2 * REF_LEN + REF_OFF, *flft = base + 3 * REF_LEN + REF_OFF;
        const int w = 1 << b.lw, h = 1 << b.lh, maxv = (1 << a.bitdepth) - 1, cm = (1 << (a.log2_ctu - b.chroma)) - 1;
        const int us = 2 - b.chroma;
        uint16_t *dst = plane + (size_t)b.y0 * pitch + b.x0;
        if (b.cclm) {                                                       // parameters by lane 0 in phase 0, prediction in phase 3
            int *prm = (int *)base;
            const uint16_t *Y = a.plane[0] + (size_t)(2 * b.y0) * a.pitch[0] + 2 * b.x0;
            if (P == 0 && lane == 0) cclm_params(b, Y, a.pitch[0], dst, pitch, cm, a.bitdepth, prm);
            if (P == 3) {
                const int abv = cclm_abv(b, cm), lft = cclm_lft(b, cm);
                const int ac = prm[0], bc = prm[1], shift = prm[2];
                for (int o = lane; o < w * h; o += nl) {
                    const int x = o & (w - 1), y = o >> b.lw;
                    const int lm = lm_luma(Y + (ptrdiff_t)(2 * y) * a.pitch[0] + 2 * x, a.pitch[0], b.coll, x == 0 && !lft, y == 0 && !abv);
                    dst[(size_t)y * pitch + x] = (uint16_t)ovd_clamp(((lm * ac) >> shift) + bc, 0, maxv);
                }
            }
            return;
        }

        if (b.mip && P > 0) {
            // vvc_intra_pred_mip(_tr): nb boundary samples per side, reduced block red x red
            const int nb = (b.lw > 2 || b.lh > 2) ? 4 : 2, log2_nb = nb == 4 ? 2 : 1;
            const int red_size = b.lh == 2 || b.lw == 2 || (b.lh <= 3 && b.lw <= 3), log2_red = red_size ? 2 : 3, red = 1 << log2_red;
            int *in_raw = (int *)(base + 2 * REF_LEN);
            uint16_t *pred = base + 3 * REF_LEN;
            if (P == 1) {
                if (lane < 2 * nb) {                                        // rcn_intra_mip.c:165-183 / :283-300
                    const int j = lane & (nb - 1), is_abv = (lane < nb) != b.mip_tr;
                    const int lb = (is_abv ? b.lw : b.lh) - log2_nb;
                    const uint16_t *r = (is_abv ? abv : lft) + 1 + (j << lb);
                    int sum = 0;
                    for (int i = 0; i < (1 << lb); i++) sum += r[i];
                    in_raw[lane] = (sum + ((1 << lb) >> 1)) >> lb;
                }
                return;
            }
            if (P == 2) {
                int in[8], sum = 0;
                const int input_offset = in_raw[0];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    int v = i < 2 * nb ? in_raw[i] : input_offset;
                    if (i == 0 && red_size) v = 1 << (a.bitdepth - 1);     // :187-189
                    in[i] = (int16_t)(v - input_offset);
                    sum += in[i];
                }
                const int rnd = 32 - 32 * sum;                              // MIP_OFFSET - MIP_OFFSET * sum
                const uint8_t *mt = (b.lw == 2 && b.lh == 2) ? &mip_w_4x4[b.mip_idx * 64] : (red_size ? &mip_w_8x8[b.mip_idx * 128] : &mip_w_16x16[b.mip_idx * 512]);
                for (int pos = lane; pos < red * red; pos += nl) {
                    const uint8_t *m = mt + pos * 2 * nb;
                    int v = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) if (i < 2 * nb) v += in[i] * m[i];
                    // the transposed variant reads the product as pred[j * red + i] (:337-341, :377-381): store it transposed
                    const int q = b.mip_tr ? ((pos & (red - 1)) << log2_red) + (pos >> log2_red) : pos;
                    pred[q] = (uint16_t)ovd_clamp(((v + rnd) >> 6) + input_offset, 0, maxv);
                }
                return;
            }
            // up_sample (:93-131) horizontally on the rows holding the reduced lines, then vertically, per output sample
            const int lsx = b.lw - log2_red, lsy = b.lh - log2_red, sx = 1 << lsx, sy = 1 << lsy;
            for (int o = lane; o < w * h; o += nl) {
                const int x = o & (w - 1), y = o >> b.lw;
                const int jx = x >> lsx, px = (x & (sx - 1)) + 1, jy = y >> lsy, py = (y & (sy - 1)) + 1;
                // H(i): value of reduced row i at column x
                int hv[2];
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int i = jy - 1 + t;
                    if (i < 0) { hv[t] = abv[1 + x]; continue; }
                    const int after = pred[(i << log2_red) + jx];
                    const int before = jx ? pred[(i << log2_red) + jx - 1] : lft[(i + 1) * sy];
                    hv[t] = lsx ? (before * (sx - px) + after * px + (sx >> 1)) >> lsx : after;
                }
                dst[(size_t)y * pitch + x] = (uint16_t)(lsy ? (hv[0] * (sy - py) + hv[1] * py + (sy >> 1)) >> lsy : hv[1]);
            }
            return;
        }

        if (b.isp && P < 2) {
            if (P == 0) {           // vvc_intra_pred_isp (rcn_intra.c:586-596): left column x0 - 1 from the block's top, above row y0 - 1 from its left
                fill_ref(lft - b.oy, plane + (size_t)(b.cb_y0 - 1) * pitch + b.x0 - 1, pitch, 1, b.cb_lh, b.cb_lw, b.cols, b.rows,
                         b.cb_y0 & cm, b.x0 & cm, a.bitdepth, lane, nl, 2, 0, b.oy);
                fill_ref(abv - b.ox, plane + (size_t)(b.y0 - 1) * pitch + b.cb_x0 - 1, 1, pitch, b.cb_lw, b.cb_lh, b.rows, b.cols,
                         b.cb_x0 & cm, b.y0 & cm, a.bitdepth, lane, nl, 2, 0, b.ox);
            } else {                // :601-607: four more samples replicated from index (block size + partition size)
                if (lane < 4) lft[(1 << b.cb_lh) + h + 1 + lane] = lft[(1 << b.cb_lh) + h];
                else if (lane < 8) abv[(1 << b.cb_lw) + w + 1 + (lane - 4)] = abv[(1 << b.cb_lw) + w];
            }
            return;
        }
        if (P == 0) {
            const uint16_t *corner = dst - (ptrdiff_t)(b.mrl + 1) * pitch - (b.mrl + 1);
            if (b.mrl) {                                                    // abv / lft = raw + m: [1 + i] is the i-th sample of line m
                fill_ref_mref(lft - b.mrl, corner, pitch, 1, b.lh, b.lw, b.cols, b.rows, b.y0 & cm, b.x0 & cm, a.bitdepth, lane, nl, b.mrl, 1);
                fill_ref_mref(abv - b.mrl, corner, 1, pitch, b.lw, b.lh, b.rows, b.cols, b.x0 & cm, b.y0 & cm, a.bitdepth, lane, nl, b.mrl, 0);
                return;
            }
            fill_ref(lft, corner, pitch, 1, b.lh, b.lw, b.cols, b.rows, b.y0 & cm, b.x0 & cm, a.bitdepth, lane, nl, us, b.chroma);
            fill_ref(abv, corner, 1, pitch, b.lw, b.lh, b.rows, b.cols, b.x0 & cm, b.y0 & cm, a.bitdepth, lane, nl, us, 0);
            return;
        }
        if (P == 1) {
            // filter_ref_samples over 2 * len samples of both arrays (the consumers read inside that range)
            for (int i = lane; i <= 2 * w; i += nl)
                fabv[i] = (uint16_t)(i == 0 ? (lft[1] + (abv[0] << 1) + abv[1] + 2) >> 2 : (i == 2 * w ? abv[i] : (abv[i + 1] + (abv[i] << 1) + abv[i - 1] + 2) >> 2));
            for (int i = lane; i <= 2 * h; i += nl)
                flft[i] = (uint16_t)(i == 0 ? (abv[1] + (lft[0] << 1) + lft[1] + 2) >> 2 : (i == 2 * h ? lft[i] : (lft[i + 1] + (lft[i] << 1) + lft[i - 1] + 2) >> 2));
            return;
        }
        // roles of the arrays for the block's class
        uint16_t *M = b.vert ? (b.filt_m ? fabv : abv) : (b.filt_m ? flft : lft);
        const uint16_t *S = b.vert ? (b.filt_s ? flft : lft) : (b.filt_s ? fabv : abv);
        if (P == 2) {
            if (b.ext) {                                                    // rcn_intra.c:118-121
                const int H = 1 << b.mlh;
                for (int k = 1 + lane; k <= H; k += nl) M[-b.mrl - k] = S[-b.mrl + ovd_min((256 + k * b.inv) >> 9, H)];
            }
            return;
        }
        if (P == 3) {
            const int tl = abv[0];
            if (b.mode == 0) {                                              // intra_planar_pdpc
                const uint16_t *A = b.filt_m ? fabv : abv, *L = b.filt_m ? flft : lft;
                const int ws = b.pd ? ovd_max(b.lw, 1) : b.lw, hs = b.pd ? ovd_max(b.lh, 1) : b.lh, scale = (b.lw + b.lh - 2) >> 2;
                for (int o = lane; o < w * h; o += nl) {
                    const int x = o & (w - 1), y = o >> b.lw;
                    const int hor = (L[y + 1] << ws) + (x + 1) * (A[w + 1] - L[y + 1]), ver = (A[x + 1] << hs) + (y + 1) * (L[h + 1] - A[x + 1]);
                    const int v = ((hor << hs) + (ver << ws) + (1 << (ws + hs))) >> (ws + hs + 1);
                    const int xw = pdpc_w(scale, x), yw = pdpc_w(scale, y);
                    int r = b.pd ? ovd_clamp((xw * L[y + 1] + yw * A[x + 1] + (64 - (xw + yw)) * v + 32) >> 6, 0, maxv) : v;
                    uint16_t *d = &dst[(size_t)y * pitch + x];
                    if (b.ciip) r = ovd_clamp((r * b.ciip + *d * (4 - b.ciip) + 2) >> 2, 0, maxv);     // put_weighted_ciip_pixels
                    *d = (uint16_t)r;
                }
            } else if (b.mode == 1) {                                       // intra_dc_pdpc
                const int shift = ovd_max(b.lw, b.lh) + (b.lw == b.lh), scale = (b.lw + b.lh - 2) >> 2;
                int dc = 0;
                if (b.lw >= b.lh) for (int i = 0; i < w; i++) dc += abv[1 + i];
                if (b.lw <= b.lh) for (int i = 0; i < h; i++) dc += lft[1 + i];
                dc = (dc + ((1 << shift) >> 1)) >> shift;
                for (int o = lane; o < w * h; o += nl) {
                    const int x = o & (w - 1), y = o >> b.lw;
                    const int tw = pdpc_w(scale, x), lwg = pdpc_w(scale, y);
                    dst[(size_t)y * pitch + x] = (uint16_t)(b.pd ? ovd_clamp((tw * lft[y + 1] + lwg * abv[x + 1] + (64 - (tw + lwg)) * dc + 32) >> 6, 0, maxv) : dc);
                }
            } else {
                for (int o = lane; o < w * h; o += nl) {
                    const int x = o & (w - 1), y = o >> b.lw;
                    const int v = b.vert ? class_pixel(b, M, S, tl, x, y, maxv) : class_pixel(b, M, S, tl, y, x, maxv);
                    dst[(size_t)y * pitch + x] = (uint16_t)v;
                }
            }
        }
    }
};
