// tr_kernels.h -- residual path of a whole picture's transform units (sm_100a; also compiled by tests/emu).
//
// Reference behaviour: tmp.rcn_tu_st / tmp.rcn_tu_c (libovvc/rcn_transform_tree.c:1228-1390) -> rcn_residual
// (:415-506), rcn_residual_c (:553-628), rcn_res_c (:720-793), rcn_jcbcr (:796-867), transform skip (:691-716,
// :1207-1225); dequantisation + sub-block de-swizzle (libovvc/rcn_dequant.c:92-312); LFNST (libovvc/drv_lfnst.c:
// 50-156, libovvc/rcn_lfnst.c:41-162); separable inverse transforms (libovvc/rcn_transform.c:45-598); residual
// add / joint-CbCr / LMCS chroma scaling (libovvc/rcn_residuals.c:46-331).
//
// One work item = one transform block (luma, Cb, Cr or a joint Cb-Cr block).  Blocks of <= 256 samples are owned
// by one warp, larger ones by a whole CTA (template parameter GROUP).  Phases:
//   0  coefficients (reference buffer order) -> dequantised block, stored TRANSPOSED in shared memory so that the
//      vertical pass reads packed pairs; positions covered by LFNST are produced by the LFNST kernel instead
//   1  vertical 1-D transform  (shift 7)           as IDP.2A dot products on packed int16 pairs x int8 matrix words
//   2  horizontal 1-D transform (shift 20 - BD)
//   3  residual add (+ joint-CbCr variants, chroma scaling) into the picture, clipped to the bit depth
// Only the rows / columns that can be non-zero (sig_sb_map, LFNST region) enter the dot products, like the
// reference's nb_row / nb_col skipping.
#pragma once
#include "dev.h"
#include "vvc_tr_tables.h"

#define TU_IMPLICIT_MTS 1
#define TU_LFNST_L 2
#define TU_LFNST_C 4
#define TU_SCALE_ANCHOR 8
#define TU_ISP 16            // luma block of an ISP partition: any size from 1xN up, explicit DST-VII / DCT-II choice, no DC shortcut

struct TuCmdDev {            // == OVGPUTuCmd
    int16_t x0, y0;
    uint8_t log2_w, log2_h, cbf_mask, tr_skip_mask;
    uint8_t cu_mts_flag, cu_mts_idx, flags, lfnst_idx;
    uint8_t qp[4], qp_ts[4];
    uint16_t last_pos[3];
    int16_t chroma_scale;
    uint8_t lfnst_mode_l, lfnst_mode_c;
    uint16_t scale_anchor;
    uint32_t coeff_off[3];
    uint32_t bdpcm;          // bit 0 luma BDPCM, bit 1 its direction, bit 2 chroma BDPCM, bit 3 its direction
    uint64_t sig_sb_map[3];
};

struct TbItemDev {           // one transform block of command `cmd`
    uint32_t cmd;
    uint32_t kind;           // 0 luma, 1 Cb, 2 Cr, 3 joint Cb-Cr
};

OVD_CONST uint8_t tr_iq_scale[2][6] = { { 40, 45, 51, 57, 64, 72 }, { 57, 64, 72, 80, 90, 102 } };
OVD_CONST uint8_t tr_lfnst_scan[16] = { 0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15 };   // rcn_lfnst.c:47
OVD_CONST uint8_t tr_wide_shift[6] = { 0, 6, 10, 12, 14, 15 };                                    // drv_lfnst.c:53

#ifndef TR_MIN_BLOCKS
#define TR_MIN_BLOCKS 8
#endif
template <int GROUP>
struct TrK {
    static const int THREADS = 128, MIN_BLOCKS = TR_MIN_BLOCKS, PHASES = 4;       // <= 64 registers: the line-tiled passes would otherwise take 80+
    static const int GROUPS = THREADS / GROUP;
    // shared memory per group, int16 units (see the size derivation in DESIGN.md "Residual")
    static const int SM_CT = GROUP == 32 ? 320 : 32 * 34;     // dequantised block, transposed: [x][k = y], line stride kv4 + 2
    static const int SM_T2 = GROUP == 32 ? 384 : 64 * 34;     // after the vertical pass: [y][k = x], line stride kh4 + 2
    static const int SM_RES = GROUP == 32 ? 256 : 64 * 64;    // residual block, raster
    static const int SM_TB = 256;                             // bytes: the block's derived parameters (Tb), shared by the phases
    static const int SM_GROUP = (SM_CT + SM_T2 + SM_RES) * 2 + SM_TB;
    static const int SMEM_BYTES = GROUPS * SM_GROUP;

    struct Args {
        uint16_t *dst[3];
        int dst_pitch[3];        // samples
        int bitdepth;
        const TuCmdDev *cmds;
        const TbItemDev *items;
        int nb_items;
        const int16_t *arena;
        int dep_quant, sh_ts_disabled, ict_type;
        const int16_t *scales;   // LMCS chroma scale per anchor (LmcsScaleK), for commands with TU_SCALE_ANCHOR
    };
    struct State { int dummy; };

    enum { MODE_TR = 0, MODE_DC = 1, MODE_TS = 2 };

    struct Tb {
        int valid, kind;
        int lw, lh;              // block size (log2) in its own plane
        int px, py;              // position in its plane
        int mode;
        int tr_h, tr_v;
        int cw, ch;              // coded extent: min(32, w), min(32, h)
        int kv, kh;              // rows / columns that can be non-zero
        int kv4, kh4;            // rounded up to 4
        int thin;                // block with a dimension < 4 (chroma, ISP): raster coefficient order with row stride sw
        int bdpcm, bdpcm_dir;    // transform-skip block of a BDPCM coding unit: running sum along rows (0) / columns (1)
        int isp, one_d, sw;      // ISP partition; 1xN / Nx1: ONE 1-D pass with shift TR_SHIFT_H + 1 (rcn_1xX_tb / rcn_Xx1_tb)
        int lfnst, lf_set, lf_tr, lf_is8, lf_idx;
        int scale, shift, neg;   // dequantisation
        uint64_t sig, sig_raw;   // sig_sb_map with bit 0 forced for the all-zero case / as recorded (reorder_tb_4x4 does not force it)
        const int16_t *src;
        int cbf_c, chroma_scale;
    };

    OVD int dequant(int c, const Tb &t)
    {
        int v = t.neg ? c * (t.scale << t.shift) : (c * t.scale + ((1 << t.shift) >> 1)) >> t.shift;
        return ovd_clamp(v, -32767, 32767);                       // ov_clip_intp2(v, 16), ovutils.h:79-93
    }

    OVD Tb load_tb(const Args &a, int idx)
    {
        Tb t;
        t.valid = idx < a.nb_items;
        if (!t.valid) return t;
        const TbItemDev it = a.items[idx];
        const TuCmdDev &c = a.cmds[it.cmd];
        t.kind = (int)it.kind;
        const int luma = t.kind == 0;
        const int slot = luma ? 2 : (t.kind == 2 ? 1 : 0);
        const int bit = luma ? 0x10 : (t.kind == 1 ? 0x2 : 0x1);              // transform-skip bit (joint: 0x1, :806)
        t.lw = c.log2_w - !luma; t.lh = c.log2_h - !luma;
        t.px = luma ? c.x0 : c.x0 >> 1; t.py = luma ? c.y0 : c.y0 >> 1;
        t.cw = 1 << ovd_min(5, t.lw); t.ch = 1 << ovd_min(5, t.lh);
        t.thin = t.lw < 2 || t.lh < 2;
        t.isp = luma && (c.flags & TU_ISP) != 0;
        t.one_d = t.isp && (t.lw == 0 || t.lh == 0);
        t.sw = 1 << t.lw;
        t.sig = t.sig_raw = c.sig_sb_map[slot];
        t.sig |= !t.sig;
        t.src = a.arena + c.coeff_off[slot];
        t.cbf_c = c.cbf_mask & 3;
        t.chroma_scale = ((c.flags & TU_SCALE_ANCHOR) && a.scales) ? a.scales[c.scale_anchor] : c.chroma_scale;
        const int ts = (c.tr_skip_mask & bit) != 0;
        int qp;
        if (luma) qp = ts ? c.qp_ts[0] : c.qp[0];
        else if (t.kind == 3) {
            if (!ts) qp = t.cbf_c == 3 ? c.qp[3] : (t.cbf_c == 1 ? c.qp[2] : c.qp[1]);
            else     qp = t.cbf_c == 3 ? c.qp_ts[3] : (t.cbf_c == 2 ? c.qp_ts[1] : c.qp_ts[2]);
        } else qp = ts ? c.qp_ts[t.kind] : c.qp[t.kind];
        // derive_dequant_{sdh,dpq,ts}, rcn_dequant.c:92-157
        const int s = t.lw + t.lh;
        int shift;
        if (ts)                 { shift = 6 - qp / 6; t.scale = tr_iq_scale[0][qp % 6]; }
        else if (a.dep_quant)   { shift = 6 + 1 - (15 - a.bitdepth) - (qp + 1) / 6 + (s >> 1) + (s & 1); t.scale = tr_iq_scale[s & 1][(qp + 1) % 6]; }
        else                    { shift = 6 - (15 - a.bitdepth) - qp / 6 + (s >> 1) + (s & 1); t.scale = tr_iq_scale[s & 1][qp % 6]; }
        t.neg = shift < 0; t.shift = shift < 0 ? -shift : shift;

        t.lfnst = 0; t.lf_set = t.lf_tr = t.lf_is8 = t.lf_idx = 0;
        t.tr_h = t.tr_v = 2;
        t.bdpcm = ts && ((c.bdpcm >> (luma ? 0 : 2)) & 1);
        t.bdpcm_dir = (c.bdpcm >> (luma ? 1 : 3)) & 1;
        if (ts) {
            t.mode = MODE_TS;
        } else if (t.isp) {
            // rcn_isp_tu / rcn_2xX_tb / rcn_Xx2_tb (rcn_transform_tree.c:870-945, 1000-1026): always the full transform; the
            // recorder passes the DST-VII / DCT-II choice of recon_isp_subtree_{v,h} (:1110-1111, 1178-1179); LFNST forces DCT-II
            t.mode = MODE_TR;
            t.tr_h = (c.cu_mts_idx & 1) ? 0 : 2; t.tr_v = (c.cu_mts_idx & 2) ? 0 : 2;
            t.lfnst = !t.thin && (c.flags & TU_LFNST_L) != 0;
            if (t.lfnst) t.tr_h = t.tr_v = 2;
        } else {
            const int lf = luma ? (c.flags & TU_LFNST_L) : (c.flags & TU_LFNST_C);
            if (luma && (c.flags & TU_IMPLICIT_MTS) && !c.cu_mts_flag && (t.lw <= 4 || t.lh <= 4) && !lf) {
                t.mode = MODE_TR; t.tr_h = t.lw <= 4 ? 0 : 2; t.tr_v = t.lh <= 4 ? 0 : 2;         // :435-449
            } else if (luma && c.cu_mts_flag) {
                t.mode = MODE_TR; t.tr_h = c.cu_mts_idx & 1; t.tr_v = c.cu_mts_idx >> 1;            // :492-504
            } else {
                t.lfnst = lf != 0;
                t.mode = (!c.last_pos[slot] && !t.lfnst) ? MODE_DC : MODE_TR;
            }
        }
        if (t.lfnst) {
            // drv_lfnst_mode_l / derive_lfnst_mode_c + process_lfnst(_luma), drv_lfnst.c:50-156
            int mode = luma ? c.lfnst_mode_l : c.lfnst_mode_c;
            if (mode > 1 && !t.isp) {                 // ISP: already mapped with the coding block's shape (drv_lfnst_mode_l, :1118)
                const int d = t.lw - t.lh, ms = tr_wide_shift[d < 0 ? -d : d];
                if (t.lw > t.lh && mode < 2 + ms) mode += 65;
                else if (t.lh > t.lw && mode > 66 - ms) mode -= 67;
            }
            if (!t.isp) mode = mode < 0 ? mode + 14 + 67 : (mode >= 67 ? mode + 14 : mode);
            t.lf_set = mode <= 1 ? 0 : ((mode <= 12 || mode >= 56) ? 1 : ((mode <= 23 || (mode >= 45 && mode <= 55)) ? 2 : 3));
            t.lf_tr = (mode < 67 && mode > 34) || mode >= 67 + 14;
            t.lf_is8 = t.lw >= 3 && t.lh >= 3;
            t.lf_idx = c.lfnst_idx & 1;
        }
        // extent that can be non-zero: nb_row / nb_col of the reference (derive_nb_rows / derive_nb_cols, LFNST: 4 or 8)
        if (t.thin) { t.kv = t.ch; t.kh = t.cw; }
        else if (t.lfnst) { t.kv = t.kh = 4 << t.lf_is8; }
        else {
            uint32_t colmask = 0; int maxrow = 0;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                uint32_t b = (uint32_t)(t.sig >> (8 * r)) & 0xff;
                if (b) maxrow = r;
                colmask |= b;
            }
            int maxcol = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) if ((colmask >> q) & 1) maxcol = q;
            t.kv = ovd_min(t.ch, 4 * (maxrow + 1)); t.kh = ovd_min(t.cw, 4 * (maxcol + 1));
        }
        t.kv4 = (t.kv + 3) & ~3; t.kh4 = (t.kh + 3) & ~3;
        return t;
    }

    // dequantised coefficient at column x, row y (0 outside the coded / significant area)
    OVD int coeff_at(const Tb &t, int x, int y)
    {
        if (x >= t.cw || y >= t.ch) return 0;
        if (t.thin) return dequant(t.src[y * t.sw + x], t);                                   // dequant_non_4x4_sb (raster, stride w)
        if (!((t.sig >> (((y >> 2) << 3) + (x >> 2))) & 1)) return 0;
        return dequant(t.src[(((y >> 2) * (t.cw >> 2) + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3)], t);
    }

    // LFNST output i of the kernel (rcn_lfnst.c:41-162); inputs = first 16 dequantised coefficients in diagonal scan
    OVD int lfnst_out(const Tb &t, int i)
    {
        const int nout = t.lf_is8 ? 48 : 16;
        const int nin = (!t.lf_is8 && t.lw == t.lh) ? 8 : 16;
        const int8_t *M = t.lf_is8 ? vvc_lfnst_8x8[t.lf_set][t.lf_idx] : vvc_lfnst_4x4[t.lf_set][t.lf_idx];
        int sum = 0;
        for (int j = 0; j < nin; j++) {
            const int p = tr_lfnst_scan[j];
            sum += coeff_at(t, p & 3, p >> 2) * M[i + j * nout];
        }
        return (int16_t)ovd_clamp((sum + 64) >> 7, -32768, 32768);       // sic: +32768 wraps in the int16 store (rcn_lfnst.c:64)
    }

    // one 1-D pass: out[j][line] = clip16((sum_k in[line][k] * T[j][k] + rnd) >> shift).  A thread owns output j of L
    // consecutive lines: one matrix word (fetched once, pointer stepped instead of re-derived) feeds 2 * L IDP.2A.
    OVD void tr_pass(const int16_t *in, int ls, int nlines, int log2n, int k4n, const uint32_t *Tw, int shift,
                     int16_t *out, int out_sj, int out_sl, int gt)
    {
        const int L = GROUP == 32 ? 2 : 4;
        const int n = 1 << log2n, total = ((nlines + L - 1) / L) << log2n, rnd = 1 << (shift - 1);
        for (int o = gt; o < total; o += GROUP) {
            const int j = o & (n - 1), line0 = (o >> log2n) * L;
            const uint32_t *tp = Tw + j;
            const uint32_t *iw[L];
            int acc[L];
#pragma unroll
            for (int l = 0; l < L; l++) { iw[l] = (const uint32_t *)(in + ovd_min(line0 + l, nlines - 1) * ls); acc[l] = 0; }
            for (int k4 = 0; k4 < k4n; k4++, tp += n) {
                const uint32_t tw = *tp;
#pragma unroll
                for (int l = 0; l < L; l++) {
                    acc[l] = ovd_dp2a_lo(iw[l][2 * k4], tw, acc[l]);
                    acc[l] = ovd_dp2a_hi(iw[l][2 * k4 + 1], tw, acc[l]);
                }
            }
#pragma unroll
            for (int l = 0; l < L; l++)
                if (line0 + l < nlines) out[j * out_sj + (line0 + l) * out_sl] = (int16_t)ovd_clamp((acc[l] + rnd) >> shift, -32768, 32767);
        }
    }

    // residual add family (rcn_residuals.c:46-226)
    OVD void add_plane(const Args &a, int comp, const Tb &t, const int16_t *res, int dcv, int sub, int half, int do_scale,
                       int scale, int gt)
    {
        const int w = 1 << t.lw, n = 1 << (t.lw + t.lh), maxv = (1 << a.bitdepth) - 1;
        // selects, not a.dst[comp]: dynamic indexing would copy the kernel parameter block to local memory
        const int pitch = comp == 0 ? a.dst_pitch[0] : (comp == 1 ? a.dst_pitch[1] : a.dst_pitch[2]);
        uint16_t *base = (comp == 0 ? a.dst[0] : (comp == 1 ? a.dst[1] : a.dst[2])) + (size_t)t.py * pitch + t.px;
        for (int o = gt; o < n; o += GROUP) {
            int v = t.mode == MODE_DC ? dcv : res[o];
            if (sub) v = -v;
            if (half) v >>= 1;
            if (do_scale) {
                const int sign = v & (1 << 15);                                   // sic, rcn_residuals.c:60
                int m = v < 0 ? -v : v;
                m = (ovd_min(m, maxv) * scale + (1 << 10)) >> 11;
                v = ovd_clamp(sign ? -m : m, -(1 << 15), 1 << 15);
            }
            uint16_t *p = base + (size_t)(o >> t.lw) * pitch + (o & (w - 1));
            *p = (uint16_t)ovd_clamp((int)*p + v, 0, maxv);
        }
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int grp = tid / GROUP, gt = tid % GROUP;
        // derived once (phase 0) and shared through shared memory: in the dataflow kernel phase 3 runs on a block's critical path
        Tb *const stb = (Tb *)(smem + grp * SM_GROUP + (SM_CT + SM_T2 + SM_RES) * 2);
        Tb t;
        if (P == 0) {
            t = load_tb(a, bx * GROUPS + grp);
            if (gt == 0) { if (t.valid) *stb = t; else stb->valid = 0; }
        } else t = *stb;
        if (!t.valid) return;
        int16_t *ct = (int16_t *)(smem + grp * SM_GROUP);
        int16_t *t2 = ct + SM_CT;
        int16_t *res = t2 + SM_T2;
        const int ls1 = t.kv4 + 2, ls2 = t.kh4 + 2;

        if (P == 0) {
            if (t.mode == MODE_TS && t.bdpcm) {
                // rcn_bdpcm_tb (:666-688): levels in raster order (reorder_tb_4x4 when they arrive by sub-block), clipped running
                // sum along the rows or columns -- one thread per line --, then every WHOLE group of 16 values is dequantised
                const int w = 1 << t.lw, h = 1 << t.lh, n = w * h, nl = t.bdpcm_dir ? w : h, len = t.bdpcm_dir ? h : w;
                const int sb_order = a.sh_ts_disabled && !t.thin;
                for (int l = gt; l < nl; l += GROUP) {
                    int acc = 0;
                    for (int k = 0; k < len; k++) {
                        const int x = t.bdpcm_dir ? l : k, y = t.bdpcm_dir ? k : l, o = y * w + x;
                        int v;
                        if (!sb_order) v = t.src[o];
                        else v = ((t.sig_raw >> (((y >> 2) << 3) + (x >> 2))) & 1) ? t.src[(((y >> 2) * (w >> 2) + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3)] : 0;
                        acc = k ? ovd_clamp(acc + v, -32768, 32767) : v;
                        res[o] = (int16_t)(o < ((n >> 4) << 4) ? dequant(acc, t) : acc);
                    }
                }
                return;
            }
            if (t.mode == MODE_TS) {
                // rcn_transform_skip_tb_{l,c} without BDPCM (:691-716, :1207-1225)
                const int n = 1 << (t.lw + t.lh), w = 1 << t.lw;
                for (int o = gt; o < n; o += GROUP) {
                    int v;
                    if (!a.sh_ts_disabled) v = t.src[o];
                    else if (!t.thin) v = coeff_at(t, o & (w - 1), o >> t.lw);
                    else v = o < ((n >> 4) << 4) ? dequant(t.src[o], t) : t.src[o];       // sic: whole groups of 16 only
                    res[o] = (int16_t)v;
                }
                return;
            }
            if (t.mode == MODE_DC) return;
            // dequantise + de-swizzle (+ LFNST) into ct[x][y]
            const int total = t.kh4 * t.kv4;
            for (int o = gt; o < total; o += GROUP) {
                const int x = o % t.kh4, y = o / t.kh4;
                int v;
                int li = -1;
                if (t.lfnst) {
                    const int lx = t.lf_tr ? y : x, ly = t.lf_tr ? x : y;
                    if (t.lf_is8) {
                        if (ly < 4 && lx < 8) li = ly * 8 + lx;
                        else if (ly < 8 && lx < 4) li = 32 + (ly - 4) * 4 + lx;
                    } else if (lx < 4 && ly < 4) li = ly * 4 + lx;
                }
                if (li >= 0) v = lfnst_out(t, li);
                else v = (x < t.kh && y < t.kv) ? coeff_at(t, x, y) : 0;
                if (t.one_d && t.lh == 0) t2[y * ls2 + x] = (int16_t)v;       // Nx1: the row is the input of the only (horizontal) pass
                else ct[x * ls1 + y] = (int16_t)v;
            }
            return;
        }
        if (t.mode != MODE_TR) {
            if (P != 3) return;
        }
        if (t.one_d) {
            // rcn_1xX_tb / rcn_Xx1_tb (:947-966, 1028-1048): one pass straight into the residual, shift TR_SHIFT_H + 1
            if (P == 1 && t.lw == 0) tr_pass(ct, ls1, 1, t.lh, t.kv4 >> 2, vvc_tr_words + vvc_tr_woff[t.tr_v][t.lh], 21 - a.bitdepth, res, 1, 0, gt);
            if (P == 2 && t.lh == 0) tr_pass(t2, ls2, 1, t.lw, t.kh4 >> 2, vvc_tr_words + vvc_tr_woff[t.tr_h][t.lw], 21 - a.bitdepth, res, 1, 0, gt);
            if (P != 3) return;
        }
        if (P == 1) {
            // vertical: TRFunc->func[tr_v][log2_h](coeffs, tmp, cb_w, nb_row, nb_col, 7)
            tr_pass(ct, ls1, t.kh4, t.lh, t.kv4 >> 2, vvc_tr_words + vvc_tr_woff[t.tr_v][t.lh], 7, t2, ls2, 1, gt);
            return;
        }
        if (P == 2) {
            // horizontal: func[tr_h][log2_w](tmp, dst, tb_h, tb_h, nb_row, 20 - bitdepth)
            tr_pass(t2, ls2, 1 << t.lh, t.lw, t.kh4 >> 2, vvc_tr_words + vvc_tr_woff[t.tr_h][t.lw], 20 - a.bitdepth, res, 1, 1 << t.lw, gt);
            return;
        }
        if (P == 3) {
            int dcv = 0;
            if (t.mode == MODE_DC) {                                            // inverse_dct_ii_dc, rcn_transform.c:579-597
                const int dc = coeff_at(t, 0, 0);
                dcv = ovd_clamp((((dc + 1) >> 1) + (1 << (13 - a.bitdepth))) >> (14 - a.bitdepth), -32768, 32767);
            }
            const int scale_on = a.ict_type & 1, sub_on = (a.ict_type >> 1) & 1;   // rcn_init_ict_functions, rcn_residuals.c:229-331
            if (t.kind == 0) {
                add_plane(a, 0, t, res, dcv, 0, 0, 0, 0, gt);
            } else if (t.kind == 3) {                                           // rcn_jcbcr, :796-867
                const int scale = (t.lw + t.lh == 2) ? 1 << 11 : t.chroma_scale;
                if (t.cbf_c == 3) {
                    add_plane(a, 1, t, res, dcv, 0, 0, scale_on, scale, gt);
                    add_plane(a, 2, t, res, dcv, sub_on, 0, scale_on, scale, gt);
                } else if (t.cbf_c == 2) {
                    add_plane(a, 1, t, res, dcv, 0, 0, scale_on, scale, gt);
                    add_plane(a, 2, t, res, dcv, sub_on, 1, scale_on, scale, gt);
                } else {
                    add_plane(a, 2, t, res, dcv, 0, 0, scale_on, scale, gt);
                    add_plane(a, 1, t, res, dcv, sub_on, 1, scale_on, scale, gt);
                }
            } else {                                                            // rcn_res_c: ict.ict[..][0] vs ict.add[..], :749-753
                const int big = t.lw + t.lh > 2;
                add_plane(a, t.kind, t, res, dcv, 0, 0, big && scale_on, t.chroma_scale, gt);
            }
        }
    }
};

// =====================================================================================================================
// LMCS chroma residual scale of a 64x64 luma region (rcn_lmcs_compute_chroma_scale, libovvc/rcn_lmcs.c:204-343): average
// of the reconstructed (reshaped-domain) luma samples above and left of the region -- the available run of 4-sample units,
// padded to 16 units with the run's last sample -- mapped through the LMCS window bounds to 2^17 / (window size + offset).
// The reference evaluates it when a coding unit starts on a 64-aligned position (coding_unit, vcl_coding_unit.c:725-730)
// from its CTU buffer; here one warp per recorded anchor reads the finished luma plane (the schedule runs all luma first).
struct LmcsAnchorDev { int16_t x0, y0; uint16_t abv_mask, lft_mask; };   // == OVGPULmcsAnchor

struct LmcsScaleK {
    static const int THREADS = 128, MIN_BLOCKS = 4, PHASES = 2;
    static const int WARPS = 4;
    static const int SMEM_BYTES = WARPS * 32 * 4;
    struct Args {
        const uint16_t *luma;
        int pitch, bitdepth;
        const LmcsAnchorDev *anchors;
        int nb;
        int16_t *scales;
        const uint16_t *wnd_bnd;         // 17 entries (device)
        int min_idx, max_idx_p1, crs_offset;
    };
    struct State { int dummy; };

    OVD int run_len(unsigned m) { int n = 0; while (m) { n++; m >>= 1; } return n; }     // the reference walks until the mask is empty

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        const int warp = tid >> 5, lane = tid & 31, idx = bx * WARPS + warp;
        if (idx >= a.nb) return;
        int *part = (int *)smem + warp * 32;
        const LmcsAnchorDev an = a.anchors[idx];
        const int n_abv = run_len(an.abv_mask), n_lft = run_len(an.lft_mask);
        const uint16_t *src = a.luma + (ptrdiff_t)an.y0 * a.pitch + an.x0;
        if (P == 0) {
            const int left = lane >> 4, u = lane & 15, n = left ? n_lft : n_abv;
            const ptrdiff_t step = left ? a.pitch : 1;
            const uint16_t *line = left ? src - 1 : src - a.pitch;
            int sum = 0;
            if (u < n) sum = line[(4 * u) * step] + line[(4 * u + 1) * step] + line[(4 * u + 2) * step] + line[(4 * u + 3) * step];
            else if (n) sum = 4 * line[(4 * n - 1) * step];
            part[lane] = sum;
            return;
        }
        if (lane) return;
        int sum = 0;
        for (int i = 0; i < 32; i++) sum += part[i];
        const int nb_units = (n_abv ? 16 : 0) + (n_lft ? 16 : 0);
        int avg = 1 << (a.bitdepth - 1);
        if (nb_units) { const int l2 = nb_units == 32 ? 6 : 5; avg = (sum + (1 << l2)) >> (l2 + 1); }
        int i = a.min_idx;
        for (; i < a.max_idx_p1; i++) if (avg < a.wnd_bnd[i + 1]) break;               // get_bwd_idx, rcn_lmcs.c:84-94
        i = ovd_min(i, 15);
        const int wnd_sz = (int)a.wnd_bnd[i + 1] - (int)a.wnd_bnd[i], den = wnd_sz + a.crs_offset;
        const int scale = (wnd_sz == 0 || den == 0) ? 1 << 11 : (1 << (a.bitdepth - 4 + 11)) / den;
        a.scales[idx] = (int16_t)(uint16_t)scale;                                      // uint16_t lmcs_chroma_scale, read as int16_t
    }
};
