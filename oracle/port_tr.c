/* oracle/port_tr.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of OpenVVC's residual path for transform units of non-intra CUs: tmp.rcn_tu_st
 * (libovvc/rcn_transform_tree.c:1228-1301) -> rcn_residual (:415-506), rcn_res_c (:720-793), rcn_jcbcr
 * (:796-867), rcn_residual_c (:553-628), transform skip (:691-716, 1208-1225), dequantisation
 * (libovvc/rcn_dequant.c:92-312), the separable inverse transforms (libovvc/rcn_transform.c:45-598) and
 * the residual add / inter-component variants (libovvc/rcn_residuals.c:46-331).
 * The transform matrices are the normative ones, read back from the compiled reference
 * (tools/gen_vvc_tables.py -> openvvc_b200/csrc/vvc_tr_tables.h, data only).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define OVD_CONST static const
#include "../openvvc_b200/csrc/vvc_tr_tables.h"

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ov_clip_intp2(val, 16): symmetric +-32767 (ovutils.h:79-93) */
static inline int clip_sym16(int v) { return v > 32767 ? 32767 : (v < -32767 ? -32767 : v); }

typedef struct { int scale, shift, neg; } Deq;
static const int iq_scale[2][6] = { { 40, 45, 51, 57, 64, 72 }, { 57, 64, 72, 80, 90, 102 } };

/* derive_dequant_sdh / _dpq / _ts, rcn_dequant.c:92-157 */
static Deq derive_dequant(int kind /*0 sdh, 1 dpq, 2 ts*/, int qp, int lw, int lh, int bitdepth)
{
    Deq d;
    int s = lw + lh, shift, scale;
    if (kind == 2)      { shift = 6 - qp / 6; scale = iq_scale[0][qp % 6]; }
    else if (kind == 1) { shift = 6 + 1 - (15 - bitdepth) - (qp + 1) / 6 + (s >> 1) + (s & 1); scale = iq_scale[s & 1][(qp + 1) % 6]; }
    else                { shift = 6 - (15 - bitdepth) - qp / 6 + (s >> 1) + (s & 1); scale = iq_scale[s & 1][qp % 6]; }
    d.neg = shift < 0; d.shift = shift < 0 ? -shift : shift; d.scale = scale;
    return d;
}

static inline int dequant_one(int c, Deq d)
{
    if (d.neg) return clip_sym16(c * (d.scale << d.shift));
    return clip_sym16((c * d.scale + ((1 << d.shift) >> 1)) >> d.shift);
}

/* dequant_tb_4x4(_neg), rcn_dequant.c:160-312: coefficients arrive as 4x4 sub-blocks of 16 contiguous
 * values, min(32,w)/4 sub-blocks per row; insignificant sub-blocks are zeroed. dst is raster, stride min(32,w). */
static void dequant_deswizzle(int16_t *dst, const int16_t *src, Deq d, int lw, int lh, uint64_t sig)
{
    int cw = 1 << imin(5, lw), chh = 1 << imin(5, lh);
    sig |= !sig;
    for (int i = 0; i < chh / 4; i++) {
        unsigned row = (unsigned)(sig >> (i << 3)) & 0xff;
        for (int j = 0; j < cw / 4; j++) {
            const int16_t *s = src + (size_t)i * 4 * cw + (j << 4);
            for (int k = 0; k < 16; k++)
                dst[(4 * i + (k >> 2)) * cw + 4 * j + (k & 3)] = (row >> j) & 1 ? (int16_t)dequant_one(s[k], d) : 0;
        }
    }
}

/* out[line][j] = clip16((sum_k in[k][line] * M[k][j] + rnd) >> shift): matrix_multiplication, rcn_transform.c:45-68.
 * in: kmax x nlines (stride in_stride), out: nlines x n */
static void tr_1d(int16_t *out, const int16_t *in, int in_stride, int nlines, int type, int log2n, int kmax, int shift)
{
    int n = 1 << log2n;
    const int8_t *T = vvc_tr_blob + vvc_tr_off[type][log2n];   /* T[j*n + k] */
    for (int i = 0; i < nlines; i++)
        for (int j = 0; j < n; j++) {
            int sum = 0;
            for (int k = 0; k < kmax; k++) sum += in[k * in_stride + i] * T[j * n + k];
            out[i * n + j] = (int16_t)clipi((sum + (1 << (shift - 1))) >> shift, -32768, 32767);
        }
}

/* 2-D inverse transform of a dequantised block C (stride cw = min(32,w)) into res (stride w). */
static void inv_transform_2d(int16_t *res, const int16_t *C, int lw, int lh, int tr_h, int tr_v, int bitdepth)
{
    int w = 1 << lw, h = 1 << lh, cw = imin(32, w), chh = imin(32, h);
    int16_t *tmp = malloc(sizeof(int16_t) * 64 * 64);
    /* vertical first: TRFunc->func[tr_v][log2_h](C, tmp, cw, ...), shift 7; tmp[x][y'] */
    tr_1d(tmp, C, cw, cw, tr_v, lh, chh, 7);
    /* then horizontal: func[tr_h][log2_w](tmp, dst, tb_h, tb_h, ...), shift 20 - bitdepth */
    tr_1d(res, tmp, h, h, tr_h, lw, cw, 20 - bitdepth);
    free(tmp);
}

static void dc_fill(int16_t *res, int lw, int lh, int dc, int bitdepth)
{   /* inverse_dct_ii_dc, rcn_transform.c:579-597 */
    int v = clipi((((dc + 1) >> 1) + (1 << (13 - bitdepth))) >> (14 - bitdepth), -32768, 32767);
    for (int i = 0; i < (1 << (lw + lh)); i++) res[i] = (int16_t)v;
}

typedef struct { int dep_quant, sh_ts_disabled, ict_type, bitdepth; } ResCfg;

/* transform-skip block (no BDPCM: non-intra CU), rcn_transform_tree.c:691-716 / 1208-1225 */
/* rcn_bdpcm_tb (:666-688): with sh_ts_disabled the levels of blocks >= 4x4 arrive in sub-block order and are put in raster
 * order first (reorder_tb_4x4 :134-206, insignificant sub-blocks zero); clipped running sum along rows (dir 0, apply_bdpcm_1)
 * or columns (dir 1, apply_bdpcm_2); then every whole group of 16 values is dequantised with the transform-skip scale */
static void bdpcm_block(int16_t *res, const int16_t *src, int lw, int lh, int qp, uint64_t sig, const ResCfg *cfg, int dir)
{
    int w = 1 << lw, h = 1 << lh, n = w * h;
    int16_t tmp[32 * 32];
    const int16_t *in = src;
    Deq d = derive_dequant(2, qp, lw, lh, cfg->bitdepth);
    if (cfg->sh_ts_disabled && lw > 1 && lh > 1) {
        for (int i = 0; i < h / 4; i++)
            for (int j = 0; j < w / 4; j++)
                for (int k = 0; k < 16; k++)
                    tmp[(4 * i + (k >> 2)) * w + 4 * j + (k & 3)] = ((sig >> (8 * i + j)) & 1) ? src[(size_t)i * 4 * w + (j << 4) + k] : 0;
        in = tmp;
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int first = dir ? y == 0 : x == 0, prev = first ? 0 : res[dir ? (y - 1) * w + x : y * w + x - 1];
            res[y * w + x] = (int16_t)(first ? in[y * w + x] : clipi(prev + in[y * w + x], -32768, 32767));
        }
    for (int i = 0; i < (n >> 4) * 16; i++) res[i] = (int16_t)dequant_one(res[i], d);
}

static void ts_block(int16_t *res, const int16_t *src, int lw, int lh, int qp, uint64_t sig, const ResCfg *cfg, int is_chroma)
{
    int n = 1 << (lw + lh);
    if (!cfg->sh_ts_disabled) { memcpy(res, src, n * sizeof(int16_t)); return; }
    Deq d = derive_dequant(2, qp, lw, lh, cfg->bitdepth);
    if (!is_chroma || (lw > 1 && lh > 1)) {
        dequant_deswizzle(res, src, d, lw, lh, sig);
    } else {
        memcpy(res, src, n * sizeof(int16_t));
        for (int i = 0; i < (n >> 4) * 16; i++) res[i] = (int16_t)dequant_one(res[i], d);   /* sic: whole groups of 16 only */
    }
}

/* drv_lfnst_mode_l / derive_lfnst_mode_c tail (drv_lfnst.c:50-118): wide-angle remap of the intra mode */
static int lfnst_wide_mode(int lw, int lh, int mode)
{
    static const uint8_t mode_shift_lut[6] = { 0, 6, 10, 12, 14, 15 };
    if (mode > 1) {
        int ms = mode_shift_lut[abs(lw - lh)];
        if (lw > lh && mode < 2 + ms) mode += 65;
        else if (lh > lw && mode > 66 - ms) mode -= 67;
    }
    return mode < 0 ? mode + 14 + 67 : (mode >= 67 ? mode + 14 : mode);
}

/* process_lfnst(_luma) (drv_lfnst.c:120-156) + compute_lfnst_{4x4,8x8}(_tr) (rcn_lfnst.c:41-162), in place on the
 * dequantised block C (stride 1 << lcw). */
static void lfnst_apply2(int16_t *C, int lcw, int lw, int lh, int lfnst_idx, int mode);
static void lfnst_apply(int16_t *C, int lcw, int lw, int lh, int lfnst_idx, int intra_mode)
{
    lfnst_apply2(C, lcw, lw, lh, lfnst_idx, lfnst_wide_mode(lw, lh, intra_mode));
}
/* mode: already mapped by drv_lfnst_mode_l (ISP maps with the coding block's size, rcn_transform_tree.c:1118, 1172) */
static void lfnst_apply2(int16_t *C, int lcw, int lw, int lh, int lfnst_idx, int mode)
{
    static const uint8_t scan[16] = { 0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15 };   /* nibbles of 0xfbe7ad369c258140 */
    int set = mode <= 1 ? 0 : ((mode <= 12 || mode >= 56) ? 1 : ((mode <= 23 || (mode >= 45 && mode <= 55)) ? 2 : 3));
    int tr = (mode < 67 && mode > 34) || mode >= 67 + 14;
    int is8 = lw >= 3 && lh >= 3;
    int16_t sb[16], in[16];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) sb[4 * r + k] = C[(r << lcw) + k];
    for (int i = 0; i < 16; i++) in[i] = sb[scan[i]];
    if (!is8) {
        const int8_t *M = vvc_lfnst_4x4[set][lfnst_idx];
        int nin = lw == lh ? 8 : 16;
        for (int i = 0; i < 16; i++) {
            int sum = 0;
            for (int j = 0; j < nin; j++) sum += in[j] * M[i + j * 16];
            int v = clipi((sum + 64) >> 7, -32768, 32768);
            C[tr ? ((i & 3) << lcw) + (i >> 2) : (i & 3) + ((i >> 2) << lcw)] = (int16_t)v;
        }
    } else {
        const int8_t *M = vvc_lfnst_8x8[set][lfnst_idx];
        for (int i = 0; i < 48; i++) {
            int sum = 0;
            for (int j = 0; j < 16; j++) sum += in[j] * M[i + j * 48];
            int v = clipi((sum + 64) >> 7, -32768, 32768);
            int x = i < 32 ? (i & 7) : (i & 3), y = i < 32 ? (i >> 3) : 4 + ((i - 32) >> 2);
            C[tr ? (x << lcw) + y : x + (y << lcw)] = (int16_t)v;
        }
    }
}

/* rcn_residual (:415-506); for non-intra CUs `is_mip |= !is_intra` disables implicit MTS */
static void luma_tb(int16_t *res, const int16_t *src, const OVOTuCmd *c, const ResCfg *cfg)
{
    int lw = c->log2_w, lh = c->log2_h;
    int lfnst = (c->flags & OVO_TU_LFNST_L) != 0;
    if ((c->tr_skip_mask & 0x10) && (c->bdpcm & 1)) { bdpcm_block(res, src, lw, lh, c->qp_ts[0], c->sig_sb_map[2], cfg, (c->bdpcm >> 1) & 1); return; }
    if (c->tr_skip_mask & 0x10) { ts_block(res, src, lw, lh, c->qp_ts[0], c->sig_sb_map[2], cfg, 0); return; }
    int16_t C[32 * 32];
    dequant_deswizzle(C, src, derive_dequant(cfg->dep_quant, c->qp[0], lw, lh, cfg->bitdepth), lw, lh, c->sig_sb_map[2]);
    if ((c->flags & OVO_TU_IMPLICIT_MTS) && !c->cu_mts_flag && (lw <= 4 || lh <= 4) && !lfnst) {
        inv_transform_2d(res, C, lw, lh, lw <= 4 ? 0 : 2, lh <= 4 ? 0 : 2, cfg->bitdepth);
    } else if (!c->cu_mts_flag) {
        if (lfnst) lfnst_apply(C, imin(5, lw), lw, lh, c->lfnst_idx, c->lfnst_mode_l);
        if (!c->last_pos[2] && !lfnst) dc_fill(res, lw, lh, C[0], cfg->bitdepth);
        else inv_transform_2d(res, C, lw, lh, 2, 2, cfg->bitdepth);
    } else {
        inv_transform_2d(res, C, lw, lh, c->cu_mts_idx & 1, c->cu_mts_idx >> 1, cfg->bitdepth);
    }
}

/* ISP transform blocks: rcn_isp_tu (:870-917, >= 4 samples both ways: dequantisation by sub-blocks, optional LFNST which
 * forces DCT-II, full 2-D transform, no DC shortcut), rcn_2xX_tb / rcn_Xx2_tb (:919-945, 1000-1026: plain dequantisation, 2-D
 * transform with a 2-point pass) and rcn_1xX_tb / rcn_Xx1_tb (:947-966, 1028-1048: ONE 1-D pass with shift TR_SHIFT_H + 1). */
void ovo_isp_residual(int16_t *res, const int16_t *arena, const OVOTuCmd *c, int dep_quant, int bitdepth)
{
    int lw = c->log2_w, lh = c->log2_h, w = 1 << lw, h = 1 << lh;
    int type_h = (c->cu_mts_idx & 1) ? 0 : 2, type_v = (c->cu_mts_idx & 2) ? 0 : 2;
    const int16_t *src = arena + c->coeff_off[2];
    Deq d = derive_dequant(dep_quant, c->qp[0], lw, lh, bitdepth);
    int16_t C[32 * 32];
    if (lw >= 2 && lh >= 2) {
        dequant_deswizzle(C, src, d, lw, lh, c->sig_sb_map[2]);
        if (c->flags & OVO_TU_LFNST_L) { lfnst_apply2(C, imin(5, lw), lw, lh, c->lfnst_idx, c->lfnst_mode_l); type_h = type_v = 2; }
        inv_transform_2d(res, C, lw, lh, type_h, type_v, bitdepth);
    } else if (lw == 0 || lh == 0) {
        for (int i = 0; i < w * h; i++) C[i] = (int16_t)dequant_one(src[i], d);
        tr_1d(res, C, 1, 1, lw == 0 ? type_v : type_h, lw == 0 ? lh : lw, imin(32, w * h), 20 - bitdepth + 1);
    } else {
        /* raster source with stride w; the 2-D transform works on the min(32, w) x min(32, h) coded area (a 64-sample side
         * carries coefficients in its first half only) */
        int cw = imin(32, w), chh = imin(32, h);
        for (int y = 0; y < chh; y++) for (int x = 0; x < cw; x++) C[y * cw + x] = (int16_t)dequant_one(src[y * w + x], d);
        inv_transform_2d(res, C, lw, lh, type_h, type_v, bitdepth);
    }
}

/* rcn_residual_c, :553-628 */
static void chroma_tb(int16_t *res, const int16_t *src, int lw, int lh, int qp, int last_pos, uint64_t sig, const ResCfg *cfg,
                      int lfnst, int lfnst_idx, int lfnst_mode)
{
    int16_t C[32 * 32];
    Deq d = derive_dequant(cfg->dep_quant, qp, lw, lh, cfg->bitdepth);
    if (lw > 1 && lh > 1) dequant_deswizzle(C, src, d, lw, lh, sig);
    else for (int i = 0; i < (1 << (lw + lh)); i++) C[i] = (int16_t)dequant_one(src[i], d);   /* dequant_non_4x4_sb */
    if (!last_pos && !lfnst) { dc_fill(res, lw, lh, C[0], cfg->bitdepth); return; }
    if (lfnst) lfnst_apply(C, lw, lw, lh, lfnst_idx, lfnst_mode);
    inv_transform_2d(res, C, lw, lh, 2, 2, cfg->bitdepth);
}

/* residual add family, rcn_residuals.c:46-226: mode 0 add, 1 sub, 2 half (sign per `sub`) */
static void add_res(uint16_t *dst, int stride, const int16_t *res, int lw, int lh, int sub, int half, int do_scale, int scale, int bitdepth)
{
    int w = 1 << lw, h = 1 << lh, maxv = (1 << bitdepth) - 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = res[y * w + x];
            if (sub) v = -v;
            if (half) v >>= 1;
            if (do_scale) {
                int sign = v & (1 << 15);                              /* sic: bit 15 of the int, rcn_residuals.c:60 */
                int a = abs(v);
                a = (imin(a, maxv) * scale + (1 << 10)) >> 11;
                v = clipi(sign ? -a : a, -(1 << 15), 1 << 15);
            }
            dst[(size_t)y * stride + x] = (uint16_t)clipi((int)dst[(size_t)y * stride + x] + v, 0, maxv);
        }
}

void ovo_residual_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                          const OVOTuCmd *cmds, int nb_cmds, const int16_t *arena,
                          int dep_quant, int sh_ts_disabled, int ict_type)
{
    ResCfg cfg = { dep_quant, sh_ts_disabled, ict_type, bitdepth };
    int16_t *res = malloc(sizeof(int16_t) * 64 * 64);
    int scale_on = ict_type & 1, sub_on = (ict_type >> 1) & 1;       /* rcn_init_ict_functions, rcn_residuals.c:229-331 */
    (void)pic_w; (void)pic_h;
    for (int n = 0; n < nb_cmds; n++) {
        const OVOTuCmd *c = &cmds[n];
        int lwc = c->log2_w - 1, lhc = c->log2_h - 1;
        uint16_t *dy = dst[0] + (size_t)c->y0 * dst_stride[0] + c->x0;
        uint16_t *dcb = dst[1] + (size_t)(c->y0 >> 1) * dst_stride[1] + (c->x0 >> 1);
        uint16_t *dcr = dst[2] + (size_t)(c->y0 >> 1) * dst_stride[2] + (c->x0 >> 1);
        if ((c->cbf_mask & 0x10) && (c->flags & OVO_TU_ISP)) {
            ovo_isp_residual(res, arena, c, dep_quant, bitdepth);
            add_res(dy, dst_stride[0], res, c->log2_w, c->log2_h, 0, 0, 0, 0, bitdepth);
        } else if (c->cbf_mask & 0x10) {
            luma_tb(res, arena + c->coeff_off[2], c, &cfg);
            add_res(dy, dst_stride[0], res, c->log2_w, c->log2_h, 0, 0, 0, 0, bitdepth);
        }
        int cbf_c = c->cbf_mask & 3, lfnst_c = (c->flags & OVO_TU_LFNST_C) != 0;
        if (c->cbf_mask & 0x8) {                                     /* rcn_jcbcr, :796-867 */
            if (!(c->tr_skip_mask & 0x1)) {
                int qp = cbf_c == 3 ? c->qp[3] : (cbf_c == 1 ? c->qp[2] : c->qp[1]);
                chroma_tb(res, arena + c->coeff_off[0], lwc, lhc, qp, c->last_pos[0], c->sig_sb_map[0], &cfg, lfnst_c, c->lfnst_idx, c->lfnst_mode_c);
            } else {
                int qp = cbf_c == 3 ? c->qp_ts[3] : (cbf_c == 2 ? c->qp_ts[1] : c->qp_ts[2]);
                if (c->bdpcm & 4) bdpcm_block(res, arena + c->coeff_off[0], lwc, lhc, qp, c->sig_sb_map[0], &cfg, (c->bdpcm >> 3) & 1);
                else ts_block(res, arena + c->coeff_off[0], lwc, lhc, qp, c->sig_sb_map[0], &cfg, 1);
            }
            int scale = (lwc + lhc == 2) ? 1 << 11 : c->chroma_scale;
            if (!scale_on) scale = 1 << 11;
            if (cbf_c == 3) {
                add_res(dcb, dst_stride[1], res, lwc, lhc, 0, 0, scale_on, scale, bitdepth);
                add_res(dcr, dst_stride[2], res, lwc, lhc, sub_on, 0, scale_on, scale, bitdepth);
            } else if (cbf_c == 2) {
                add_res(dcb, dst_stride[1], res, lwc, lhc, 0, 0, scale_on, scale, bitdepth);
                add_res(dcr, dst_stride[2], res, lwc, lhc, sub_on, 1, scale_on, scale, bitdepth);
            } else {
                add_res(dcr, dst_stride[2], res, lwc, lhc, 0, 0, scale_on, scale, bitdepth);
                add_res(dcb, dst_stride[1], res, lwc, lhc, sub_on, 1, scale_on, scale, bitdepth);
            }
        } else if (cbf_c) {                                          /* rcn_res_c, :720-793 */
            for (int k = 0; k < 2; k++) {
                int bit = k == 0 ? 0x2 : 0x1;
                if (!(cbf_c & bit)) continue;
                if (!(c->tr_skip_mask & bit))
                    chroma_tb(res, arena + c->coeff_off[k], lwc, lhc, c->qp[1 + k], c->last_pos[k], c->sig_sb_map[k], &cfg, lfnst_c, c->lfnst_idx, c->lfnst_mode_c);
                else
                    if (c->bdpcm & 4) bdpcm_block(res, arena + c->coeff_off[k], lwc, lhc, c->qp_ts[1 + k], c->sig_sb_map[k], &cfg, (c->bdpcm >> 3) & 1);
                    else ts_block(res, arena + c->coeff_off[k], lwc, lhc, c->qp_ts[1 + k], c->sig_sb_map[k], &cfg, 1);
                int big = lwc + lhc > 2;                             /* ict.ict[..][0] vs ict.add[..], :749-753 */
                add_res(k == 0 ? dcb : dcr, dst_stride[1 + k], res, lwc, lhc, 0, 0, big && scale_on, c->chroma_scale, bitdepth);
            }
        }
    }
    free(res);
}
