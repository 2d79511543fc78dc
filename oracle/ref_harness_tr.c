/* oracle/ref_harness_tr.c -- TEST INFRASTRUCTURE ONLY (see ref_harness.c).
 *
 * Drives the reference's residual path: tmp.rcn_tu_st (rcn_transform_tree.c:1228) per transform unit of a
 * non-intra CU (dequantisation, inverse transform / transform skip, residual add, JCbCr, LMCS chroma scale),
 * and exposes the 1-D inverse transforms of its function table (tr.func[type][log2N], rcn_transform.c) so
 * that the normative transform matrices can be read back by impulse probing.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "ovdefs.h"
#include "ovframe.h"
#include "ctudec.h"
#include "dec_structures.h"
#include "rcn.h"
#include "rcn_structures.h"
#include "vcl.h"
#include "cu_utils.h"

#include "oracle.h"

void rcn_init_tr_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_dequant_10(struct RCNFunctions *rcn_func);
void rcn_init_ict_functions_10(struct RCNFunctions *rcn_func, uint8_t type, uint8_t bitdepth);
void rcn_init_lfnst_functions(struct RCNFunctions *rcn_func);

/* private to rcn_transform_tree.c:51-66 and vcl_transform_unit.c -- must match field for field */
struct TBInfo {
   uint16_t last_pos;
   uint64_t sig_sb_map;
};
struct TUInfo {
   uint8_t is_sbt;
   uint8_t cbf_mask;
   uint16_t pos_offset;
   uint8_t tr_skip_mask;
   uint8_t cu_mts_flag;
   uint8_t cu_mts_idx;
   uint8_t lfnst_flag;
   uint8_t lfnst_idx;
   struct TBInfo tb_info[3];
};

/* M[k][j] of the 1-D inverse transform `type` (0 DST-VII, 1 DCT-VIII, 2 DCT-II) of size 1<<log2n:
 * an impulse of 128 at input k with shift 7 returns exactly row k (rcn_transform.c:45-68). */
int
ovref_tr_matrix(int type, int log2n, int16_t *out)
{
    struct RCNFunctions f;
    int n = 1 << log2n;
    /* 8 lines of input (the SIMD kernels work on groups of lines); only line 0 carries the impulse */
    int16_t src[64 * 8], dst[64 * 8];
    memset(&f, 0, sizeof(f));
    rcn_init_functions(&f, 0, 1, 0, 0, 10);
    if (!f.tr.func[type][log2n]) return -1;
    for (int k = 0; k < n; k++) {
        memset(src, 0, sizeof(src));
        memset(dst, 0, sizeof(dst));
        src[k * 8] = 128;                       /* src[k * src_stride + line] */
        /* the 4-point SIMD kernels need a group of lines; larger ones are probed with one line so that the
         * 32-point DST-VII/DCT-VIII rows 16..31 (zeroed out by the SIMD variant for >= 8 lines) are visible too */
        f.tr.func[type][log2n](src, dst, 8, n == 4 ? 8 : 1, n, 7);
        memcpy(out + k * n, dst, n * sizeof(int16_t));   /* dst[line * n + j], line 0 */
    }
    return 0;
}

static void
no_intra_pred_c(const struct OVRCNCtx *const rcn_ctx, uint8_t intra_mode, int x0, int y0, int log2_pb_w, int log2_pb_h, CUFlags cu_flags)
{
    (void)rcn_ctx; (void)intra_mode; (void)x0; (void)y0; (void)log2_pb_w; (void)log2_pb_h; (void)cu_flags;
}

void
ovref_residual_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int log2_ctu_s,
                       int bitdepth, const OVOTuCmd *cmds, int nb_cmds, const int16_t *arena,
                       int dep_quant, int sh_ts_disabled, int ict_type, int simd)
{
    OVCTUDec *c;
    OVPartInfo part;
    enum { SCR_STRIDE = 384 };
    uint16_t (*scr)[(128 + 8) * SCR_STRIDE] = calloc(3, sizeof(*scr));
    ctudec_init(&c);
    memset(&part, 0, sizeof(part));
    part.log2_ctu_s = log2_ctu_s;
    part.log2_min_cb_s = 2;
    c->part_ctx = &part;
    c->rcn_ctx.ctudec = c;
    c->pic_w = pic_w;
    c->pic_h = pic_h;
    rcn_init_functions(&c->rcn_funcs, ict_type, 1, 0, ict_type & 1, bitdepth);
    if (!simd && bitdepth == 10) {      /* scalar C templates = arithmetic of record (rcn.c:151-178) */
        rcn_init_tr_functions_10(&c->rcn_funcs);
        rcn_init_dequant_10(&c->rcn_funcs);
        rcn_init_ict_functions_10(&c->rcn_funcs, ict_type, bitdepth);
        rcn_init_lfnst_functions(&c->rcn_funcs);
    }
    c->residual_coding_l = dep_quant ? &residual_coding_dpq : &residual_coding_sdh;
    c->sh_ts_disabled = sh_ts_disabled;
    c->transform_unit = &transform_unit_st;
    c->lmcs_info.scale_c_flag = ict_type & 1;
    c->rcn_ctx.ctu_buff.y = scr[0]; c->rcn_ctx.ctu_buff.cb = scr[1]; c->rcn_ctx.ctu_buff.cr = scr[2];
    c->rcn_ctx.ctu_buff.stride = SCR_STRIDE; c->rcn_ctx.ctu_buff.stride_c = SCR_STRIDE;
    c->part_ctx_c = &part;
    c->rcn_funcs.intra_pred_c = &no_intra_pred_c;  /* rcn_tu_st / rcn_tu_c predict chroma first: prediction is already in dst */

    for (int n = 0; n < nb_cmds; n++) {
        const OVOTuCmd *m = &cmds[n];
        int ctu_s = 1 << log2_ctu_s;
        int x0 = m->x0 & (ctu_s - 1), y0 = m->y0 & (ctu_s - 1);
        int w = 1 << m->log2_w, h = 1 << m->log2_h;
        struct TUInfo tu;
        memset(&tu, 0, sizeof(tu));
        tu.cbf_mask = m->cbf_mask;
        tu.tr_skip_mask = m->tr_skip_mask;
        tu.cu_mts_flag = m->cu_mts_flag;
        tu.cu_mts_idx = m->cu_mts_idx;
        tu.pos_offset = 0;
        for (int k = 0; k < 3; k++) { tu.tb_info[k].last_pos = m->last_pos[k]; tu.tb_info[k].sig_sb_map = m->sig_sb_map[k]; }
        /* prediction into the CTU-local buffer, coefficients into the per-CU buffers (ctudec.h:816-818) */
        for (int k = 0; k < 3; k++) {
            int sh = k ? 1 : 0;
            for (int y = 0; y < (h >> sh); y++)
                memcpy(scr[k] + (size_t)((y0 >> sh) + y) * SCR_STRIDE + (x0 >> sh),
                       dst[k] + (size_t)((m->y0 >> sh) + y) * dst_stride[k] + (m->x0 >> sh), (w >> sh) * sizeof(uint16_t));
        }
        int nl = (w < 32 ? w : 32) * (h < 32 ? h : 32);
        int nc = ((w >> 1) < 32 ? (w >> 1) : 32) * ((h >> 1) < 32 ? (h >> 1) : 32);
        if (m->cbf_mask & 0x10) memcpy(c->residual_y, arena + m->coeff_off[2], nl * sizeof(int16_t));
        if (m->cbf_mask & 0x0a) memcpy(c->residual_cb, arena + m->coeff_off[0], nc * sizeof(int16_t));
        if (m->cbf_mask & 0x01) memcpy(c->residual_cr, arena + m->coeff_off[1], nc * sizeof(int16_t));
        c->dequant_luma.qp = m->qp[0]; c->dequant_cb.qp = m->qp[1]; c->dequant_cr.qp = m->qp[2]; c->dequant_joint_cb_cr.qp = m->qp[3];
        c->dequant_luma_skip.qp = m->qp_ts[0]; c->dequant_cb_skip.qp = m->qp_ts[1]; c->dequant_cr_skip.qp = m->qp_ts[2];
        c->dequant_jcbcr_skip.qp = m->qp_ts[3];
        c->lmcs_info.lmcs_chroma_scale = m->chroma_scale;
        tu.lfnst_flag = (m->flags & (OVO_TU_LFNST_L | OVO_TU_LFNST_C)) != 0;
        tu.lfnst_idx = m->lfnst_idx;
        c->intra_mode = m->lfnst_mode_l;
        c->intra_mode_c = m->lfnst_mode_c;
        if (m->flags & OVO_TU_LFNST_C) {
            /* dual-tree chroma TU: tmp.rcn_tu_c takes chroma coordinates and sizes (rcn_transform_tree.c:1349) */
            CUFlags flc = flg_pred_mode_flag;
            if (m->bdpcm & 4) flc |= flg_intra_bdpcm_chroma_flag | ((m->bdpcm & 8) ? flg_intra_bdpcm_chroma_dir : 0);
            c->rcn_funcs.tmp.rcn_tu_c(c, x0 >> 1, y0 >> 1, m->log2_w - 1, m->log2_h - 1, flc, m->cbf_mask, (void *)&tu);
        } else {
            /* intra CU <=> implicit MTS or LFNST requested (an inter CU never uses either, :1247-1249) */
            CUFlags fl = (m->flags & (OVO_TU_IMPLICIT_MTS | OVO_TU_LFNST_L)) ? flg_pred_mode_flag : 0;
            if (m->bdpcm & 1) fl |= flg_pred_mode_flag | flg_intra_bdpcm_luma_flag | ((m->bdpcm & 2) ? flg_intra_bdpcm_luma_dir : 0);
            if (m->bdpcm & 4) fl |= flg_pred_mode_flag | flg_intra_bdpcm_chroma_flag | ((m->bdpcm & 8) ? flg_intra_bdpcm_chroma_dir : 0);
            c->mts_implicit = (m->flags & OVO_TU_IMPLICIT_MTS) != 0;
            c->rcn_funcs.tmp.rcn_tu_st(c, x0, y0, m->log2_w, m->log2_h, fl, m->cbf_mask, (void *)&tu);
        }
        for (int k = 0; k < 3; k++) {
            int sh = k ? 1 : 0;
            for (int y = 0; y < (h >> sh); y++)
                memcpy(dst[k] + (size_t)((m->y0 >> sh) + y) * dst_stride[k] + (m->x0 >> sh),
                       scr[k] + (size_t)((y0 >> sh) + y) * SCR_STRIDE + (x0 >> sh), (w >> sh) * sizeof(uint16_t));
        }
    }
    free(c);
    free(scr);
}
