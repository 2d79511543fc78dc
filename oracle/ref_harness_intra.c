/* oracle/ref_harness_intra.c -- TEST INFRASTRUCTURE ONLY (see ref_harness.c).
 *
 * Drives the reference's luma intra prediction the way the decoder does: CTUs in raster order, init_ctu_bitfield(_border)
 * at the start of each CTU (slicedec.c:731, :784), for every coding unit in decode order the rcn_funcs.intra_pred slot
 * (vvc_intra_pred, rcn_intra.c:484) on a "CTU buffer" that aliases the picture (so neighbours are the picture's
 * samples), then ctu_field_set_rect_bitfield (rcn_transform_tree.c:1419).  The two availability words the slot reads
 * (progress_field.vfield[x0 >> 2], .hfield[y0 >> 2]) are written back into the command -- what the integration shim's
 * recorder stores for the GPU.  Prediction only: with no residual the reconstruction of a block is its prediction.
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
#include "drv_utils.h"
#include "cu_utils.h"
#include "rcn_intra_angular.h"
#include "drv.h"

#include "oracle.h"

void rcn_init_intra_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_dc_planar_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_intra_angular_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_fill_ref_10(struct RCNFunctions *rcn_func);
void rcn_init_mip_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_tr_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_dequant_10(struct RCNFunctions *rcn_func);
void rcn_init_ict_functions_10(struct RCNFunctions *rcn_func, uint8_t type, uint8_t bitdepth);
void rcn_init_lfnst_functions(struct RCNFunctions *rcn_func);
uint64_t residual_coding_dpq(OVCTUDec *const ctu_dec, int16_t *const dst, uint8_t log2_tb_w, uint8_t log2_tb_h, uint16_t last_pos);
uint64_t residual_coding_sdh(OVCTUDec *const ctu_dec, int16_t *const dst, uint8_t log2_tb_w, uint8_t log2_tb_h, uint16_t last_pos);

/* private to rcn_transform_tree.c:51-76 and vcl_transform_unit.c -- must match field for field */
struct HTBInfo { uint16_t last_pos; uint64_t sig_sb_map; };
struct HISPTUInfo {
   uint8_t cbf_mask;
   uint8_t tr_skip_mask;
   uint8_t cu_mts_flag;
   uint8_t cu_mts_idx;
   uint8_t lfnst_flag;
   uint8_t lfnst_idx;
   struct HTBInfo tb_info[4];
};

/* One ISP coding unit through the reference's own slot (tmp.recon_isp_subtree_{v,h}): the whole block is marked in the
 * progress field first (vcl_transform_unit.c:1878), the coefficients of partition i sit at residual_y + (i << (log2 w +
 * log2 h)) (rcn_transform_tree.c:981, 1065).  The availability words the prediction reads are written back into `info`. */
static void
isp_cu(OVCTUDec *c, const OVOIntraCmd *m, OVOIspInfo *info, const int16_t *arena, int x0, int y0, int mts_enabled)
{
    int ver = (m->flags & OVO_INTRA_ISP_VER) != 0, lcw = m->log2_w, lch = m->log2_h;
    int lp = (ver ? lcw : lch) - 2, lo = ver ? lch : lcw;
    if (lo < 4 && lp <= 4 - lo) lp = 4 - lo;
    int nb = (1 << (ver ? lcw : lch)) >> lp, q = 0;
    struct HISPTUInfo tu;
    memset(&tu, 0, sizeof(tu));
    ctu_field_set_rect_bitfield(&c->rcn_ctx.progress_field, x0 >> 2, y0 >> 2, (1 << lcw) >> 2, (1 << lch) >> 2);
    info->common = ver ? c->rcn_ctx.progress_field.hfield[y0 >> 2] : c->rcn_ctx.progress_field.vfield[x0 >> 2];
    for (int i = 0; i < nb; i++) {
        int off = i << lp;
        if (ver) { if (!(off & 3)) info->part[q++] = c->rcn_ctx.progress_field.vfield[(x0 + off) >> 2]; }
        else info->part[q++] = c->rcn_ctx.progress_field.hfield[((y0 + off) >> 2) + !!((y0 + off) % 4)];
        const OVOTuCmd *t = &info->tu[i];
        if (!(t->cbf_mask & 0x10)) continue;
        tu.cbf_mask |= 1 << (nb - i - 1);
        tu.tb_info[i].last_pos = t->last_pos[2];
        tu.tb_info[i].sig_sb_map = t->sig_sb_map[2];
        if (t->flags & OVO_TU_LFNST_L) { tu.lfnst_flag = 1; tu.lfnst_idx = t->lfnst_idx; }
        c->dequant_luma.qp = t->qp[0];
        int tw = 1 << t->log2_w, th = 1 << t->log2_h;
        int n = (t->log2_w >= 2 && t->log2_h >= 2) ? (tw < 32 ? tw : 32) * (th < 32 ? th : 32) : tw * th;
        memcpy(c->residual_y + (i << (t->log2_w + t->log2_h)), arena + t->coeff_off[2], n * sizeof(int16_t));
    }
    c->mts_enabled = mts_enabled;
    if (ver) c->rcn_funcs.tmp.recon_isp_subtree_v(c, x0, y0, lcw, lch, m->mode, (const void *)&tu);
    else     c->rcn_funcs.tmp.recon_isp_subtree_h(c, x0, y0, lcw, lch, m->mode, (const void *)&tu);
}


/* rcn_ciip / rcn_ciip_b (rcn_inter.c:3012-3067) from the point where the inter prediction exists: the block's current
 * content plays tmp_inter (the MC stage wrote it), the reference's own slots do the rest -- intra_pred / intra_pred_c
 * with OVINTRA_PLANAR into the CTU buffer, ciip.weighted for the blend, mc_c.unidir[0][0] for the narrow-chroma copy
 * (rcn_ciip_weighted_sum, :2968-3010; the weight comes with the command, the shim derives it from part_map.cu_mode_x/y). */
static void
ciip_block(OVCTUDec *c, const OVOIntraCmd *m, int x0, int y0, uint16_t *const inter_pic[3], const int stride[3])
{
    /* The AVX2 blend works on 16-sample groups whatever the width (x86/rcn_mc_avx2.c:5024-5095): inside the reference that
     * spills into CTU-buffer samples that are rewritten later; here the CTU buffer aliases the picture, so the blend runs
     * on padded scratch copies and only the w x h block is stored back.  For the same reason (SIMD intra kernels of narrow
     * blocks also spill to the right) the inter prediction is taken from a copy of the input picture made up front. */
    enum { ST = 96 };
    static uint16_t inter[3][64 * ST + 32], intra[64 * ST + 32], out[64 * ST + 32];
    struct OVBuffInfo *b = &c->rcn_ctx.ctu_buff;
    uint16_t *dst[3] = { b->y + x0 + (ptrdiff_t)y0 * b->stride, b->cb + (x0 >> 1) + (ptrdiff_t)(y0 >> 1) * b->stride_c,
                         b->cr + (x0 >> 1) + (ptrdiff_t)(y0 >> 1) * b->stride_c };
    int w = 1 << m->log2_w, h = 1 << m->log2_h;
    for (int k = 0; k < 3; k++) {
        int bw = k ? w >> 1 : w, bh = k ? h >> 1 : h, st = k ? b->stride_c : b->stride;
        const uint16_t *src = inter_pic[k] + (size_t)(m->y0 >> (k != 0)) * stride[k] + (m->x0 >> (k != 0));
        (void)st;
        for (int y = 0; y < bh; y++) memcpy(&inter[k][y * ST], src + (size_t)y * stride[k], bw * 2);
    }
    c->rcn_funcs.intra_pred(&c->rcn_ctx, b, OVINTRA_PLANAR, x0, y0, m->log2_w, m->log2_h, 0);
    c->rcn_funcs.intra_pred_c(&c->rcn_ctx, OVINTRA_PLANAR, x0 >> 1, y0 >> 1, m->log2_w - 1, m->log2_h - 1, 0);
    for (int k = 0; k < 3; k++) {
        int bw = k ? w >> 1 : w, bh = k ? h >> 1 : h, st = k ? b->stride_c : b->stride;
        for (int y = 0; y < bh; y++) memcpy(&intra[y * ST], dst[k] + (ptrdiff_t)y * st, bw * 2);
        if (k && m->log2_w <= 2) c->rcn_funcs.mc_c.unidir[0][0](out, ST, inter[k], ST, bh, 0, 0, bw);
        else c->rcn_funcs.ciip.weighted(out, ST, intra, inter[k], ST, ST, bw, bh, m->ciip_wt);
        for (int y = 0; y < bh; y++) memcpy(dst[k] + (ptrdiff_t)y * st, &out[y * ST], bw * 2);
    }
}

void ovref_intra_picture2(uint16_t *const user_planes[3], const int user_stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth, int simd,
                          OVOIntraCmd *cmds, int nb_cmds, OVOIspInfo *isp, const int16_t *arena, int dep_quant, int mts_enabled);

/* Intra block copy through the reference's own slots (rcn_funcs.rcn_ibc_l / rcn_ibc_c).  The harness keeps the CTUs in the picture
 * planes, the reference copies inside its CTU ring buffer, so a ring is set up for the call: 256 * 128 >> log2_ctu_s luma columns
 * (rcn_ctu.c:558-560), slot (ctb_x & mask) = the current CTU as reconstructed so far, every other slot = the CTU decoded most
 * recently into it in this CTU row (slots no CTU of the row has reached yet stay zero); the copied block then goes back to the planes. */
static void
ibc_block(OVCTUDec *c, const OVOIntraCmd *m, uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int cx, int cy)
{
    const int ctu_s = 1 << log2_ctu_s, nslots = 256 * 128 >> (2 * log2_ctu_s), ring_w = nslots << log2_ctu_s;
    const int slot = cx & (nslots - 1), ctb_pos = slot << log2_ctu_s;
    struct OVBuffInfo saved = c->rcn_ctx.ctu_buff;
    uint16_t *ring[3];
    int rstride[3];
    for (int k = 0; k < 3; k++) {
        const int sh = k != 0;
        rstride[k] = (ring_w >> sh) + 64;
        ring[k] = calloc((size_t)rstride[k] * ((ctu_s >> sh) + 1), sizeof(uint16_t));
        for (int s = 0; s < nslots; s++) {
            const int src_cx = cx - ((slot - s) & (nslots - 1));
            if (src_cx < 0) continue;
            for (int y = 0; y < (ctu_s >> sh) && ((cy << log2_ctu_s) >> sh) + y < (pic_h >> sh); y++) {
                int n = (pic_w >> sh) - ((src_cx << log2_ctu_s) >> sh);
                if (n > (ctu_s >> sh)) n = ctu_s >> sh;
                memcpy(ring[k] + (size_t)y * rstride[k] + ((s << log2_ctu_s) >> sh),
                       planes[k] + (size_t)(((cy << log2_ctu_s) >> sh) + y) * stride[k] + ((src_cx << log2_ctu_s) >> sh), n * 2);
            }
        }
    }
    c->ctb_x = cx;
    c->rcn_ctx.ctu_buff.y = ring[0] + ctb_pos; c->rcn_ctx.ctu_buff.stride = rstride[0];
    c->rcn_ctx.ctu_buff.cb = ring[1] + (ctb_pos >> 1); c->rcn_ctx.ctu_buff.cr = ring[2] + (ctb_pos >> 1); c->rcn_ctx.ctu_buff.stride_c = rstride[1];
    IBCMV mv = { (int16_t)(m->pad[0] | (m->pad[1] << 8)), (int16_t)(m->pad[2] | (m->pad[3] << 8)) };
    const int x0 = m->x0 & (ctu_s - 1), y0 = m->y0 & (ctu_s - 1);
    if (m->flags & OVO_INTRA_LUMA) c->rcn_funcs.rcn_ibc_l(c, x0, y0, m->log2_w, m->log2_h, log2_ctu_s, mv);
    if (m->flags & OVO_INTRA_CHROMA) c->rcn_funcs.rcn_ibc_c(c, x0, y0, m->log2_w, m->log2_h, log2_ctu_s, mv);
    for (int k = 0; k < 3; k++) {
        const int sh = k != 0;
        if ((k == 0 && (m->flags & OVO_INTRA_LUMA)) || (k && (m->flags & OVO_INTRA_CHROMA)))
            for (int y = 0; y < (1 << m->log2_h) >> sh; y++)
                memcpy(planes[k] + (size_t)((m->y0 >> sh) + y) * stride[k] + (m->x0 >> sh),
                       ring[k] + (size_t)((y0 >> sh) + y) * rstride[k] + ((ctb_pos + x0) >> sh), ((1 << m->log2_w) >> sh) * 2);
        free(ring[k]);
    }
    c->rcn_ctx.ctu_buff = saved;
}

void
ovref_intra_picture(uint16_t *const user_planes[3], const int user_stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth, int simd,
                    OVOIntraCmd *cmds, int nb_cmds)
{
    ovref_intra_picture2(user_planes, user_stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd, cmds, nb_cmds, NULL, NULL, 1, 0);
}

void
ovref_intra_picture2(uint16_t *const user_planes[3], const int user_stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth, int simd,
                     OVOIntraCmd *cmds, int nb_cmds, OVOIspInfo *isp, const int16_t *arena, int dep_quant, int mts_enabled)
{
    /* SIMD kernels of 2- and 4-sample-wide blocks store whole vectors; the reference's CTU buffer has padding for that
     * (RCN_CTB_PADDING), so the harness works on copies of the planes with 64 samples of row padding */
    uint16_t *planes[3];
    int stride[3];
    for (int k = 0; k < 3; k++) {
        int ph = pic_h >> (k != 0), pw = pic_w >> (k != 0);
        stride[k] = pw + 64;
        planes[k] = calloc((size_t)stride[k] * (ph + 1), sizeof(uint16_t));
        for (int y = 0; y < ph; y++) memcpy(planes[k] + (size_t)y * stride[k], user_planes[k] + (size_t)y * user_stride[k], pw * 2);
    }
    OVCTUDec *c;
    OVPartInfo part;
    int ctu_s = 1 << log2_ctu_s;
    int ncw = (pic_w + ctu_s - 1) >> log2_ctu_s, nch = (pic_h + ctu_s - 1) >> log2_ctu_s;
    ctudec_init(&c);
    memset(&part, 0, sizeof(part));
    part.log2_ctu_s = log2_ctu_s; part.log2_min_cb_s = 2;
    c->part_ctx = &part;
    c->rcn_ctx.ctudec = c;
    c->pic_w = pic_w; c->pic_h = pic_h;
    int collocated = 0;
    for (int i = 0; i < nb_cmds; i++) collocated |= (cmds[i].flags & OVO_INTRA_CCLM_COLLOCATED) != 0;
    rcn_init_functions(&c->rcn_funcs, 0, 1, collocated, 0, bitdepth);
    if (!simd && bitdepth == 10) {      /* scalar C templates = arithmetic of record (rcn.c:98-117, 151-178) */
        extern const struct IntraAngularFunctions angular_gauss_h_10, angular_gauss_v_10, angular_cubic_h_10, angular_cubic_v_10,
                                                  angular_c_h_10, angular_c_v_10, angular_nofrac_v_10, angular_nofrac_h_10;
        rcn_init_intra_functions_10(&c->rcn_funcs);
        rcn_init_dc_planar_functions_10(&c->rcn_funcs);
        rcn_init_fill_ref_10(&c->rcn_funcs);
        rcn_init_mip_functions_10(&c->rcn_funcs);
        rcn_init_tr_functions_10(&c->rcn_funcs);
        rcn_init_dequant_10(&c->rcn_funcs);
        rcn_init_ict_functions_10(&c->rcn_funcs, 0, bitdepth);
        rcn_init_lfnst_functions(&c->rcn_funcs);
        c->rcn_funcs.intra_angular_gauss_h = &angular_gauss_h_10; c->rcn_funcs.intra_angular_gauss_v = &angular_gauss_v_10;
        c->rcn_funcs.intra_angular_cubic_h = &angular_cubic_h_10; c->rcn_funcs.intra_angular_cubic_v = &angular_cubic_v_10;
        c->rcn_funcs.intra_angular_c_h = &angular_c_h_10; c->rcn_funcs.intra_angular_c_v = &angular_c_v_10;
        c->rcn_funcs.intra_angular_nofrac_v = &angular_nofrac_v_10; c->rcn_funcs.intra_angular_nofrac_h = &angular_nofrac_h_10;
    }
    c->residual_coding_l = dep_quant ? &residual_coding_dpq : &residual_coding_sdh;
    uint16_t *inter_pic[3] = { NULL, NULL, NULL };
    for (int i = 0; i < nb_cmds; i++)
        if (((cmds[i].flags & OVO_INTRA_CIIP) || !cmds[i].flags) && !inter_pic[0])
            for (int k = 0; k < 3; k++) {
                size_t sz = (size_t)stride[k] * (pic_h >> (k != 0)) * sizeof(uint16_t);
                inter_pic[k] = malloc(sz);
                memcpy(inter_pic[k], planes[k], sz);
            }
    int n = 0;
    for (int cy = 0; cy < nch; cy++)
        for (int cx = 0; cx < ncw; cx++) {
            int ctu_w = pic_w - (cx << log2_ctu_s) < ctu_s ? pic_w - (cx << log2_ctu_s) : ctu_s;
            int ctu_h = pic_h - (cy << log2_ctu_s) < ctu_s ? pic_h - (cy << log2_ctu_s) : ctu_s;
            uint8_t flags = (cx ? CTU_LFT_FLG : 0) | (cy ? CTU_UP_FLG : 0) | (cx && cy ? CTU_UPLFT_FLG : 0) |
                            (cy && cx + 1 < ncw ? CTU_UPRGT_FLG : 0);
            c->ctu_ngh_flags = flags;
            if (ctu_w == ctu_s && ctu_h == ctu_s) init_ctu_bitfield(&c->rcn_ctx, flags, log2_ctu_s);
            else init_ctu_bitfield_border(&c->rcn_ctx, flags, log2_ctu_s, ctu_w, ctu_h);
            if (cx + 2 == ncw && (pic_w & (ctu_s - 1))) {          /* slicedec.c:733-737: the above-right CTU is a partial one */
                uint64_t mask = ((uint64_t)1 << ((((pic_w & (ctu_s - 1)) + ctu_s) >> 2) + 1)) - 1;
                c->rcn_ctx.progress_field_c.hfield[0] &= mask;
                c->rcn_ctx.progress_field.hfield[0] &= mask;
            }
            c->rcn_ctx.ctu_buff.y = planes[0] + (size_t)(cy << log2_ctu_s) * stride[0] + (cx << log2_ctu_s);
            c->rcn_ctx.ctu_buff.stride = stride[0];
            c->rcn_ctx.ctu_buff.cb = planes[1] + (size_t)(cy << (log2_ctu_s - 1)) * stride[1] + (cx << (log2_ctu_s - 1));
            c->rcn_ctx.ctu_buff.cr = planes[2] + (size_t)(cy << (log2_ctu_s - 1)) * stride[2] + (cx << (log2_ctu_s - 1));
            c->rcn_ctx.ctu_buff.stride_c = stride[1];
            while (n < nb_cmds && (cmds[n].x0 >> log2_ctu_s) == cx && (cmds[n].y0 >> log2_ctu_s) == cy) {
                OVOIntraCmd *m = &cmds[n++];
                int x0 = m->x0 & (ctu_s - 1), y0 = m->y0 & (ctu_s - 1);
                m->avail_cols = c->rcn_ctx.progress_field.vfield[x0 >> 2];
                m->avail_rows = c->rcn_ctx.progress_field.hfield[y0 >> 2];
                if (!m->flags) {
                    /* a CU that is not intra: its reconstruction lands in the CTU buffer now (and replaces whatever the
                     * vector stores of a narrow left neighbour spilled into it) */
                    for (int k = 0; k < 3; k++) {
                        int sh = k != 0, bw = (1 << m->log2_w) >> sh;
                        for (int y = 0; y < (1 << m->log2_h) >> sh; y++) {
                            size_t off = (size_t)((m->y0 >> sh) + y) * stride[k] + (m->x0 >> sh);
                            memcpy(planes[k] + off, inter_pic[k] + off, bw * 2);
                        }
                    }
                } else if (m->flags & OVO_INTRA_IBC) {
                    ibc_block(c, m, planes, stride, pic_w, pic_h, log2_ctu_s, cx, cy);
                } else if (m->flags & OVO_INTRA_CIIP) {
                    ciip_block(c, m, x0, y0, inter_pic, stride);
                } else if (m->flags & OVO_INTRA_ISP) {
                    /* isp_subtree_{v,h} (vcl_transform_unit.c:1344-1760): luma sub-tree, then the chroma blocks of the CU (rcn_tu_c) */
                    isp_cu(c, m, &isp[m->pad[0] | (m->pad[1] << 8)], arena, x0, y0, mts_enabled);
                    if (m->flags & OVO_INTRA_CHROMA)
                        c->rcn_funcs.intra_pred_c(&c->rcn_ctx, m->mode_c, x0 >> 1, y0 >> 1, m->log2_w - 1, m->log2_h - 1, flg_pred_mode_flag);
                } else {
                if ((m->flags & OVO_INTRA_LUMA) && (m->flags & OVO_INTRA_MIP))       /* rcn_intra_tu, rcn_transform_tree.c:1403-1406 */
                    c->rcn_funcs.mip.rcn_intra_mip(&c->rcn_ctx, x0, y0, m->log2_w, m->log2_h, m->mode);
                else if ((m->flags & OVO_INTRA_LUMA) && m->mrl_idx)       /* rcn_transform_tree.c:1413-1419 */
                    c->rcn_funcs.intra_pred_mrl(c, c->rcn_ctx.ctu_buff.y, c->rcn_ctx.ctu_buff.stride, m->mode, x0, y0, m->log2_w, m->log2_h, m->mrl_idx);
                else if (m->flags & OVO_INTRA_LUMA)
                    c->rcn_funcs.intra_pred(&c->rcn_ctx, &c->rcn_ctx.ctu_buff, m->mode, x0, y0, m->log2_w, m->log2_h, flg_pred_mode_flag);
                if (m->flags & OVO_INTRA_CHROMA) {
                    /* single tree: the chroma field is updated with the same rectangles as the luma one (rcn_tu_st, :1275) */
                    if (c->rcn_ctx.progress_field_c.vfield[x0 >> 2] != m->avail_cols || c->rcn_ctx.progress_field_c.hfield[y0 >> 2] != m->avail_rows) abort();
                    c->rcn_funcs.intra_pred_c(&c->rcn_ctx, m->mode_c, x0 >> 1, y0 >> 1, m->log2_w - 1, m->log2_h - 1, flg_pred_mode_flag);
                }
                }
                ctu_field_set_rect_bitfield(&c->rcn_ctx.progress_field, x0 >> 2, y0 >> 2, (1 << m->log2_w) >> 2, (1 << m->log2_h) >> 2);
                ctu_field_set_rect_bitfield(&c->rcn_ctx.progress_field_c, x0 >> 2, y0 >> 2, (1 << m->log2_w) >> 2, (1 << m->log2_h) >> 2);
            }
        }
    for (int k = 0; k < 3; k++) {
        int ph = pic_h >> (k != 0), pw = pic_w >> (k != 0);
        for (int y = 0; y < ph; y++) memcpy(user_planes[k] + (size_t)y * user_stride[k], planes[k] + (size_t)y * stride[k], pw * 2);
        free(planes[k]);
        free(inter_pic[k]);
    }
    free(c);
}
