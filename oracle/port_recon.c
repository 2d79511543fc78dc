/* port_recon.c -- TEST INFRASTRUCTURE ONLY (oracle; never linked into the product).
 *
 * Reconstruction of a picture in the reference's CU order: for every coding unit, in decode order, the prediction and
 * then the residual, exactly as coding_unit() drives it -- intra CU in a single tree: rcn_intra_tu (luma prediction,
 * libovvc/rcn_transform_tree.c:1392-1431), then rcn_tu_st (:1228-1300): luma residual, chroma prediction (intra_pred_c,
 * after the luma residual so that CCLM reads reconstructed luma), chroma residual.  CIIP: rcn_ciip blends both
 * components (libovvc/rcn_inter.c:3012-3067) before the transform tree.  Inter CUs: their prediction is already in the
 * picture (motion compensation); only the residual is added.
 * Built from the two pinned pieces ovo_intra_picture / ovo_residual_picture, one command at a time. */
#include <string.h>
#include "oracle.h"

/* cmds[i] and tus[i] describe the same coding unit (intra flags == 0: not intra; cbf_mask == 0: no residual) */
void ovo_recon_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                       const OVOIntraCmd *cmds, const OVOTuCmd *tus, int nb_cu, const int16_t *arena,
                       int dep_quant, int sh_ts_disabled, int ict_type,
                       const OVOLmcsAnchor *anchors, int nb_anchors, const OVOLmcsLuts *luts, const OVOLmcsData *lmcs)
{
    ovo_recon_picture2(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, cmds, tus, nb_cu, arena, dep_quant, sh_ts_disabled, ict_type,
                       anchors, nb_anchors, luts, lmcs, NULL);
}

void ovo_recon_picture2(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                        const OVOIntraCmd *cmds, const OVOTuCmd *tus, int nb_cu, const int16_t *arena,
                        int dep_quant, int sh_ts_disabled, int ict_type,
                        const OVOLmcsAnchor *anchors, int nb_anchors, const OVOLmcsLuts *luts, const OVOLmcsData *lmcs,
                        const OVOIspInfo *isp)
{
    int next_anchor = 0, cur_scale = 1 << 11;
    for (int i = 0; i < nb_cu; i++) {
        OVOIntraCmd c = cmds[i];
        OVOTuCmd t = tus[i], tl, tc;
        /* coding_unit (vcl_coding_unit.c:725-730): the chroma scale is refreshed when a CU starts on a 64-aligned position,
         * from the CTU buffer as it is then, and stays in force until the next refresh */
        if (next_anchor < nb_anchors && anchors[next_anchor].x0 == c.x0 && anchors[next_anchor].y0 == c.y0) {
            const OVOLmcsAnchor *a = &anchors[next_anchor++];
            cur_scale = ovo_lmcs_chroma_scale(planes[0] + (size_t)a->y0 * stride[0] + a->x0, stride[0], a->abv_mask, a->lft_mask,
                                              luts, lmcs, bitdepth);
        }
        if (t.flags & 8) t.chroma_scale = (int16_t)cur_scale;
        tl = t; tc = t;
        tl.cbf_mask = t.cbf_mask & 0x10;
        tc.cbf_mask = t.cbf_mask & 0x0f;
        if (c.flags & OVO_INTRA_CIIP) {
            ovo_intra_picture(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, &c, 1);
            if (t.cbf_mask) ovo_residual_picture(planes, stride, pic_w, pic_h, bitdepth, &t, 1, arena, dep_quant, sh_ts_disabled, ict_type);
            continue;
        }
        if (c.flags & OVO_INTRA_ISP) {
            ovo_isp_cu(planes[0], stride[0], bitdepth, (1 << log2_ctu_s) - 1, &c, &isp[c.pad[0] | (c.pad[1] << 8)], arena, dep_quant);
            tl.cbf_mask = 0;
            c.flags &= ~(OVO_INTRA_ISP | OVO_INTRA_ISP_VER | OVO_INTRA_LUMA);
        }
        if (c.flags & OVO_INTRA_LUMA) {
            OVOIntraCmd l = c;
            l.flags &= ~OVO_INTRA_CHROMA;
            ovo_intra_picture(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, &l, 1);
        }
        if (tl.cbf_mask) ovo_residual_picture(planes, stride, pic_w, pic_h, bitdepth, &tl, 1, arena, dep_quant, sh_ts_disabled, ict_type);
        if (c.flags & OVO_INTRA_CHROMA) {
            OVOIntraCmd k = c;
            k.flags &= ~OVO_INTRA_LUMA;
            ovo_intra_picture(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, &k, 1);
        }
        if (tc.cbf_mask) ovo_residual_picture(planes, stride, pic_w, pic_h, bitdepth, &tc, 1, arena, dep_quant, sh_ts_disabled, ict_type);
    }
}
