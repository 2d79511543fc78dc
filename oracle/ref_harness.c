/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin driver around the UNMODIFIED OpenVVC reference (linked as oracle/_ref/libovvc_ref.so): it builds
 * the minimum decoder state each reference entry point reads, then calls the
 * reference's own function through its RCNFunctions table.  No reference code is restated here; the
 * arithmetic all runs inside the reference.  Compiled by oracle/Makefile with -I/root/reference/libovvc
 * into oracle/_ref/libovref_harness.so (git-ignored, travels to the GPU box as a prebuilt file).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "ovdefs.h"
#include "ovframe.h"
#include "ctudec.h"
#include "dec_structures.h"
#include "nvcl_structures.h"
#include "rcn.h"
#include "rcn_alf.h"
#include "rcn_structures.h"

#include "oracle.h"

/* C-only init functions of the 10-bit template (rcn.c:151-178) used to force the scalar path */
void rcn_init_alf_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_sao_functions_10(struct RCNFunctions *rcn_func);

int ovref_harness_version(void) { return 2; }

typedef struct HarnessCtx {
    OVCTUDec *ctudec;
    OVPartInfo part;
    OVFrame frame;
} HarnessCtx;

static HarnessCtx *
harness_open(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s,
             int bitdepth, int simd)
{
    HarnessCtx *h = calloc(1, sizeof(*h));
    if (ctudec_init(&h->ctudec) < 0) { free(h); return NULL; }
    OVCTUDec *c = h->ctudec;

    memset(&h->part, 0, sizeof(h->part));
    h->part.log2_ctu_s    = log2_ctu_s;
    h->part.log2_min_cb_s = 2;
    c->part_ctx = &h->part;
    c->rcn_ctx.ctudec = c;
    c->pic_w = pic_w;
    c->pic_h = pic_h;

    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 0, bitdepth);
    if (!simd && bitdepth == 10) {
        rcn_init_alf_functions_10(&c->rcn_funcs);
        rcn_init_sao_functions_10(&c->rcn_funcs);
    }

    for (int i = 0; i < 3; i++) {
        h->frame.data[i]     = planes[i];
        h->frame.linesize[i] = (size_t)stride[i] * sizeof(uint16_t);
    }
    h->frame.width  = pic_w;
    h->frame.height = pic_h;
    c->rcn_ctx.frame_start = &h->frame;

    int nb_ctu_w = (pic_w + (1 << log2_ctu_s) - 1) >> log2_ctu_s;
    /* margin 3 as in slicedec.c:1319-1320 */
    c->rcn_funcs.rcn_alloc_filter_buffers(&c->rcn_ctx, nb_ctu_w, 3, log2_ctu_s);
    return h;
}

static void
harness_close(HarnessCtx *h)
{
    OVCTUDec *c = h->ctudec;
    /* the in-loop parameter arrays belong to the caller */
    c->alf_info.ctb_alf_params = NULL;
    c->alf_info.ctb_cc_alf_filter_idx[0] = NULL;
    c->alf_info.ctb_cc_alf_filter_idx[1] = NULL;
    c->sao_info.sao_params = NULL;
    c->rcn_funcs.rcn_buff_uninit(&c->rcn_ctx);
    free(c);
    free(h);
}

static void
aps_to_ref(OVALFData *d, const OVOAlfAps *a)
{
    memset(d, 0, sizeof(*d));
    d->alf_luma_clip_flag = a->luma_clip_flag;
    d->alf_luma_num_filters_signalled_minus1 = a->num_luma_filters - 1;
    memcpy(d->alf_luma_coeff_delta_idx, a->luma_coeff_delta_idx, 25);
    memcpy(d->alf_luma_coeff, a->luma_coeff, sizeof(d->alf_luma_coeff));
    memcpy(d->alf_luma_clip_idx, a->luma_clip_idx, sizeof(d->alf_luma_clip_idx));
    d->alf_chroma_clip_flag = a->chroma_clip_flag;
    d->alf_chroma_num_alt_filters_minus1 = a->num_chroma_alts - 1;
    memcpy(d->alf_chroma_coeff, a->chroma_coeff, sizeof(d->alf_chroma_coeff));
    memcpy(d->alf_chroma_clip_idx, a->chroma_clip_idx, sizeof(d->alf_chroma_clip_idx));
}

/* Runs the reference's alf.rcn_alf_reconstruct_coeff_APS (rcn_alf.c:264) and copies RCNALF out. */
void
ovref_alf_build_tables(OVOAlfTables *t, const OVOAlfAps *luma_aps, int nb_luma_aps,
                       const OVOAlfAps *chroma_aps, int bitdepth)
{
    OVCTUDec *c;
    OVALFData *d = calloc(9, sizeof(*d));
    ctudec_init(&c);
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 0, bitdepth);
    c->alf_info.num_alf_aps_ids_luma = nb_luma_aps;
    for (int i = 0; i < nb_luma_aps; i++) {
        aps_to_ref(&d[i], &luma_aps[i]);
        c->alf_info.aps_alf_data[i] = &d[i];
    }
    if (chroma_aps) {
        aps_to_ref(&d[8], chroma_aps);
        c->alf_info.aps_alf_data_c = &d[8];
    }
    c->rcn_funcs.alf.rcn_alf_reconstruct_coeff_APS(&c->alf_info.rcn_alf, c, 1, chroma_aps != NULL);
    memcpy(t->luma_coeff, c->alf_info.rcn_alf.filter_coeff_dec, sizeof(t->luma_coeff));
    memcpy(t->luma_clip,  c->alf_info.rcn_alf.filter_clip_dec,  sizeof(t->luma_clip));
    memcpy(t->chroma_coeff, c->alf_info.rcn_alf.chroma_coeff_final, sizeof(t->chroma_coeff));
    memcpy(t->chroma_clip,  c->alf_info.rcn_alf.chroma_clip_final,  sizeof(t->chroma_clip));
    free(c);
    free(d);
}

/* In-place whole-picture ALF through alf.rcn_alf_filter_line (rcn_alf.c:1285), rows in decode order. */
void
ovref_alf_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h,
                  int log2_ctu_s, int bitdepth, int simd,
                  const OVOAlfTables *tab, const OVOAlfCtuParams *p)
{
    HarnessCtx *h = harness_open(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd);
    OVCTUDec *c = h->ctudec;
    struct ALFInfo *ai = &c->alf_info;
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    int nb_h = (pic_h + ctu_s - 1) >> log2_ctu_s;
    int n = nb_w * nb_h;
    ALFParamsCtu *params = calloc(n, sizeof(*params));
    OVALFData cc[2];
    memset(cc, 0, sizeof(cc));

    memcpy(ai->rcn_alf.filter_coeff_dec, tab->luma_coeff, sizeof(tab->luma_coeff));
    memcpy(ai->rcn_alf.filter_clip_dec,  tab->luma_clip,  sizeof(tab->luma_clip));
    memcpy(ai->rcn_alf.chroma_coeff_final, tab->chroma_coeff, sizeof(tab->chroma_coeff));
    memcpy(ai->rcn_alf.chroma_clip_final,  tab->chroma_clip,  sizeof(tab->chroma_clip));
    memcpy(cc[0].alf_cc_mapped_coeff[0], tab->cc_coeff[0], sizeof(tab->cc_coeff[0]));
    memcpy(cc[1].alf_cc_mapped_coeff[1], tab->cc_coeff[1], sizeof(tab->cc_coeff[1]));

    for (int i = 0; i < n; i++) {
        params[i].ctb_alf_flag   = p->ctb_alf_flag[i];
        params[i].ctb_alf_idx    = p->ctb_alf_idx[i];
        params[i].cb_alternative = p->cb_alternative[i];
        params[i].cr_alternative = p->cr_alternative[i];
    }
    ai->alf_luma_enabled_flag = 1;
    ai->alf_cb_enabled_flag = 1;
    ai->alf_cr_enabled_flag = 1;
    ai->cc_alf_cb_enabled_flag = p->cc_cb_enabled;
    ai->cc_alf_cr_enabled_flag = p->cc_cr_enabled;
    ai->aps_cc_alf_data_cb = &cc[0];
    ai->aps_cc_alf_data_cr = &cc[1];
    ai->ctb_alf_params = params;
    ai->ctb_cc_alf_filter_idx[0] = (uint8_t *)p->cc_idx_cb;
    ai->ctb_cc_alf_filter_idx[1] = (uint8_t *)p->cc_idx_cr;

    struct RectEntryInfo e;
    memset(&e, 0, sizeof(e));
    e.nb_ctu_w = nb_w;
    e.nb_ctu_h = nb_h;
    e.nb_ctu_rect = n;
    e.nb_ctb_pic_w = nb_w;

    for (int row = 0; row < nb_h; row++)
        c->rcn_funcs.alf.rcn_alf_filter_line(c, &e, row);

    harness_close(h);
    free(params);
}

/* In-place whole-picture SAO following the call order of decode_ctu_line's tail (slicedec.c:934-956). */
void
ovref_sao_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h,
                  int log2_ctu_s, int bitdepth, int simd, const OVOSaoCtu *p)
{
    HarnessCtx *h = harness_open(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd);
    OVCTUDec *c = h->ctudec;
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    int nb_h = (pic_h + ctu_s - 1) >> log2_ctu_s;
    int n = nb_w * nb_h;
    SAOParamsCtu *params = calloc(n, sizeof(*params));

    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            memcpy(params[i].offset_val[k], p[i].offset_val[k], sizeof(params[i].offset_val[k]));
            params[i].band_position[k] = p[i].band_position[k];
            params[i].eo_class[k]      = p[i].eo_class[k];
            params[i].type_idx[k]      = p[i].type_idx[k];
        }
    c->sao_info.sao_luma_flag = 1;
    c->sao_info.sao_chroma_flag = 1;
    c->sao_info.chroma_format_idc = 1;
    c->sao_info.sao_params = params;

    struct RectEntryInfo e;
    memset(&e, 0, sizeof(e));
    e.nb_ctu_w = nb_w;
    e.nb_ctu_h = nb_h;
    e.nb_ctu_rect = n;
    e.nb_ctb_pic_w = nb_w;

    const struct SAOFunctions *sao = &c->rcn_funcs.sao;
    for (int row = 0; row < nb_h; row++) {
        if (row == 0) {
            sao->rcn_sao_first_pix_rows(c, &e, 0);
            if (nb_h == 1) sao->rcn_sao_filter_line(c, &e, 0);
        } else if (row == nb_h - 1) {
            sao->rcn_sao_filter_line(c, &e, row - 1);
            sao->rcn_sao_filter_line(c, &e, row);
        } else {
            sao->rcn_sao_filter_line(c, &e, row - 1);
        }
    }

    harness_close(h);
    free(params);
}

/* ===================================================================================================
 * Motion compensation through the reference's coarse slots rcn_mcp_b / rcn_mcp_b_l / rcn_mcp_b_c
 * (rcn_structures.h:650-676).  The destination OVBuffInfo points at the CTU origin inside the output
 * picture so that the reference, which writes CTU-relative, lands on picture coordinates.
 * =================================================================================================== */
#include "ovdpb.h"
#include "rcn_lmcs.h"

void rcn_init_mc_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_prof_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_bdof_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_dmvr_functions_10(struct RCNFunctions *rcn_func);

static void harness_no_synchro(const OVPicture *const ref_pic, int a, int b, int c, int d) { (void)ref_pic; (void)a; (void)b; (void)c; (void)d; }

/* the reference's own slot on a synthetic CTU state: the picture plays the CTU buffer, the two availability words are
 * placed in the progress field where the slot looks them up (rcn_lmcs.c:325-331) */
int
ovref_lmcs_chroma_scale(const uint16_t *src, int stride, unsigned abv_mask, unsigned lft_mask,
                        const OVOLmcsLuts *luts, const OVOLmcsData *d, int bitdepth)
{
    OVCTUDec *c;
    OVLMCSData data;
    struct CTUBitField pf;
    int ret;
    ctudec_init(&c);
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 1, bitdepth);
    memset(&data, 0, sizeof(data));
    data.lmcs_min_bin_idx = d->lmcs_min_bin_idx;
    data.lmcs_delta_max_bin_idx = d->lmcs_delta_max_bin_idx;
    data.lmcs_delta_cw_prec_minus1 = d->lmcs_delta_cw_prec_minus1;
    memcpy(data.lmcs_delta_abs_cw, d->lmcs_delta_abs_cw, sizeof(data.lmcs_delta_abs_cw));
    memcpy(data.lmcs_delta_sign_cw_flag, d->lmcs_delta_sign_cw_flag, sizeof(data.lmcs_delta_sign_cw_flag));
    data.lmcs_delta_abs_crs = d->lmcs_delta_abs_crs;
    data.lmcs_delta_sign_crs_flag = d->lmcs_delta_sign_crs_flag;
    c->rcn_funcs.rcn_init_lmcs(&c->lmcs_info, &data);
    if (memcmp(c->lmcs_info.luts, luts, sizeof(*luts))) abort();      /* same LUTs as the caller's */
    memset(&pf, 0, sizeof(pf));
    pf.hfield[0] = (uint64_t)abv_mask << 1;
    pf.vfield[0] = (uint64_t)lft_mask << 1;
    c->rcn_funcs.rcn_lmcs_compute_chroma_scale(&c->lmcs_info, (int16_t)stride, &pf, src, 0, 0);
    ret = (int16_t)c->lmcs_info.lmcs_chroma_scale;
    free(c->lmcs_info.luts);
    free(c);
    return ret;
}

/* the reference's own lmcs_reshape_backward slot (rcn_lmcs.c:295; called per CTU by slicedec.c:745-750) over a whole luma plane:
 * CTU by CTU like the decoder does, with the caller's LUTs in the slot's private struct LMCSLUTs layout */
void
ovref_lmcs_reshape_backward_plane(uint16_t *plane, int stride, int w, int h, const OVOLmcsLuts *luts, int log2_ctu, int bitdepth)
{
    OVCTUDec *c;
    const int ctu = 1 << log2_ctu;
    ctudec_init(&c);
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 1, bitdepth);
    for (int y = 0; y < h; y += ctu)
        for (int x = 0; x < w; x += ctu)
            c->rcn_funcs.lmcs_reshape_backward(plane + (size_t)y * stride + x, stride, (const struct LMCSLUTs *)luts,
                                               x + ctu <= w ? ctu : w - x, y + ctu <= h ? ctu : h - y);
    free(c);
}

void
ovref_lmcs_build_luts(OVOLmcsLuts *out, const OVOLmcsData *d, int bitdepth)
{
    OVCTUDec *c;
    OVLMCSData data;
    ctudec_init(&c);
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 1, bitdepth);
    memset(&data, 0, sizeof(data));
    data.lmcs_min_bin_idx = d->lmcs_min_bin_idx;
    data.lmcs_delta_max_bin_idx = d->lmcs_delta_max_bin_idx;
    data.lmcs_delta_cw_prec_minus1 = d->lmcs_delta_cw_prec_minus1;
    memcpy(data.lmcs_delta_abs_cw, d->lmcs_delta_abs_cw, sizeof(data.lmcs_delta_abs_cw));
    memcpy(data.lmcs_delta_sign_cw_flag, d->lmcs_delta_sign_cw_flag, sizeof(data.lmcs_delta_sign_cw_flag));
    data.lmcs_delta_abs_crs = d->lmcs_delta_abs_crs;
    data.lmcs_delta_sign_crs_flag = d->lmcs_delta_sign_crs_flag;
    c->rcn_funcs.rcn_init_lmcs(&c->lmcs_info, &data);
    /* struct LMCSLUTs (rcn_lmcs.c:77-82) is private to the reference: fwd[1<<BD], bwd[1<<BD], bnd[17] */
    memcpy(out, c->lmcs_info.luts, sizeof(*out));
    free(c->lmcs_info.luts);
    free(c);
}

/* sc != NULL: reference picture resampling -- the references have their own sizes and the slice's scale factors
 * (inter_ctx->scale_fact_rpl0/1, ctudec.c:40-84) send the stock slots down their rcn_mcp_rpr_* paths (rcn_inter.c:1990-2747) */
static void
mc_picture_impl(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int log2_ctu_s,
                int bitdepth, int simd, const OVOMcRefs *refs, const OVOMcScaling *sc, const OVOMcCmd *cmds, int nb_cmds,
                const OVOProfInfo *prof_tab, const OVOLmcsLuts *lmcs, OVODmvrMv *dmvr_out)
{
    OVCTUDec *c;
    OVPartInfo part;
    /* heap, not static: bench.py drives one picture per host thread */
    OVPicture (*pics)[16] = calloc(2, sizeof(*pics));
    OVFrame (*frames)[16] = calloc(2, sizeof(*frames));
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
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, lmcs != NULL, bitdepth);
    if (!simd && bitdepth == 10) {      /* scalar C templates = arithmetic of record */
        rcn_init_mc_functions_10(&c->rcn_funcs);
        rcn_init_prof_functions_10(&c->rcn_funcs);
        rcn_init_bdof_functions_10(&c->rcn_funcs);
        rcn_init_dmvr_functions_10(&c->rcn_funcs);
    }
    if (lmcs) {
        c->lmcs_info.luts = malloc(sizeof(*lmcs));
        memcpy(c->lmcs_info.luts, lmcs, sizeof(*lmcs));
        c->lmcs_info.lmcs_enabled_flag = 1;
    }
    struct InterDRVCtx *ic = &c->drv_ctx.inter_ctx;
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            /* rcn_mcp_rpr_c reads rpl0[ref_idx]->scale_info whatever the list (rcn_inter.c:2305): every entry gets a picture */
            if (sc) {
                pics[l][i].scale_info.chroma_hor_col_flag = sc->chroma_hor_col;
                pics[l][i].scale_info.chroma_ver_col_flag = sc->chroma_ver_col;
                if (l == 0) ic->rpl0[i] = &pics[l][i]; else ic->rpl1[i] = &pics[l][i];
            }
            if (!refs->plane[l][i][0]) continue;
            OVFrame *f = &frames[l][i];
            for (int k = 0; k < 3; k++) {
                f->data[k] = (void *)refs->plane[l][i][k];
                f->linesize[k] = (size_t)refs->stride[l][i][k] * sizeof(uint16_t);
            }
            f->width = sc ? sc->ref_w[l][i] : pic_w; f->height = sc ? sc->ref_h[l][i] : pic_h;
            pics[l][i].frame = f;
            pics[l][i].poc = refs->poc[l][i];
            atomic_init(&pics[l][i].idx_function, 0);
            pics[l][i].ovdpb_frame_synchro[0] = harness_no_synchro;
            pics[l][i].ovdpb_frame_synchro[1] = harness_no_synchro;
            if (l == 0) { ic->rpl0[i] = &pics[l][i]; ic->scale_fact_rpl0[i][0] = ic->scale_fact_rpl0[i][1] = 1 << 14; }
            else        { ic->rpl1[i] = &pics[l][i]; ic->scale_fact_rpl1[i][0] = ic->scale_fact_rpl1[i][1] = 1 << 14; }
            if (sc)
                for (int k = 0; k < 2; k++) {
                    if (l == 0) ic->scale_fact_rpl0[i][k] = sc->scale[l][i][k];
                    else        ic->scale_fact_rpl1[i][k] = sc->scale[l][i][k];
                }
        }

    for (int n = 0; n < nb_cmds; n++) {
        const OVOMcCmd *m = &cmds[n];
        int ctu_s = 1 << log2_ctu_s;
        c->ctb_x = m->x0 >> log2_ctu_s;
        c->ctb_y = m->y0 >> log2_ctu_s;
        int x0 = m->x0 & (ctu_s - 1), y0 = m->y0 & (ctu_s - 1);
        /* The reference writes into its CTU-local buffer (stride 324, rcn_ctu.c:558) and its SIMD kernels may
         * store a few samples past a narrow block; give it the same kind of padded scratch and copy the PU out. */
        struct OVBuffInfo b;
        memset(&b, 0, sizeof(b));
        b.stride = SCR_STRIDE;
        b.stride_c = SCR_STRIDE;
        b.y = scr[0]; b.cb = scr[1]; b.cr = scr[2];
        OVMV mv0 = { .x = m->mv0_x, .y = m->mv0_y, .ref_idx = m->ref_idx0, .bcw_idx_plus1 = m->bcw_idx_plus1 };
        OVMV mv1 = { .x = m->mv1_x, .y = m->mv1_y, .ref_idx = m->ref_idx1, .bcw_idx_plus1 = m->bcw_idx_plus1 };
        ic->prec_amvr = (m->flags & OVO_MC_HPEL_AMVR) ? 3 /* MV_PRECISION_HALF, drv_utils.h:102 */ : 0;
        if (m->tool == OVO_MC_TOOL_DMVR) {
            c->rcn_funcs.rcn_dmvr_mv_refine(c, b, x0, y0, m->log2_w, m->log2_h, &mv0, &mv1, m->ref_idx0, m->ref_idx1, m->aux & 1);
            if (dmvr_out) { dmvr_out->mv0_x = mv0.x; dmvr_out->mv0_y = mv0.y; dmvr_out->mv1_x = mv1.x; dmvr_out->mv1_y = mv1.y; dmvr_out++; }
        } else if (m->tool == OVO_MC_TOOL_BDOF)
            c->rcn_funcs.rcn_bdof_mcp_l(c, b, x0, y0, m->log2_w, m->log2_h, mv0, mv1, m->ref_idx0, m->ref_idx1);
        else if (m->tool == OVO_MC_TOOL_GPM) {
            static int gpm_init = 0;
            if (!gpm_init) { rcn_init_gpm_params(); gpm_init = 1; }          /* process-global tables, ovdec.c */
            struct VVCGPM g;
            memset(&g, 0, sizeof(g));
            g.split_dir = m->aux & 63; g.inter_dir0 = 1 + ((m->aux >> 8) & 1); g.inter_dir1 = 1 + ((m->aux >> 9) & 1);
            g.mv0 = mv0; g.mv1 = mv1;
            c->rcn_ctx.ctu_buff = b;                                           /* rcn_gpm_b writes to the CTU buffer */
            c->rcn_funcs.rcn_gpm_b(c, &g, x0, y0, m->log2_w, m->log2_h);
        } else if (m->tool == OVO_MC_TOOL_PROF)      /* struct PROFInfo (rcn_inter.c:1128) has OVOProfInfo's layout */
            c->rcn_funcs.rcn_prof_mcp_b_l(c, b, ic, &part, mv0, mv1, x0, y0, 2, 2, m->inter_dir, m->ref_idx0, m->ref_idx1,
                                          m->aux & 3, (const void *)&prof_tab[m->aux >> 8]);
        else if ((m->flags & 3) == 3)
            c->rcn_funcs.rcn_mcp_b(c, b, ic, &part, mv0, mv1, x0, y0, m->log2_w, m->log2_h, m->inter_dir, m->ref_idx0, m->ref_idx1);
        else if (m->flags & OVO_MC_LUMA)
            c->rcn_funcs.rcn_mcp_b_l(c, b, ic, &part, mv0, mv1, x0, y0, m->log2_w, m->log2_h, m->inter_dir, m->ref_idx0, m->ref_idx1);
        else if (m->flags & OVO_MC_CHROMA)
            c->rcn_funcs.rcn_mcp_b_c(c, b, ic, &part, mv0, mv1, x0, y0, m->log2_w, m->log2_h, m->inter_dir, m->ref_idx0, m->ref_idx1);
        for (int k = 0; k < 3; k++) {
            int sh = k ? 1 : 0;
            if ((k == 0 && !(m->flags & OVO_MC_LUMA)) || (k && !(m->flags & OVO_MC_CHROMA))) continue;
            int bw = (1 << m->log2_w) >> sh, bh = (1 << m->log2_h) >> sh;
            for (int y = 0; y < bh; y++)
                memcpy(dst[k] + (size_t)((m->y0 >> sh) + y) * dst_stride[k] + (m->x0 >> sh),
                       scr[k] + (size_t)((y0 >> sh) + y) * SCR_STRIDE + (x0 >> sh), bw * sizeof(uint16_t));
        }
    }
    if (lmcs) free(c->lmcs_info.luts);
    free(c);
    free(pics); free(frames); free(scr);
}

void
ovref_mc_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int log2_ctu_s,
                 int bitdepth, int simd, const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const OVOProfInfo *prof_tab,
                 const OVOLmcsLuts *lmcs, OVODmvrMv *dmvr_out)
{
    mc_picture_impl(dst, dst_stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd, refs, NULL, cmds, nb_cmds, prof_tab, lmcs, dmvr_out);
}

void
ovref_mc_picture_rpr(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int log2_ctu_s,
                     int bitdepth, int simd, const OVOMcRefs *refs, const OVOMcScaling *sc, const OVOMcCmd *cmds, int nb_cmds,
                     const OVOProfInfo *prof_tab, const OVOLmcsLuts *lmcs)
{
    mc_picture_impl(dst, dst_stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd, refs, sc, cmds, nb_cmds, prof_tab, lmcs, NULL);
}

/* Taps of the resampling filters (ov_mc_filters_rpr / ov_mc_filters_rpr_c, static in rcn_mc.c:165-379), read back through the
 * scalar mc_l / mc_c .rpr_h[1][1] slots by impulse probing: a sample of value 4 at tap k gives taps[k] * 4 >> (10 - 8). */
int
ovref_rpr_taps(int chroma, int filter_idx, int frac, int16_t *taps)
{
    struct RCNFunctions f;
    memset(&f, 0, sizeof(f));
    rcn_init_mc_functions_10(&f);
    const int nt = chroma ? 4 : 8, before = chroma ? 1 : 3;
    if (filter_idx < 0 || filter_idx >= (chroma ? 3 : 6) || frac < 0 || frac >= (chroma ? 32 : 16)) return -1;
    for (int k = 0; k < nt; k++) {
        uint16_t src[16] = { 0 }, out[4] = { 0 };
        src[4 - before + k] = 4;
        (chroma ? f.mc_c.rpr_h[1][1] : f.mc_l.rpr_h[1][1])(out, 4, src + 4, 16, 1, frac, 0, 1, (uint8_t)filter_idx);
        taps[k] = (int16_t)out[0];
    }
    return 0;
}

/* the reference's output-stage up-scaler on one plane (pp_sample_rate_conv is exported: post_proc.h:68) */
void pp_sample_rate_conv(uint16_t *scaled_dst, uint16_t scaled_stride, int scaledWidth, int scaledHeight, uint16_t *orgSrc, uint16_t org_stride,
                         int orgWidth, int orgHeight, const struct ScalingInfo *const scale_info, uint8_t luma_flag);
void
ovref_pp_upscale_plane(uint16_t *dst, int dst_stride, int dst_w, int dst_h, const uint16_t *src, int src_stride, int src_w, int src_h,
                       const OVOScalingWindow *win, int luma)
{
    struct ScalingInfo si;
    memset(&si, 0, sizeof(si));
    si.scaling_win_left = win->left; si.scaling_win_right = win->right; si.scaling_win_top = win->top; si.scaling_win_bottom = win->bottom;
    si.chroma_hor_col_flag = win->chroma_hor_col; si.chroma_ver_col_flag = win->chroma_ver_col;
    pp_sample_rate_conv(dst, (uint16_t)dst_stride, dst_w, dst_h, (uint16_t *)src, (uint16_t)src_stride, src_w, src_h, &si, (uint8_t)luma);
}

/* the reference's film grain synthesis (pp_film_grain.c; both functions are exported, post_proc.h) */
void fg_data_base_generation(uint8_t enableDeblocking);
void fg_simulate_grain_blk8x8(int32_t *grainStripe, uint32_t grainStripeOffsetBlk8, uint32_t width, uint8_t log2ScaleFactor, int16_t scaleFactor,
                              uint32_t kOffset, uint32_t lOffset, uint8_t h, uint8_t v, uint32_t xSize);
void fg_grain_apply_pic(int16_t **dstComp, int16_t **srcComp, struct OVSEIFGrain *fgrain, int pic_w, int pic_h, int poc, uint8_t isIdrPic, uint8_t enableDeblocking);

/* its pattern data base (a static array), read back 8x8 window by window through fg_simulate_grain_blk8x8 with a unit scale */
void
ovref_fg_database(int8_t *db, int deblock)
{
    fg_data_base_generation((uint8_t)deblock);
    for (int h = 0; h < 13; h++)
        for (int v = 0; v < 13; v++)
            for (int lo = 0; lo < 64; lo += 8)
                for (int ko = 0; ko < 64; ko += 8) {
                    int32_t blk[64];
                    fg_simulate_grain_blk8x8(blk, 0, 8, 0, 64, ko, lo, (uint8_t)h, (uint8_t)v, 8);
                    for (int l = 0; l < 8; l++)
                        for (int k = 0; k < 8; k++) db[(size_t)(h * 13 + v) * 4096 + (lo + l) * 64 + ko + k] = (int8_t)blk[l * 8 + k];
                }
}

void
ovref_fg_apply(uint16_t *const dst[3], const uint16_t *const src[3], int w, int h, const OVOFilmGrain *fg, int32_t poc)
{
    struct OVSEIFGrain sei;
    memset(&sei, 0, sizeof(sei));
    sei.fg_log2_scale_factor = fg->log2_scale_factor;
    for (int c = 0; c < 3; c++) {
        sei.fg_comp_model_present_flag[c] = fg->comp_model_present_flag[c];
        sei.fg_num_model_values_minus1[c] = fg->num_model_values_minus1[c];
        for (int i = 0; i < 8; i++) {
            sei.fg_intensity_interval_lower_bound[c][i] = fg->intensity_interval_lower_bound[c][i];
            sei.fg_intensity_interval_upper_bound[c][i] = fg->intensity_interval_upper_bound[c][i];
            for (int k = 0; k < 3; k++) sei.fg_comp_model_value[c][i][k] = fg->comp_model_value[c][i][k];
        }
    }
    int16_t *d[3] = { (int16_t *)dst[0], (int16_t *)dst[1], (int16_t *)dst[2] }, *s_[3] = { (int16_t *)src[0], (int16_t *)src[1], (int16_t *)src[2] };
    fg_grain_apply_pic(d, s_, &sei, w, h, poc, 0, 1);
}

/* SAO of a picture cut into tiles: one rect entry per tile, each driven like a picture of its own (slicedec.c runs entries this
 * way, one OVCTUDec per entry job; parameters are indexed relative to the entry, rcn_sao.c:217) */
void
ovref_sao_picture_tiles(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h,
                        int log2_ctu_s, int bitdepth, int simd, const OVOSaoCtu *p, const OVOTileGrid *tg)
{
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    for (int tr = 0; tr < tg->nb_rows; tr++)
        for (int tc = 0; tc < tg->nb_cols; tc++) {
            HarnessCtx *h = harness_open(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd);
            OVCTUDec *c = h->ctudec;
            struct RectEntryInfo e;
            memset(&e, 0, sizeof(e));
            e.ctb_x = tg->col_bd[tc]; e.ctb_y = tg->row_bd[tr];
            e.nb_ctu_w = tg->col_bd[tc + 1] - tg->col_bd[tc]; e.nb_ctu_h = tg->row_bd[tr + 1] - tg->row_bd[tr];
            e.nb_ctu_rect = e.nb_ctu_w * e.nb_ctu_h;
            e.nb_ctb_pic_w = nb_w;
            SAOParamsCtu *params = calloc(e.nb_ctu_rect, sizeof(*params));
            for (int y = 0; y < e.nb_ctu_h; y++)
                for (int x = 0; x < e.nb_ctu_w; x++) {
                    const OVOSaoCtu *q = &p[(e.ctb_y + y) * nb_w + e.ctb_x + x];
                    SAOParamsCtu *d = &params[y * e.nb_ctu_w + x];
                    for (int k = 0; k < 3; k++) {
                        memcpy(d->offset_val[k], q->offset_val[k], sizeof(d->offset_val[k]));
                        d->band_position[k] = q->band_position[k]; d->eo_class[k] = q->eo_class[k]; d->type_idx[k] = q->type_idx[k];
                    }
                }
            c->sao_info.sao_luma_flag = 1; c->sao_info.sao_chroma_flag = 1; c->sao_info.chroma_format_idc = 1;
            c->sao_info.sao_params = params;
            const struct SAOFunctions *sao = &c->rcn_funcs.sao;
            for (int row = 0; row < e.nb_ctu_h; row++) {
                if (row == 0) {
                    sao->rcn_sao_first_pix_rows(c, &e, 0);
                    if (e.nb_ctu_h == 1) sao->rcn_sao_filter_line(c, &e, 0);
                } else if (row == e.nb_ctu_h - 1) {
                    sao->rcn_sao_filter_line(c, &e, row - 1);
                    sao->rcn_sao_filter_line(c, &e, row);
                } else {
                    sao->rcn_sao_filter_line(c, &e, row - 1);
                }
            }
            harness_close(h);
            free(params);
        }
}

/* ALF of a picture cut into tiles: one rect entry per tile (parameters and CC-ALF indices relative to the entry, rcn_alf.c:1319-1321, 1399) */
void
ovref_alf_picture_tiles(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h,
                        int log2_ctu_s, int bitdepth, int simd,
                        const OVOAlfTables *tab, const OVOAlfCtuParams *p, const OVOTileGrid *tg)
{
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    /* in place like the decoder: an entry pads its own edges and never reads another entry's samples, so the order is free */
    for (int tr = 0; tr < tg->nb_rows; tr++)
        for (int tc = 0; tc < tg->nb_cols; tc++) {
            HarnessCtx *h = harness_open(planes, stride, pic_w, pic_h, log2_ctu_s, bitdepth, simd);
            OVCTUDec *c = h->ctudec;
            struct ALFInfo *ai = &c->alf_info;
            struct RectEntryInfo e;
            memset(&e, 0, sizeof(e));
            e.ctb_x = tg->col_bd[tc]; e.ctb_y = tg->row_bd[tr];
            e.nb_ctu_w = tg->col_bd[tc + 1] - tg->col_bd[tc]; e.nb_ctu_h = tg->row_bd[tr + 1] - tg->row_bd[tr];
            e.nb_ctu_rect = e.nb_ctu_w * e.nb_ctu_h;
            e.nb_ctb_pic_w = nb_w;
            ALFParamsCtu *params = calloc(e.nb_ctu_rect, sizeof(*params));
            uint8_t *ccb = calloc(e.nb_ctu_rect, 1), *ccr = calloc(e.nb_ctu_rect, 1);
            OVALFData cc[2];
            memset(cc, 0, sizeof(cc));
            memcpy(ai->rcn_alf.filter_coeff_dec, tab->luma_coeff, sizeof(tab->luma_coeff));
            memcpy(ai->rcn_alf.filter_clip_dec,  tab->luma_clip,  sizeof(tab->luma_clip));
            memcpy(ai->rcn_alf.chroma_coeff_final, tab->chroma_coeff, sizeof(tab->chroma_coeff));
            memcpy(ai->rcn_alf.chroma_clip_final,  tab->chroma_clip,  sizeof(tab->chroma_clip));
            memcpy(cc[0].alf_cc_mapped_coeff[0], tab->cc_coeff[0], sizeof(tab->cc_coeff[0]));
            memcpy(cc[1].alf_cc_mapped_coeff[1], tab->cc_coeff[1], sizeof(tab->cc_coeff[1]));
            for (int y = 0; y < e.nb_ctu_h; y++)
                for (int x = 0; x < e.nb_ctu_w; x++) {
                    const int i = (e.ctb_y + y) * nb_w + e.ctb_x + x, k = y * e.nb_ctu_w + x;
                    params[k].ctb_alf_flag = p->ctb_alf_flag[i]; params[k].ctb_alf_idx = p->ctb_alf_idx[i];
                    params[k].cb_alternative = p->cb_alternative[i]; params[k].cr_alternative = p->cr_alternative[i];
                    ccb[k] = p->cc_idx_cb[i]; ccr[k] = p->cc_idx_cr[i];
                }
            ai->alf_luma_enabled_flag = 1; ai->alf_cb_enabled_flag = 1; ai->alf_cr_enabled_flag = 1;
            ai->cc_alf_cb_enabled_flag = p->cc_cb_enabled; ai->cc_alf_cr_enabled_flag = p->cc_cr_enabled;
            ai->aps_cc_alf_data_cb = &cc[0]; ai->aps_cc_alf_data_cr = &cc[1];
            ai->ctb_alf_params = params;
            ai->ctb_cc_alf_filter_idx[0] = ccb; ai->ctb_cc_alf_filter_idx[1] = ccr;
            for (int row = 0; row < e.nb_ctu_h; row++)
                c->rcn_funcs.alf.rcn_alf_filter_line(c, &e, row);
            harness_close(h);
            free(params); free(ccb); free(ccr);
        }
}
