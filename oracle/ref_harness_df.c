/* oracle/ref_harness_df.c -- TEST INFRASTRUCTURE ONLY (see ref_harness.c).
 *
 * Drives the reference's deblocking filter the way slicedec.c does (decode_ctu :753-758, decode_ctu_line :872-877):
 * for every CTU in raster order the DBFInfo of the ctudec is filled from a block list with the reference's OWN
 * helpers (fill_ctb_bound / fill_bs_map / dbf_fill_qp_map / dbf_fill_aff_map, dbf_utils.h), df.rcn_dbf_ctu (or
 * rcn_dbf_truncated_ctu) filters the CTU in place in the frame, and dbf_store_info / dbf_load_info (drv_lines.c:754-
 * 771) carry the state to the next CTU.  At each rcn_dbf_ctu call the DBFInfo is also converted to the picture-level
 * OVODbfMaps layout -- exactly what the integration shim's recorder does -- so tests can check our
 * map generator against the reference's bookkeeping.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "ovdefs.h"
#include "ovframe.h"
#include "ctudec.h"
#include "dec_structures.h"
#include "slicedec.h"
#include "rcn.h"
#include "rcn_structures.h"
#include "dbf_utils.h"

#include "oracle.h"

void dbf_load_info(struct DBFInfo *const dbf_info, const struct DBFLines *const dbf_lines, uint8_t log2_ctu_s, int ctb_x);
void dbf_store_info(struct DBFInfo *const dbf_info, const struct DBFLines *const dbf_lines, uint8_t log2_ctu_s, int ctb_x);
void rcn_init_df_functions_10(struct RCNFunctions *rcn_func);

static void
record_ctu(OVODbfMaps *m, const struct DBFInfo *d, int cx, int cy, int n_u, int nuw, int nuh)
{
    int w_u = m->width >> 2, h_u = m->height >> 2;
    uint64_t vmask = nuh >= 64 ? ~0ull : (1ull << nuh) - 1, hmask = nuw >= 64 ? ~0ull : (1ull << nuw) - 1;
    const uint64_t *vsrc[OVO_DBF_NB_MAPS] = { d->ctb_bound_ver + 8, d->aff_edg_ver + 8, d->bs1_map.ver, d->bs2_map.ver, d->affine_map.ver + 1,
                                               d->ctb_bound_ver_c + 8, d->bs1_map_cb.ver, d->bs1_map_cr.ver, d->bs2_map_c.ver };
    const uint64_t *hsrc[OVO_DBF_NB_MAPS] = { d->ctb_bound_hor + 8, d->aff_edg_hor + 8, d->bs1_map.hor, d->bs2_map.hor, d->affine_map.hor + 1,
                                               d->ctb_bound_hor_c + 8, d->bs1_map_cb.hor, d->bs1_map_cr.hor, d->bs2_map_c.hor };
    for (int k = 0; k < OVO_DBF_NB_MAPS; k++) {
        /* column nuw / row nuh (the far CTU border) belongs to the next CTU's record, except at the picture border */
        int last_c = (cx * n_u + nuw == w_u) ? nuw + 1 : nuw, last_r = (cy * n_u + nuh == h_u) ? nuh + 1 : nuh;
        if (k == OVO_DBF_AFFINE) { last_c = nuw; last_r = nuh; }
        for (int i = 0; i < last_c; i++) m->ver[k][cy * (w_u + 1) + cx * n_u + i] = (uint32_t)(vsrc[k][i] & vmask);
        for (int i = 0; i < last_r; i++) m->hor[k][cx * (h_u + 1) + cy * n_u + i] = (uint32_t)((hsrc[k][i] >> 2) & hmask);
    }
    for (int y = 0; y < nuh; y++)
        for (int x = 0; x < nuw; x++) {
            int pos = 34 * (1 + y) + 2 + x;
            m->qp[0][(cy * n_u + y) * w_u + cx * n_u + x] = d->qp_map_y.hor[pos];
            m->qp[1][(cy * n_u + y) * w_u + cx * n_u + x] = d->qp_map_cb.hor[pos];
            m->qp[2][(cy * n_u + y) * w_u + cx * n_u + x] = d->qp_map_cr.hor[pos];
        }
}

void
ovref_dbf_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth, int simd,
                  const OVODbfBlk *blks, int nb_blks, int beta_offset, int tc_offset, OVODbfMaps *rec)
{
    OVCTUDec *c;
    OVPartInfo part;
    int ctu_s = 1 << log2_ctu_s, n_u = ctu_s >> 2;
    int ncw = (pic_w + ctu_s - 1) >> log2_ctu_s, nch = (pic_h + ctu_s - 1) >> log2_ctu_s;
    struct DBFLines l;
    ctudec_init(&c);
    memset(&part, 0, sizeof(part));
    part.log2_ctu_s = log2_ctu_s; part.log2_min_cb_s = 2;
    c->part_ctx = &part;
    c->rcn_ctx.ctudec = c;
    c->pic_w = pic_w; c->pic_h = pic_h;
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 0, bitdepth);
    if (!simd && bitdepth == 10) rcn_init_df_functions_10(&c->rcn_funcs);
    c->tmp_slice_type = 2;      /* skip the MV pre-processing: MV/ref-mismatch edges arrive as bs1 bits (flag 16) */
    /* DBFLines (slicedec.h:127-148), zeroed like dbf_clear_lines (drv_lines.c:560-579) */
    l.qp_x_map = calloc(ncw * 32 + 64, 1); l.qp_x_map_cb = calloc(ncw * 32 + 64, 1); l.qp_x_map_cr = calloc(ncw * 32 + 64, 1);
    l.small_map = calloc(ncw + 2, 8); l.dbf_bs2_hor = calloc(ncw + 2, 8); l.dbf_bs2_hor_c = calloc(ncw + 2, 8);
    l.dbf_bs1_hor = calloc(ncw + 2, 8); l.dbf_bs1_hor_cb = calloc(ncw + 2, 8); l.dbf_bs1_hor_cr = calloc(ncw + 2, 8);
    l.dbf_affine = calloc(ncw + 2, 8); l.large_map_c = calloc(ncw + 2, 8);
    struct DBFInfo *d = &c->dbf_info;
    d->beta_offset = beta_offset; d->tc_offset = tc_offset;

    /* blocks are bucketed per CTU first (the generator emits them CTU by CTU, but do not rely on it) */
    int *first = calloc(ncw * nch + 1, sizeof(int)), *order = malloc(sizeof(int) * (nb_blks + 1));
    for (int i = 0; i < nb_blks; i++) first[(blks[i].y0 >> log2_ctu_s) * ncw + (blks[i].x0 >> log2_ctu_s) + 1]++;
    for (int i = 0; i < ncw * nch; i++) first[i + 1] += first[i];
    int *fill = calloc(ncw * nch, sizeof(int));
    for (int i = 0; i < nb_blks; i++) { int k = (blks[i].y0 >> log2_ctu_s) * ncw + (blks[i].x0 >> log2_ctu_s); order[first[k] + fill[k]++] = i; }

    for (int cy = 0; cy < nch; cy++) {
        dbf_load_info(d, &l, log2_ctu_s, 0);                                     /* drv_line_next_line, drv_lines.c:1025 */
        for (int cx = 0; cx < ncw; cx++) {
            int ctu_w = pic_w - (cx << log2_ctu_s) < ctu_s ? pic_w - (cx << log2_ctu_s) : ctu_s;
            int ctu_h = pic_h - (cy << log2_ctu_s) < ctu_s ? pic_h - (cy << log2_ctu_s) : ctu_s;
            for (int k = first[cy * ncw + cx]; k < first[cy * ncw + cx + 1]; k++) {
                const OVODbfBlk *b = &blks[order[k]];
                int x0 = b->x0 & (ctu_s - 1), y0 = b->y0 & (ctu_s - 1), lw = b->log2_w, lh = b->log2_h;
                fill_ctb_bound(d, x0, y0, lw, lh);                              /* rcn_tu_st, rcn_transform_tree.c:1299-1300 */
                fill_ctb_bound_c(d, x0, y0, lw, lh);
                if (b->flags & 1) { fill_bs_map(&d->bs2_map, x0, y0, lw, lh); fill_bs_map(&d->bs2_map_c, x0, y0, lw, lh); }
                if (b->flags & (2 | 16)) fill_bs_map(&d->bs1_map, x0, y0, lw, lh);
                if (b->flags & 4) fill_bs_map(&d->bs1_map_cb, x0, y0, lw, lh);
                if (b->flags & 8) fill_bs_map(&d->bs1_map_cr, x0, y0, lw, lh);
                if (b->flags & 32) {
                    int x0u = x0 >> 2, y0u = y0 >> 2, nw = (1 << lw) >> 2, nh = (1 << lh) >> 2;
                    uint64_t mv = ((1ull << nh) - 1) << y0u, mh = ((1ull << nw) - 1) << (2 + x0u);
                    dbf_fill_aff_map(&d->affine_map, x0u, y0u, nw, nh);         /* drv_affine_mvp.c:3082 */
                    for (int i = 2; i < nw; i += 2) d->aff_edg_ver[8 + x0u + i] |= mv;   /* what dbf_set_sb_edges (:3051-3070) records */
                    for (int i = 2; i < nh; i += 2) d->aff_edg_hor[8 + y0u + i] |= mh;
                }
                dbf_fill_qp_map(&d->qp_map_y, x0, y0, lw, lh, b->qp_y);
                dbf_fill_qp_map(&d->qp_map_cb, x0, y0, lw, lh, b->qp_cb);
                dbf_fill_qp_map(&d->qp_map_cr, x0, y0, lw, lh, b->qp_cr);
            }
            c->ctu_ngh_flags = (cx ? CTU_LFT_FLG : 0) | (cy ? CTU_UP_FLG : 0);
            c->rcn_ctx.frame_buff.y = planes[0] + (size_t)(cy << log2_ctu_s) * stride[0] + (cx << log2_ctu_s);
            c->rcn_ctx.frame_buff.cb = planes[1] + (size_t)(cy << (log2_ctu_s - 1)) * stride[1] + (cx << (log2_ctu_s - 1));
            c->rcn_ctx.frame_buff.cr = planes[2] + (size_t)(cy << (log2_ctu_s - 1)) * stride[2] + (cx << (log2_ctu_s - 1));
            c->rcn_ctx.frame_buff.stride = stride[0]; c->rcn_ctx.frame_buff.stride_c = stride[1];
            if (rec) record_ctu(rec, d, cx, cy, n_u, ctu_w >> 2, ctu_h >> 2);
            if (ctu_w == ctu_s && ctu_h == ctu_s)
                c->rcn_funcs.df.rcn_dbf_ctu(&c->rcn_ctx, d, log2_ctu_s, cx == ncw - 1, cy == nch - 1);
            else
                c->rcn_funcs.df.rcn_dbf_truncated_ctu(&c->rcn_ctx, d, log2_ctu_s, cx == ncw - 1, cy == nch - 1, ctu_w, ctu_h);
            dbf_store_info(d, &l, log2_ctu_s, cx);
            if (cx + 1 < ncw) dbf_load_info(d, &l, log2_ctu_s, cx + 1);
        }
    }
    free(first); free(order); free(fill);
    free(l.qp_x_map); free(l.qp_x_map_cb); free(l.qp_x_map_cr); free(l.small_map); free(l.dbf_bs2_hor); free(l.dbf_bs2_hor_c);
    free(l.dbf_bs1_hor); free(l.dbf_bs1_hor_cb); free(l.dbf_bs1_hor_cr); free(l.dbf_affine); free(l.large_map_c);
    free(c);
}
