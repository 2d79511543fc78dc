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

#include "oracle.h"

void rcn_init_intra_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_dc_planar_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_intra_angular_functions_10(struct RCNFunctions *rcn_func);
void rcn_init_fill_ref_10(struct RCNFunctions *rcn_func);

void
ovref_intra_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth, int simd,
                    OVOIntraCmd *cmds, int nb_cmds)
{
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
    rcn_init_functions(&c->rcn_funcs, 0, 1, 0, 0, bitdepth);
    if (!simd && bitdepth == 10) {      /* scalar C templates = arithmetic of record (rcn.c:98-117, 151-178) */
        extern const struct IntraAngularFunctions angular_gauss_h_10, angular_gauss_v_10, angular_cubic_h_10, angular_cubic_v_10,
                                                  angular_c_h_10, angular_c_v_10, angular_nofrac_v_10, angular_nofrac_h_10;
        rcn_init_intra_functions_10(&c->rcn_funcs);
        rcn_init_dc_planar_functions_10(&c->rcn_funcs);
        rcn_init_fill_ref_10(&c->rcn_funcs);
        c->rcn_funcs.intra_angular_gauss_h = &angular_gauss_h_10; c->rcn_funcs.intra_angular_gauss_v = &angular_gauss_v_10;
        c->rcn_funcs.intra_angular_cubic_h = &angular_cubic_h_10; c->rcn_funcs.intra_angular_cubic_v = &angular_cubic_v_10;
        c->rcn_funcs.intra_angular_c_h = &angular_c_h_10; c->rcn_funcs.intra_angular_c_v = &angular_c_v_10;
        c->rcn_funcs.intra_angular_nofrac_v = &angular_nofrac_v_10; c->rcn_funcs.intra_angular_nofrac_h = &angular_nofrac_h_10;
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
                if (m->mode != OVO_INTRA_NO_LUMA)
                    c->rcn_funcs.intra_pred(&c->rcn_ctx, &c->rcn_ctx.ctu_buff, m->mode, x0, y0, m->log2_w, m->log2_h, flg_pred_mode_flag);
                if (m->flags & OVO_INTRA_CHROMA) {
                    /* single tree: the chroma field is updated with the same rectangles as the luma one (rcn_tu_st, :1275) */
                    if (c->rcn_ctx.progress_field_c.vfield[x0 >> 2] != m->avail_cols || c->rcn_ctx.progress_field_c.hfield[y0 >> 2] != m->avail_rows) abort();
                    c->rcn_funcs.intra_pred_c(&c->rcn_ctx, m->flags & 0x7f, x0 >> 1, y0 >> 1, m->log2_w - 1, m->log2_h - 1, flg_pred_mode_flag);
                }
                ctu_field_set_rect_bitfield(&c->rcn_ctx.progress_field, x0 >> 2, y0 >> 2, (1 << m->log2_w) >> 2, (1 << m->log2_h) >> 2);
                ctu_field_set_rect_bitfield(&c->rcn_ctx.progress_field_c, x0 >> 2, y0 >> 2, (1 << m->log2_w) >> 2, (1 << m->log2_h) >> 2);
            }
        }
    free(c);
}
