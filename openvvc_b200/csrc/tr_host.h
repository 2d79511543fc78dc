// tr_host.h -- host-side preparation of the residual work lists, shared by tr.cu and tests/emu.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include "../../include/ovvc_b200.h"
#include "tr_kernels.h"

// Validates one command against the coefficient arena; returns 0 when well-formed.
static inline int tu_check_cmd(const OVGPUTuCmd *c, size_t nb_coeffs)
{
    if (c->log2_w < 2 || c->log2_w > 6 || c->log2_h < 2 || c->log2_h > 6 || c->x0 < 0 || c->y0 < 0) return -1;
    if (c->cu_mts_idx > 3 || c->lfnst_idx > 1 || c->lfnst_mode_l > 66 || c->lfnst_mode_c > 66) return -1;
    if (c->cu_mts_flag && (c->log2_w > 5 || c->log2_h > 5)) return -1;
    for (int k = 0; k < 3; k++) {
        int used = k == 2 ? (c->cbf_mask & 0x10) : (k == 0 ? (c->cbf_mask & 0x0a) : ((c->cbf_mask & 0x01) && !(c->cbf_mask & 0x08)));
        if (!used) continue;
        int lw = c->log2_w - (k != 2), lh = c->log2_h - (k != 2);
        int ts = k == 2 ? (c->tr_skip_mask & 0x10) : ((c->cbf_mask & 0x08) ? (c->tr_skip_mask & 0x1) : (c->tr_skip_mask & (k == 0 ? 0x2 : 0x1)));
        if (ts && (lw > 5 || lh > 5)) return -1;
        size_t n = ts ? (size_t)1 << (lw + lh) : (size_t)(lw < 5 ? 1 << lw : 32) * (lh < 5 ? 1 << lh : 32);
        if ((size_t)c->coeff_off[k] + n > nb_coeffs) return -1;
    }
    for (int k = 0; k < 4; k++) if (c->qp[k] > 75 || c->qp_ts[k] > 75) return -1;
    return 0;
}

// Splits the TUs into transform blocks: `small` (<= 256 samples, one warp each) and `large` (one CTA each).
// Either output may be NULL to only count. Returns the totals through nb_small / nb_large.
static inline void tu_build_items(TbItemDev *small, TbItemDev *large, int *nb_small, int *nb_large,
                                  const OVGPUTuCmd *cmds, int n)
{
    int ns = 0, nl = 0;
    for (int i = 0; i < n; i++) {
        const OVGPUTuCmd *c = &cmds[i];
        int kinds[3], nk = 0;
        if (c->cbf_mask & 0x10) kinds[nk++] = 0;
        if (c->cbf_mask & 0x08) kinds[nk++] = 3;
        else {
            if (c->cbf_mask & 0x02) kinds[nk++] = 1;
            if (c->cbf_mask & 0x01) kinds[nk++] = 2;
        }
        for (int k = 0; k < nk; k++) {
            int s = c->log2_w + c->log2_h - (kinds[k] ? 2 : 0);
            TbItemDev it = { (uint32_t)i, (uint32_t)kinds[k] };
            if (s <= 8) { if (small) small[ns] = it; ns++; }
            else        { if (large) large[nl] = it; nl++; }
        }
    }
    *nb_small = ns; *nb_large = nl;
}

// ---- reconstruction order ----------------------------------------------------------------------------------------
// The reference reconstructs CU by CU: prediction, then rcn_transform_tree (vcl_coding_unit.c -> rcn_res_wrap), so an
// intra block reads neighbours that already hold prediction + residual.  The picture-level schedule (intra_host.h:
// recon_for_each_launch) keeps that order per plane by dependency LEVEL: transform blocks covered by an intra command run
// right after that command's level; all others (inter CUs) run first.  `map_l` / `map_c` hold, per 4x4 luma unit, the
// 1-based level of the intra command covering it (0 = none) for the luma / chroma plane.  Groups: luma blocks 0 .. nl_l,
// chroma blocks nl_l + 1 .. nl_l + 1 + nl_c.
struct TuGroups {
    int nb_groups;
    int *start_small, *n_small, *start_large, *n_large;   // per group, offsets into the re-ordered item array
};

static inline int tu_group_of(const OVGPUTuCmd *c, int kind, const uint16_t *map_l, const uint16_t *map_c, int w_u, int nl_l)
{
    const uint16_t *m = kind == 0 ? map_l : map_c;
    const int lvl = m ? m[(size_t)(c->y0 >> 2) * w_u + (c->x0 >> 2)] : 0;
    return kind == 0 ? lvl : nl_l + 1 + lvl;
}

static inline void tu_groups_free(TuGroups *g)
{
    free(g->start_small); g->start_small = g->n_small = g->start_large = g->n_large = NULL; g->nb_groups = 0;
}

// items: room for every transform block of the batch.  Returns the number of items or -1 (out of memory).
static inline int tu_build_grouped_items(TbItemDev *items, TuGroups *g, const OVGPUTuCmd *cmds, int n,
                                         const uint16_t *map_l, const uint16_t *map_c, int w_u, int nl_l, int nl_c)
{
    tu_groups_free(g);
    const int ng = nl_l + nl_c + 2;
    int *mem = (int *)calloc((size_t)4 * ng, sizeof(int));
    if (!mem) return -1;
    g->nb_groups = ng;
    g->start_small = mem; g->n_small = mem + ng; g->start_large = mem + 2 * ng; g->n_large = mem + 3 * ng;
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < n; i++) {
            const OVGPUTuCmd *c = &cmds[i];
            int kinds[3], nk = 0;
            if (c->cbf_mask & 0x10) kinds[nk++] = 0;
            if (c->cbf_mask & 0x08) kinds[nk++] = 3;
            else {
                if (c->cbf_mask & 0x02) kinds[nk++] = 1;
                if (c->cbf_mask & 0x01) kinds[nk++] = 2;
            }
            for (int k = 0; k < nk; k++) {
                const int s = c->log2_w + c->log2_h - (kinds[k] ? 2 : 0);
                const int grp = tu_group_of(c, kinds[k], map_l, map_c, w_u, nl_l);
                if (pass == 0) { if (s <= 8) g->n_small[grp]++; else g->n_large[grp]++; }
                else {
                    TbItemDev it = { (uint32_t)i, (uint32_t)kinds[k] };
                    if (s <= 8) items[g->start_small[grp] + g->n_small[grp]++] = it;
                    else        items[g->start_large[grp] + g->n_large[grp]++] = it;
                }
            }
        }
        if (pass == 0) {
            int acc = 0;
            for (int q = 0; q < ng; q++) {
                g->start_small[q] = acc; acc += g->n_small[q]; g->n_small[q] = 0;
                g->start_large[q] = acc; acc += g->n_large[q]; g->n_large[q] = 0;
            }
        }
    }
    return ng ? g->start_large[ng - 1] + g->n_large[ng - 1] : 0;
}
