// tr_host.h -- host-side preparation of the residual work lists, shared by tr.cu and tests/emu.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include "../../include/ovvc_b200.h"
#include "recon_kernels.h"

// Validates one command against the coefficient arena; returns 0 when well-formed.
static inline int tu_check_cmd(const OVGPUTuCmd *c, size_t nb_coeffs)
{
    if (c->flags & OVGPU_TU_ISP) {          /* luma block of an ISP partition: 1xN .. 64x16, at least 16 samples, luma only */
        if (c->log2_w > 6 || c->log2_h > 6 || c->log2_w + c->log2_h < 4 || c->x0 < 0 || c->y0 < 0 || c->cbf_mask != 0x10 || c->tr_skip_mask ||
            c->cu_mts_flag || c->bdpcm || c->cu_mts_idx > 3 || c->lfnst_idx > 1 || c->lfnst_mode_l > 127 || (c->flags & (OVGPU_TU_IMPLICIT_MTS | OVGPU_TU_LFNST_C)) ||
            c->qp[0] > 75)
            return -1;
        const int thin = c->log2_w < 2 || c->log2_h < 2;
        const size_t n = thin ? (size_t)1 << (c->log2_w + c->log2_h) : (size_t)(c->log2_w < 5 ? 1 << c->log2_w : 32) * (c->log2_h < 5 ? 1 << c->log2_h : 32);
        return (size_t)c->coeff_off[2] + n > nb_coeffs ? -1 : 0;
    }
    if (c->log2_w < 2 || c->log2_w > 6 || c->log2_h < 2 || c->log2_h > 6 || c->x0 < 0 || c->y0 < 0) return -1;
    if (c->cu_mts_idx > 3 || c->lfnst_idx > 1 || c->lfnst_mode_l > 66 || c->lfnst_mode_c > 66 || c->bdpcm > 15) return -1;
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
// Both lists are ordered by DECREASING block size (counting sort on log2 of the sample count; transform blocks write disjoint
// samples, so any order gives the same picture): the warps of a CTA, which meet at the barriers between the phases, then
// hold blocks of one size instead of waiting for the largest, and the last wave of a launch is made of the cheapest blocks.
static inline int tu_kinds(const OVGPUTuCmd *c, int kinds[3])
{
    int nk = 0;
    if (c->cbf_mask & 0x10) kinds[nk++] = 0;
    if (c->cbf_mask & 0x08) kinds[nk++] = 3;
    else {
        if (c->cbf_mask & 0x02) kinds[nk++] = 1;
        if (c->cbf_mask & 0x01) kinds[nk++] = 2;
    }
    return nk;
}
// cursors of a size-bucketed list: keys (log2 samples) hi .. lo in decreasing order, starting at `base`; returns the end
static inline int tu_bucket_starts(const int cnt[16], int cur[16], int hi, int lo, int base)
{
    for (int k = hi; k >= lo; k--) { cur[k] = base; base += cnt[k]; }
    return base;
}
static inline void tu_build_items(TbItemDev *small, TbItemDev *large, int *nb_small, int *nb_large,
                                  const OVGPUTuCmd *cmds, int n)
{
    int cnt[16] = { 0 }, cur[16], kinds[3];
    for (int i = 0; i < n; i++) {
        const int nk = tu_kinds(&cmds[i], kinds);
        for (int k = 0; k < nk; k++) cnt[(cmds[i].log2_w + cmds[i].log2_h - (kinds[k] ? 2 : 0)) & 15]++;
    }
    const int ns = tu_bucket_starts(cnt, cur, 8, 0, 0);
    const int nl = tu_bucket_starts(cnt, cur, 15, 9, 0);
    *nb_small = ns; *nb_large = nl;
    if (!small && !large) return;
    for (int i = 0; i < n; i++) {
        const int nk = tu_kinds(&cmds[i], kinds);
        for (int k = 0; k < nk; k++) {
            const int s = (cmds[i].log2_w + cmds[i].log2_h - (kinds[k] ? 2 : 0)) & 15;
            TbItemDev it = { (uint32_t)i, (uint32_t)kinds[k] };
            if (s <= 8) { if (small) small[cur[s]] = it; }
            else        { if (large) large[cur[s]] = it; }
            cur[s]++;
        }
    }
}

// ---- reconstruction order ----------------------------------------------------------------------------------------
// The reference reconstructs CU by CU: prediction, then rcn_transform_tree (vcl_coding_unit.c -> rcn_res_wrap), so an
// intra block reads neighbours that already hold prediction + residual.  ovgpu_recon_picture keeps that order per plane:
// the transform blocks that lie outside intra commands (inter CUs, whose prediction motion compensation wrote before) run
// first as ordinary launches; the blocks covered by an intra command are attached to that command and run inside the
// dataflow kernel right after its prediction (recon_kernels.h).  `pos_l` / `pos_c` hold, per 4x4 luma unit, 1 + the queue
// position of the intra command covering it in the luma / chroma plane (0 = none).
struct TuReconLists {
    int built;
    int start_small[2], n_small[2], start_large[2], n_large[2];    // [plane] blocks outside intra commands
    int nb_tb;                                                     // all transform blocks of the batch
};

// items: room for every transform block of the batch; ritems[p]: one entry per intra command of plane p in queue order
// (cmd already filled by the caller), receives the block lists.  Returns the number of blocks.
static inline int tu_build_recon_lists(TbItemDev *items, TuReconLists *g, ReconItemDev *const ritems[2], const int nb_ritems[2],
                                       const OVGPUTuCmd *cmds, int n, const uint32_t *pos_l, const uint32_t *pos_c, int w_u)
{
    memset(g, 0, sizeof(*g));
    for (int p = 0; p < 2; p++) for (int q = 0; q < nb_ritems[p]; q++) { ritems[p][q].tb_start = 0; ritems[p][q].nb_tb = 0; ritems[p][q].large = 0; }
    int cnt[2][16], cur[2][16];                         // blocks outside intra commands, per plane and log2 size (see tu_build_items)
    memset(cnt, 0, sizeof(cnt));
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < n; i++) {
            const OVGPUTuCmd *c = &cmds[i];
            int kinds[3];
            const int nk = tu_kinds(c, kinds);
            for (int k = 0; k < nk; k++) {
                const int p = kinds[k] != 0, s = (c->log2_w + c->log2_h - (p ? 2 : 0)) & 15, large = s > 8;
                const uint32_t *m = p ? pos_c : pos_l;
                const uint32_t pos = m ? m[(size_t)(c->y0 >> 2) * w_u + (c->x0 >> 2)] : 0;
                TbItemDev it = { (uint32_t)i, (uint32_t)kinds[k] };
                if (!pos) {
                    if (pass == 0) cnt[p][s]++;
                    else items[cur[p][s]++] = it;
                } else {
                    ReconItemDev *r = &ritems[p][pos - 1];
                    if (pass == 0) { r->nb_tb++; r->large = (uint16_t)large; }
                    else items[r->tb_start + r->nb_tb++] = it;
                }
            }
        }
        if (pass == 0) {
            int acc = 0;
            for (int p = 0; p < 2; p++) {
                g->start_small[p] = acc; acc = tu_bucket_starts(cnt[p], cur[p], 8, 0, acc);  g->n_small[p] = acc - g->start_small[p];
                g->start_large[p] = acc; acc = tu_bucket_starts(cnt[p], cur[p], 15, 9, acc); g->n_large[p] = acc - g->start_large[p];
            }
            for (int p = 0; p < 2; p++)
                for (int q = 0; q < nb_ritems[p]; q++) { ritems[p][q].tb_start = (uint32_t)acc; acc += ritems[p][q].nb_tb; ritems[p][q].nb_tb = 0; }
            g->nb_tb = acc;
        }
    }
    g->built = 1;
    return g->nb_tb;
}
