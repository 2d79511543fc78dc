// tr_host.h -- host-side preparation of the residual work lists, shared by tr.cu and tests/emu.
#pragma once
#include <stdint.h>
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
