// mc_host.h -- host-side preparation of the MC work list, shared by mc.cu and tests/emu.
#pragma once
#include <stdint.h>
#include "../../include/ovvc_b200.h"
#include "mc_tools_kernels.h"
#include "mc_big_kernels.h"

#define MC_ROUTE_BIG 6

// Which kernel runs a command: plain 4x4 luma-only commands (affine sub-blocks without PROF) take the packed 4x4 path of
// ProfK (four sub-blocks per warp, refinement off) instead of a whole McK warp each.
static inline int mc_route(const OVGPUMcCmd *c)
{
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w == 2 && c->log2_h == 2 && (c->flags & 3) == OVGPU_MC_LUMA) return OVGPU_MC_TOOL_PROF;
    // plain chroma-only commands of 8x8 luma samples (the chroma blocks of affine CUs): two per warp in McSmallCK
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w == 3 && c->log2_h == 3 && (c->flags & 3) == OVGPU_MC_CHROMA && !(c->x0 & 3)) return MC_ROUTE_SMALL_C;
    // plain commands of at least 32 x 16 luma samples: 32 x 32 / 32 x 16 tiles in McBigK
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w >= 5 && c->log2_h >= 4) return MC_ROUTE_BIG;
    return c->tool == OVGPU_MC_TOOL_GPM ? OVGPU_MC_TOOL_PLAIN : c->tool;
}

// number of <=16x16 work items of the plain (tool 0) commands of a PU list
static inline int mc_count_items(const OVGPUMcCmd *cmds, int n)
{
    int total = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i]) != OVGPU_MC_TOOL_PLAIN) continue;
        int w = 1 << cmds[i].log2_w, h = 1 << cmds[i].log2_h;
        total += ((w + 15) >> 4) * ((h + 15) >> 4);
    }
    return total;
}

// BDOF / PROF commands are one work item each (BdofK: one warp per item, ProfK: four items per warp)
static inline int mc_count_tool_items(const OVGPUMcCmd *cmds, int n, int tool)
{
    int total = 0;
    for (int i = 0; i < n; i++) total += mc_route(&cmds[i]) == tool;
    return total;
}

static inline int mc_build_tool_items(McItemDev *items, const OVGPUMcCmd *cmds, int n, int tool)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i]) != tool) continue;
        items[k].cmd = (uint32_t)i; items[k].ox = items[k].oy = 0;
        items[k].w = (uint8_t)(1 << cmds[i].log2_w); items[k].h = (uint8_t)(1 << cmds[i].log2_h);
        k++;
    }
    return k;
}

// well-formedness of one command; nb_prof = entries of the PROF table
static inline int mc_check_cmd(const OVGPUMcCmd *c, int nb_prof)
{
    if (c->log2_w < 2 || c->log2_w > 7 || c->log2_h < 2 || c->log2_h > 7 || c->inter_dir < 1 || c->inter_dir > 3 ||
        c->ref_idx0 < 0 || c->ref_idx0 > 15 || c->ref_idx1 < 0 || c->ref_idx1 > 15 || c->bcw_idx_plus1 > 5 || c->x0 < 0 || c->y0 < 0)
        return -1;
    if (c->tool == OVGPU_MC_TOOL_PLAIN) return 0;
    if (c->tool == OVGPU_MC_TOOL_GPM)       /* rcn_gpm_b: CU of 8..64 samples a side, two uni-predictions, luma and chroma, default BCW */
        return (c->inter_dir == 3 && c->log2_w >= 3 && c->log2_w <= 6 && c->log2_h >= 3 && c->log2_h <= 6 && (c->flags & 3) == 3 &&
                (c->bcw_idx_plus1 == 0 || c->bcw_idx_plus1 == 3) && (c->aux & ~0x33fu) == 0) ? 0 : -1;
    if (c->tool == OVGPU_MC_TOOL_BDOF)      /* rcn_bdof_mcp_l: bi-predicted luma sub-block of 8 or 16 samples a side */
        return (c->inter_dir == 3 && c->log2_w >= 3 && c->log2_w <= 4 && c->log2_h >= 3 && c->log2_h <= 4 &&
                (c->flags & 3) == OVGPU_MC_LUMA) ? 0 : -1;
    if (c->tool == OVGPU_MC_TOOL_PROF)      /* rcn_prof_mcp_b_l: 4x4 luma sub-block, table entry must exist */
        return (c->log2_w == 2 && c->log2_h == 2 && (c->flags & 3) == OVGPU_MC_LUMA && (int)(c->aux >> 8) < nb_prof &&
                (c->inter_dir == 3 || (c->aux & 3) == c->inter_dir)) ? 0 : -1;
    if (c->tool == OVGPU_MC_TOOL_DMVR)      /* rcn_dmvr_mv_refine: bi-predicted 8/16-sample sub-block, luma and chroma, no BCW */
        return (c->inter_dir == 3 && c->log2_w >= 3 && c->log2_w <= 4 && c->log2_h >= 3 && c->log2_h <= 4 &&
                (c->flags & 3) == 3 && (c->bcw_idx_plus1 == 0 || c->bcw_idx_plus1 == 3)) ? 0 : -1;
    return -1;
}

// 32 x 32 (32 x 16 for 16-row commands) tiles of the commands routed to McBigK
static inline int mc_build_big_items(McItemDev *items, const OVGPUMcCmd *cmds, int n)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i]) != MC_ROUTE_BIG) continue;
        const int w = 1 << cmds[i].log2_w, h = 1 << cmds[i].log2_h, th = h >= 32 ? 32 : 16;
        for (int oy = 0; oy < h; oy += th)
            for (int ox = 0; ox < w; ox += 32) {
                if (items) { items[k].cmd = (uint32_t)i; items[k].ox = (uint8_t)ox; items[k].oy = (uint8_t)oy; items[k].w = 32; items[k].h = (uint8_t)th; }
                k++;
            }
    }
    return k;
}

static inline int mc_build_items(McItemDev *items, const OVGPUMcCmd *cmds, int n)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i]) != OVGPU_MC_TOOL_PLAIN) continue;
        int w = 1 << cmds[i].log2_w, h = 1 << cmds[i].log2_h;
        for (int oy = 0; oy < h; oy += 16)
            for (int ox = 0; ox < w; ox += 16) {
                items[k].cmd = (uint32_t)i;
                items[k].ox = (uint8_t)ox; items[k].oy = (uint8_t)oy;
                items[k].w = (uint8_t)(w - ox < 16 ? w - ox : 16);
                items[k].h = (uint8_t)(h - oy < 16 ? h - oy : 16);
                k++;
            }
    }
    return k;
}

// LMCS LUT construction (rcn_lmcs.c:84-203), host side
static inline void lmcs_build_luts_host(OVGPULmcsLuts *luts, const OVGPULmcsData *d, int bitdepth)
{
    const int nb_wnd = 16, prec = 11, rnd = 1 << (prec - 1);
    const int rng = 1 << bitdepth, log2_wnd = bitdepth - 4, smp_wnd = rng >> 4, maxv = rng - 1;
    const int min_idx = d->lmcs_min_bin_idx, max_idx_p1 = nb_wnd - d->lmcs_delta_max_bin_idx;
    uint16_t fwd_step[16] = { 0 }, bwd_step[16] = { 0 }, bnd[17] = { 0 };
    for (int i = min_idx; i < max_idx_p1; i++) {
        int cw = d->lmcs_delta_sign_cw_flag[i] ? -(int)d->lmcs_delta_abs_cw[i] : (int)d->lmcs_delta_abs_cw[i];
        int wnd_sz = smp_wnd + (int16_t)cw;
        if (wnd_sz) {
            fwd_step[i] = (uint16_t)(((wnd_sz << prec) + (1 << (log2_wnd - 1))) >> log2_wnd);
            bwd_step[i] = (uint16_t)((smp_wnd << prec) / wnd_sz);
        }
        bnd[i + 1] = (uint16_t)(bnd[i] + wnd_sz);
    }
    for (int i = max_idx_p1; i < nb_wnd; i++) bnd[i + 1] = bnd[i];
    for (int val = 0; val < rng; val++) {
        int idx = val >> log2_wnd, nb_step = val - (idx << log2_wnd);
        int v = (int16_t)bnd[idx] + (((int32_t)fwd_step[idx] * nb_step + rnd) >> prec);
        luts->fwd_lut[val] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
    }
    for (int val = 0; val < rng; val++) {
        int i = min_idx;
        for (; i < max_idx_p1; i++)
            if (val < bnd[i + 1]) break;
        int idx = i < nb_wnd - 1 ? i : nb_wnd - 1;
        int v = (idx << log2_wnd) + (((int32_t)bwd_step[idx] * (val - bnd[idx]) + rnd) >> prec);
        luts->bwd_lut[val] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
    }
    for (int i = 0; i < 17; i++) luts->wnd_bnd[i] = bnd[i];
}

// whole-plane LUT pass (lmcs_reshape_backward on every CTU, slicedec.c:745-750): 8 samples per thread
struct LmcsReshapeK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 1;
    struct Args { uint16_t *plane; int pitch, w, h, bitdepth; const uint16_t *lut; };
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st; (void)bz;
        const int x = (bx * 32 + (tid & 31)) * 8, y = by * 8 + (tid >> 5);
        if (x >= a.w || y >= a.h) return;
        const int msk = (1 << a.bitdepth) - 1;
        uint16_t *p = a.plane + (size_t)y * a.pitch + x;
        ovd_u4 v = ovd_ld16_rw(p);         // in place: not the read-only (.nc) path
        uint32_t in[4] = { v.x, v.y, v.z, v.w }, o[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            o[i] = (uint32_t)a.lut[in[i] & msk] | ((uint32_t)a.lut[(in[i] >> 16) & msk] << 16);
        ovd_u4 r; r.x = o[0]; r.y = o[1]; r.z = o[2]; r.w = o[3];
        ovd_st16(p, r);
    }
};
