// mc_host.h -- routing of MC commands to kernels and the work lists: built ON THE DEVICE from the uploaded commands (McPrep*K
// below; the host only copies the commands and reads back six counters), with a plain host restatement that tests/emu checks
// the kernels against.
#pragma once
#include <stdint.h>
#include "../../include/ovvc_b200.h"
#include "mc_tools_kernels.h"
#include "mc_big_kernels.h"
#include "mc_rpr_kernels.h"

#define MC_ROUTE_BIG 6
#define MC_ROUTE_RPR 7

// Reference picture resampling: `sm` = per list the reference indices whose scale factor is not 1 (from the batch's scaling
// table, ovgpu_mc_batch_set_scaling; NULL: none).  A command that names such a reference takes the resampling kernel.
OVHD int mc_uses_scaled(const OVGPUMcCmd *c, const uint32_t *sm)
{
    if (!sm || !(sm[0] | sm[1])) return 0;
    if (c->tool == OVGPU_MC_TOOL_GPM) return (int)(((sm[(c->aux >> 8) & 1] >> (c->ref_idx0 & 15)) | (sm[(c->aux >> 9) & 1] >> (c->ref_idx1 & 15))) & 1);
    return (int)((((c->inter_dir & 1) ? sm[0] >> (c->ref_idx0 & 15) : 0) | ((c->inter_dir & 2) ? sm[1] >> (c->ref_idx1 & 15) : 0)) & 1);
}

// Which kernel runs a command: plain 4x4 luma-only commands (affine sub-blocks without PROF) take the packed 4x4 path of
// ProfK (four sub-blocks per warp, refinement off) instead of a whole McK warp each.
OVHD int mc_route(const OVGPUMcCmd *c, const uint32_t *sm = NULL)
{
    if (mc_uses_scaled(c, sm)) return MC_ROUTE_RPR;
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w == 2 && c->log2_h == 2 && (c->flags & 3) == OVGPU_MC_LUMA) return OVGPU_MC_TOOL_PROF;
    // plain chroma-only commands of 8x8 luma samples (the chroma blocks of affine CUs): two per warp in McSmallCK
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w == 3 && c->log2_h == 3 && (c->flags & 3) == OVGPU_MC_CHROMA && !(c->x0 & 3)) return MC_ROUTE_SMALL_C;
    // plain commands of at least 32 x 16 luma samples: 32 x 32 / 32 x 16 tiles in McBigK
    if (c->tool == OVGPU_MC_TOOL_PLAIN && c->log2_w >= 5 && c->log2_h >= 4) return MC_ROUTE_BIG;
    return c->tool == OVGPU_MC_TOOL_GPM ? OVGPU_MC_TOOL_PLAIN : c->tool;
}

// number of <=16x16 work items of the plain (tool 0) commands of a PU list
static inline int mc_count_items(const OVGPUMcCmd *cmds, int n, const uint32_t *sm = NULL, int route = OVGPU_MC_TOOL_PLAIN)
{
    int total = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i], sm) != route) continue;
        int w = 1 << cmds[i].log2_w, h = 1 << cmds[i].log2_h;
        total += ((w + 15) >> 4) * ((h + 15) >> 4);
    }
    return total;
}

// BDOF / PROF commands are one work item each (BdofK: one warp per item, ProfK: four items per warp)
static inline int mc_count_tool_items(const OVGPUMcCmd *cmds, int n, int tool, const uint32_t *sm = NULL)
{
    int total = 0;
    for (int i = 0; i < n; i++) total += mc_route(&cmds[i], sm) == tool;
    return total;
}

static inline int mc_build_tool_items(McItemDev *items, const OVGPUMcCmd *cmds, int n, int tool, const uint32_t *sm = NULL)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i], sm) != tool) continue;
        items[k].cmd = (uint32_t)i; items[k].ox = items[k].oy = 0;
        items[k].w = (uint8_t)(1 << cmds[i].log2_w); items[k].h = (uint8_t)(1 << cmds[i].log2_h);
        k++;
    }
    return k;
}

// well-formedness of one command; nb_prof = entries of the PROF table
OVHD int mc_check_cmd(const OVGPUMcCmd *c, int nb_prof, const uint32_t *sm = NULL)
{
    if (c->log2_w < 2 || c->log2_w > 7 || c->log2_h < 2 || c->log2_h > 7 || c->inter_dir < 1 || c->inter_dir > 3 ||
        c->ref_idx0 < 0 || c->ref_idx0 > 15 || c->ref_idx1 < 0 || c->ref_idx1 > 15 || c->bcw_idx_plus1 > 5 || c->x0 < 0 || c->y0 < 0)
        return -1;
    // the decoder switches BDOF and DMVR off for CUs with a scaled reference (and the reference has no such path)
    if ((c->tool == OVGPU_MC_TOOL_BDOF || c->tool == OVGPU_MC_TOOL_DMVR) && mc_uses_scaled(c, sm)) return -1;
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
static inline int mc_build_big_items(McItemDev *items, const OVGPUMcCmd *cmds, int n, const uint32_t *sm = NULL)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i], sm) != MC_ROUTE_BIG) continue;
        const int w = 1 << cmds[i].log2_w, h = 1 << cmds[i].log2_h, th = h >= 32 ? 32 : 16;
        for (int oy = 0; oy < h; oy += th)
            for (int ox = 0; ox < w; ox += 32) {
                if (items) { items[k].cmd = (uint32_t)i; items[k].ox = (uint8_t)ox; items[k].oy = (uint8_t)oy; items[k].w = 32; items[k].h = (uint8_t)th; }
                k++;
            }
    }
    return k;
}

static inline int mc_build_items(McItemDev *items, const OVGPUMcCmd *cmds, int n, const uint32_t *sm = NULL, int route = OVGPU_MC_TOOL_PLAIN)
{
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (mc_route(&cmds[i], sm) != route) continue;
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
#ifndef LMCS_MIN_BLOCKS
#define LMCS_MIN_BLOCKS 4
#endif
struct LmcsReshapeK {
    static const int THREADS = 256, MIN_BLOCKS = LMCS_MIN_BLOCKS, PHASES = 1;
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

// ==================================================================================================================
// Work lists built on the device.  The host used to walk the picture's commands (validate, route, split into work items:
// 0.87 ms per 4K picture and host thread, the largest single piece of host work of the end-to-end path); now it copies the
// commands and three small kernels do the walk:
//   McPrepCountK    one thread per command: validate, route, number of work items; per block and region the item count,
//                   picture-wide the furthest sample written, the references named, the first malformed command
//   McPrepScanK     one CTA: exclusive scan of the block counts per region -> block offsets, totals (the six counters the
//                   host reads back to size the launches), overflow of a region's capacity
//   McPrepScatterK  one thread per command again: its items written at block offset + offset inside the block -- so every
//                   region lists its items in COMMAND ORDER (the refined MVs of DMVR items are returned in that order)
// Regions of the item array: 0 plain (<=16x16 blocks, GPM), 1 BDOF, 2 PROF / plain 4x4, 3 DMVR, 4 small chroma, 5 big tiles,
// 6 commands with a scaled reference (<=16x16 blocks, McRprK).
#define MC_NB_REGIONS 7
// region and number of work items of a (well-formed) command
OVHD int mc_region_items(const OVGPUMcCmd *c, int *nb, const uint32_t *sm = NULL)
{
    const int route = mc_route(c, sm), w = 1 << c->log2_w, h = 1 << c->log2_h;
    *nb = 1;
    if (route == MC_ROUTE_RPR) { *nb = ((w + 15) >> 4) * ((h + 15) >> 4); return 6; }
    if (route == OVGPU_MC_TOOL_PLAIN) { *nb = ((w + 15) >> 4) * ((h + 15) >> 4); return 0; }
    if (route == MC_ROUTE_BIG) { *nb = (w >> 5) * (h >= 32 ? h >> 5 : 1); return 5; }
    return route == OVGPU_MC_TOOL_BDOF ? 1 : route == OVGPU_MC_TOOL_DMVR ? 3 : route == MC_ROUTE_SMALL_C ? 4 : 2;
}

struct McPrepInfo {                  // device result, copied to pinned host memory
    int32_t count[MC_NB_REGIONS];    // work items per region
    int32_t bad_rev;                 // nb_cmds - index of the first malformed command (0: every command is well-formed)
    int32_t max_x, max_y;            // furthest sample any command writes
    uint32_t used_refs[2];           // reference indices named, per list
    int32_t overflow;                // 1 + a region whose items did not fit (0: none)
    int32_t seq;                     // (unused)
    // written by McGuardK when the batch runs on a picture:
    int32_t limit[MC_NB_REGIONS];    // items the MC kernels may process: count[], or 0 everywhere when status != 0
    int32_t status;                  // 0 ok, 1 malformed command (bad_rev), 2 too many items, 3 commands outside the picture, 4 reference missing
    int32_t status_arg;              // the command / region / list * 16 + index concerned
};

// Number of work items per region as the HOST counts them, to size the launches without waiting for the device: no validation,
// no lists, three byte loads and a table look-up per command (fields masked so that garbage cannot index out of range; the
// device-side walk is the one that validates, and the kernels stop at ITS counts).
struct McCountLut {                  // (tool, log2_w, log2_h, flags & 3, x0 & 3 != 0) -> region << 8 | items
    uint16_t v[8 * 8 * 8 * 4 * 2];
    McCountLut()
    {
        for (int key = 0; key < 8 * 8 * 8 * 4 * 2; key++) {
            OVGPUMcCmd c;
            memset(&c, 0, sizeof(c));
            c.tool = (uint8_t)(key >> 9); c.log2_w = (key >> 6) & 7; c.log2_h = (key >> 3) & 7; c.flags = (key >> 1) & 3; c.x0 = (int16_t)(key & 1);
            int k;
            const int r = mc_region_items(&c, &k);
            v[key] = (uint16_t)((r << 8) | k);
        }
    }
};
static inline void mc_count_regions(const OVGPUMcCmd *cmds, int n, int cap, int cnt[MC_NB_REGIONS], const uint32_t *sm = NULL)
{
    static const McCountLut lut;
    long long c64[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const int any_scaled = sm && (sm[0] | sm[1]);
    for (int i = 0; i < n; i++) {
        const OVGPUMcCmd *c = &cmds[i];
        if (any_scaled && mc_uses_scaled(c, sm)) { c64[6] += (((1 << (c->log2_w & 7)) + 15) >> 4) * (((1 << (c->log2_h & 7)) + 15) >> 4); continue; }
        const unsigned key = ((c->tool & 7u) << 9) | ((c->log2_w & 7u) << 6) | ((c->log2_h & 7u) << 3) | ((c->flags & 3u) << 1) | ((c->x0 & 3) != 0);
        const unsigned e = lut.v[key];
        c64[e >> 8] += e & 0xff;
    }
    for (int r = 0; r < MC_NB_REGIONS; r++) cnt[r] = (int)(c64[r] < cap ? c64[r] : cap);
}

// One warp: the walk's findings against the picture the batch is about to run on -> limit[] / status
struct McGuardArgs { McPrepInfo *info; int pic_w, pic_h; uint32_t present[2]; McPrepInfo *info_host; };   // info_host: mapped pinned copy (or NULL)
struct McGuardK {
    static const int THREADS = 32, MIN_BLOCKS = 1, PHASES = 1;
    typedef McGuardArgs Args;
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st; (void)bx; (void)by; (void)bz;
        if (tid) return;
        McPrepInfo *f = a.info;
        int status = 0, arg = 0;
        if (f->bad_rev) { status = 1; arg = f->bad_rev; }
        else if (f->overflow) { status = 2; arg = f->overflow - 1; }
        else if (f->max_x > a.pic_w || f->max_y > a.pic_h) status = 3;
        else
            for (int l = 0; l < 2 && !status; l++) {
                const uint32_t miss = f->used_refs[l] & ~a.present[l];
                if (miss) { status = 4; arg = l * 16 + ovd_floor_log2(miss & (0u - miss)); }
            }
        for (int r = 0; r < MC_NB_REGIONS; r++) f->limit[r] = status ? 0 : f->count[r];
        f->status = status; f->status_arg = arg;
        if (a.info_host) *a.info_host = *f;                // the verdict goes to the host as a store over PCIe, not as a DMA copy
    }
};

// The commands are read where the recorder wrote them -- PINNED HOST memory, through the mapping every cudaMallocHost buffer has
// (zero copy) -- and McPrepCountK leaves the device copy the MC kernels use; the counters go back the same way (a store to mapped
// No copy engine is involved: a DMA copy of the commands would queue behind the megabytes of coefficient and picture copies of
// the other pictures in flight.  The HOST NEVER WAITS for the walk: it sizes the launches from its own cheap count of the commands
// (mc_count_regions) and the kernels stop at the walk's counts; what the walk refuses (a malformed command ...) zeroes those
// counts (McGuardK) and is reported by ovgpu_mc_batch_status / the next call that synchronises with the batch.
struct McPrepArgs {
    const McCmdDev *src;             // commands as recorded (mapped pinned host memory, or plain memory under emulation)
    McCmdDev *cmds;                  // device copy written by McPrepCountK (== src: none needed)
    const uint32_t *prof_src;        // PROF table as recorded / its device copy, in 32-bit words
    uint32_t *prof;
    int nb_cmds, nb_prof, nb_blocks, max_items;
    uint32_t scaled[2];              // reference indices with a scale factor other than 1, per list (0: no resampling)
    int32_t *block_sum;              // [region][block] item counts, then (after the scan) exclusive offsets
    McItemDev *items;                // MC_NB_REGIONS regions of max_items entries
    McPrepInfo *info;                // device
};

struct McPrepCountK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 2;
    static const int SMEM_BYTES = THREADS * 6 * 4;
    typedef McPrepArgs Args;
    struct State { int dummy; };
    // per thread: region (-1: no command / malformed), items, extents, references
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        int *s_reg = (int *)smem, *s_cnt = s_reg + THREADS, *s_mx = s_cnt + THREADS, *s_my = s_mx + THREADS;
        uint32_t *s_u0 = (uint32_t *)(s_my + THREADS), *s_u1 = s_u0 + THREADS;
        if (P == 0) {
            const int i = bx * THREADS + tid;
            int reg = -1, cnt = 0, mx = 0, my = 0;
            uint32_t u0 = 0, u1 = 0;
            if (a.prof != a.prof_src)        // PROF table: 32 words per entry
                for (int j = i; j < a.nb_prof * 32; j += a.nb_blocks * THREADS) a.prof[j] = a.prof_src[j];
            if (i < a.nb_cmds) {
                const McCmdDev cd = a.src[i];
                if (a.cmds != a.src) a.cmds[i] = cd;
                const OVGPUMcCmd *c = (const OVGPUMcCmd *)&cd;
                if (mc_check_cmd(c, a.nb_prof, a.scaled)) ovd_atomic_max(&a.info->bad_rev, a.nb_cmds - i);
                else {
                    reg = mc_region_items(c, &cnt, a.scaled);
                    mx = c->x0 + (1 << c->log2_w); my = c->y0 + (1 << c->log2_h);
                    if (c->tool == OVGPU_MC_TOOL_GPM) { if ((c->aux >> 8) & 1) u1 |= 1u << c->ref_idx0; else u0 |= 1u << c->ref_idx0;
                                                        if ((c->aux >> 9) & 1) u1 |= 1u << c->ref_idx1; else u0 |= 1u << c->ref_idx1; }
                    else { if (c->inter_dir & 1) u0 |= 1u << c->ref_idx0; if (c->inter_dir & 2) u1 |= 1u << c->ref_idx1; }
                }
            }
            s_reg[tid] = reg; s_cnt[tid] = cnt; s_mx[tid] = mx; s_my[tid] = my; s_u0[tid] = u0; s_u1[tid] = u1;
            return;
        }
        if (tid < MC_NB_REGIONS) {
            int sum = 0;
            for (int t = 0; t < THREADS; t++) if (s_reg[t] == tid) sum += s_cnt[t];
            a.block_sum[tid * a.nb_blocks + bx] = sum;
        } else if (tid == 32) {
            int mx = 0, my = 0;
            uint32_t u0 = 0, u1 = 0;
            for (int t = 0; t < THREADS; t++) { mx = ovd_max(mx, s_mx[t]); my = ovd_max(my, s_my[t]); u0 |= s_u0[t]; u1 |= s_u1[t]; }
            ovd_atomic_max(&a.info->max_x, mx); ovd_atomic_max(&a.info->max_y, my);
            if (u0) ovd_atomic_or(&a.info->used_refs[0], u0);
            if (u1) ovd_atomic_or(&a.info->used_refs[1], u1);
        }
    }
};

struct McPrepScanK {
    static const int THREADS = 256, MIN_BLOCKS = 1, PHASES = 3;
    static const int SMEM_BYTES = MC_NB_REGIONS * THREADS * 4;
    typedef McPrepArgs Args;
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)bx; (void)by; (void)bz;
        int *part = (int *)smem;                                  // [region][thread]: sum of the thread's chunk of blocks, then its offset
        const int chunk = (a.nb_blocks + THREADS - 1) / THREADS, b0 = tid * chunk, b1 = ovd_min(b0 + chunk, a.nb_blocks);
        if (P == 0) {
            for (int r = 0; r < MC_NB_REGIONS; r++) {
                int s = 0;
                for (int b = b0; b < b1; b++) s += a.block_sum[r * a.nb_blocks + b];
                part[r * THREADS + tid] = s;
            }
            return;
        }
        if (P == 1) {
            if (tid < MC_NB_REGIONS) {
                int acc = 0;
                for (int t = 0; t < THREADS; t++) { const int v = part[tid * THREADS + t]; part[tid * THREADS + t] = acc; acc += v; }
                a.info->count[tid] = acc;
                if (acc > a.max_items) ovd_atomic_max(&a.info->overflow, tid + 1);
            }
            return;
        }
        for (int r = 0; r < MC_NB_REGIONS; r++) {
            int acc = part[r * THREADS + tid];
            for (int b = b0; b < b1; b++) { const int v = a.block_sum[r * a.nb_blocks + b]; a.block_sum[r * a.nb_blocks + b] = acc; acc += v; }
        }

    }
};

struct McPrepScatterK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 3;
    static const int SMEM_BYTES = THREADS * 3 * 4;
    typedef McPrepArgs Args;
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st; (void)by; (void)bz;
        int *s_reg = (int *)smem, *s_cnt = s_reg + THREADS, *s_off = s_cnt + THREADS;
        const int i = bx * THREADS + tid;
        if (P == 0) {
            int reg = -1, cnt = 0;
            if (i < a.nb_cmds) {
                const McCmdDev cd = a.cmds[i];
                const OVGPUMcCmd *c = (const OVGPUMcCmd *)&cd;
                if (!mc_check_cmd(c, a.nb_prof, a.scaled)) reg = mc_region_items(c, &cnt, a.scaled);
            }
            s_reg[tid] = reg; s_cnt[tid] = cnt;
            return;
        }
        if (P == 1) {
            if (tid < MC_NB_REGIONS) {
                int acc = a.block_sum[tid * a.nb_blocks + bx];
                for (int t = 0; t < THREADS; t++) if (s_reg[t] == tid) { s_off[t] = acc; acc += s_cnt[t]; }
            }
            return;
        }
        const int reg = s_reg[tid], cnt = s_cnt[tid], off = reg >= 0 ? s_off[tid] : 0;
        if (reg < 0 || off + cnt > a.max_items) return;           // (an overflowing region is reported by McPrepScanK)
        McItemDev *it = a.items + (size_t)reg * a.max_items + off;
        const McCmdDev cd = a.cmds[i];
        const int w = 1 << cd.log2_w, h = 1 << cd.log2_h;
        if (reg == 0 || reg == 5 || reg == 6) {
            const int tw = reg == 5 ? 32 : 16, th = reg == 5 ? (h >= 32 ? 32 : 16) : 16;
            int k = 0;
            for (int oy = 0; oy < h; oy += th)
                for (int ox = 0; ox < w; ox += tw, k++) {
                    McItemDev v;
                    v.cmd = (uint32_t)i; v.ox = (uint8_t)ox; v.oy = (uint8_t)oy;
                    v.w = (uint8_t)(w - ox < tw ? w - ox : tw); v.h = (uint8_t)(h - oy < th ? h - oy : th);
                    it[k] = v;
                }
        } else {
            McItemDev v;
            v.cmd = (uint32_t)i; v.ox = v.oy = 0; v.w = (uint8_t)w; v.h = (uint8_t)h;
            it[0] = v;
        }
    }
};
