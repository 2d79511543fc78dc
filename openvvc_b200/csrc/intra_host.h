// intra_host.h -- dependency levels of a picture's intra blocks, shared by intra.cu and tests/emu.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/ovvc_b200.h"
#include "intra_kernels.h"

static inline int intra_check_cmd(const OVGPUIntraCmd *c, int pic_w, int pic_h)
{
    if (c->log2_w < 2 || c->log2_w > 6 || c->log2_h < 2 || c->log2_h > 6 || c->x0 < 0 || c->y0 < 0) return -1;
    if (!(c->flags & (OVGPU_INTRA_LUMA | OVGPU_INTRA_CHROMA)) || (c->flags & ~31)) return -1;
    if (c->flags & OVGPU_INTRA_MIP) {
        const int small = c->log2_w == 2 && c->log2_h == 2, red = c->log2_w == 2 || c->log2_h == 2 || (c->log2_w <= 3 && c->log2_h <= 3);
        if (!(c->flags & OVGPU_INTRA_LUMA) || (c->flags & OVGPU_INTRA_CIIP) || c->mrl_idx || (c->mode & 0x40) ||
            (c->mode & 0x3f) >= (small ? 16 : (red ? 8 : 6))) return -1;
    } else if ((c->flags & OVGPU_INTRA_LUMA) && c->mode > 66) return -1;
    if ( ((c->flags & OVGPU_INTRA_CHROMA) && c->mode_c > 69)) return -1;
    if ((c->flags & OVGPU_INTRA_CIIP) && (c->ciip_wt < 1 || c->ciip_wt > 3)) return -1;
    if (c->mrl_idx > 2 || (c->mrl_idx && (c->y0 < 4 || c->x0 + (1 << c->log2_w) > pic_w))) return -1;
    if ((c->x0 & 3) || (c->y0 & 3) || c->x0 + (1 << c->log2_w) > pic_w || c->y0 + (1 << c->log2_h) > pic_h) return -1;
    return 0;
}

// level[i] = 1 + the highest level among earlier blocks that cover one of block i's reference samples (left column of
// 2h + 1 and above row of 2w + 1 samples, on the 4x4 unit grid).  Luma and chroma keep separate level maps (dual tree:
// a chroma-only command depends on earlier chroma blocks, and -- for CCLM -- on the luma blocks under and around it).
// order[] = command indices sorted by level (stable), level_start[l] .. level_start[l + 1] = the blocks of level l
// (0-based).  map_l / map_c (optional, w/4 x h/4 entries each) receive the 1-based level of the command covering each
// unit (0 = none): the residual schedule (tr_host.h) places transform blocks with them.  Returns the number of levels or -1.
static inline int intra_neigh_level(const uint16_t *m, int w_u, int h_u, int xu, int yu, int nw, int nh, int own)
{
    int lv = 0;
    if (xu > 0)
        for (int y = yu - 1; y < yu + 2 * nh && y < h_u; y++) if (y >= 0 && m[y * w_u + xu - 1] > lv) lv = m[y * w_u + xu - 1];
    if (yu > 0)
        for (int x = xu - 1; x < xu + 2 * nw && x < w_u; x++) if (x >= 0 && m[(yu - 1) * w_u + x] > lv) lv = m[(yu - 1) * w_u + x];
    if (own)
        for (int y = yu; y < yu + nh; y++) for (int x = xu; x < xu + nw; x++) if (m[y * w_u + x] > lv) lv = m[y * w_u + x];
    return lv;
}

static inline int intra_build_levels(uint32_t *order, uint32_t *level_start, int max_levels, const OVGPUIntraCmd *cmds, int n,
                                     int pic_w, int pic_h, uint16_t *map_l_out = NULL, uint16_t *map_c_out = NULL)
{
    const int w_u = pic_w >> 2, h_u = pic_h >> 2;
    uint16_t *map_l = map_l_out ? map_l_out : (uint16_t *)malloc((size_t)w_u * h_u * sizeof(uint16_t));
    uint16_t *map_c = map_c_out ? map_c_out : (uint16_t *)malloc((size_t)w_u * h_u * sizeof(uint16_t));
    uint16_t *lvl = (uint16_t *)malloc(sizeof(uint16_t) * (n > 0 ? n : 1));
    uint32_t *fill = NULL;
    int nb_levels = 0, ret = -1;
    if (!map_l || !map_c || !lvl) goto done;
    memset(map_l, 0, (size_t)w_u * h_u * sizeof(uint16_t));
    memset(map_c, 0, (size_t)w_u * h_u * sizeof(uint16_t));
    for (int i = 0; i < n; i++) {
        const OVGPUIntraCmd *c = &cmds[i];
        const int xu = c->x0 >> 2, yu = c->y0 >> 2, nw = (1 << c->log2_w) >> 2, nh = (1 << c->log2_h) >> 2;
        const int luma = c->flags & OVGPU_INTRA_LUMA, chroma = c->flags & OVGPU_INTRA_CHROMA;
        int m = 0, v;
        if (luma) m = intra_neigh_level(map_l, w_u, h_u, xu, yu, nw, nh, 0);
        if (chroma) {
            if ((v = intra_neigh_level(map_c, w_u, h_u, xu, yu, nw, nh, 0)) > m) m = v;
            if (c->mode_c >= 67 && !luma && (v = intra_neigh_level(map_l, w_u, h_u, xu, yu, nw, nh, 1)) > m) m = v;
        }
        if (m + 1 > 65535 || m + 1 > max_levels) goto done;
        lvl[i] = (uint16_t)(m + 1);
        if (m + 1 > nb_levels) nb_levels = m + 1;
        for (int y = yu; y < yu + nh; y++)
            for (int x = xu; x < xu + nw; x++) {
                if (luma) map_l[y * w_u + x] = lvl[i];
                if (chroma) map_c[y * w_u + x] = lvl[i];
            }
    }
    // counting sort by level
    memset(level_start, 0, sizeof(uint32_t) * (nb_levels + 2));
    for (int i = 0; i < n; i++) level_start[lvl[i]]++;                 // counts at index level (1-based)
    {
        uint32_t acc = 0;
        for (int l = 1; l <= nb_levels; l++) { uint32_t cnt = level_start[l]; level_start[l] = acc; acc += cnt; }
        level_start[nb_levels + 1] = acc;
    }
    fill = (uint32_t *)malloc(sizeof(uint32_t) * (nb_levels + 2));
    if (!fill) goto done;
    memcpy(fill, level_start, sizeof(uint32_t) * (nb_levels + 2));
    for (int i = 0; i < n; i++) order[fill[lvl[i]]++] = (uint32_t)i;
    // shift to 0-based: level_start[l] for l in 0..nb_levels
    for (int l = 0; l <= nb_levels; l++) level_start[l] = level_start[l + 1];
    ret = nb_levels;
done:
    free(fill); free(lvl);
    if (!map_l_out) free(map_l);
    if (!map_c_out) free(map_c);
    return ret;
}

// Launch sequence of a picture: one grid per dependency level covering the components present; a level whose chroma
// blocks may use CCLM (they read the luma block of the same command) runs luma first, then Cb / Cr.
// f(first command of the level in order[], number of commands, first component, number of components) -> 0 on success
template <typename F>
static inline int intra_for_each_launch(int nb_levels, const uint32_t *level_start, int has_chroma, int has_cclm, F f)
{
    for (int l = 0; l < nb_levels; l++) {
        const int start = (int)level_start[l], n = (int)level_start[l + 1] - start;
        if (n <= 0) continue;
        int rc;
        if (has_cclm) { if ((rc = f(start, n, 0, 1)) != 0 || (rc = f(start, n, 1, 2)) != 0) return rc; }
        else if ((rc = f(start, n, 0, has_chroma ? 3 : 1)) != 0) return rc;
    }
    return 0;
}

// Whole reconstruction order of a picture (replaces the reference's CU loop: prediction, then rcn_transform_tree per CU,
// vcl_coding_unit.c / rcn_transform_tree.c:1433): transform blocks outside intra commands first (group 0: their
// predictions come from motion compensation, done before), then level by level the intra predictions followed by the
// residual of the blocks they cover.  With CCLM a level is run luma first (prediction + residual), then chroma.
// fi(first command in order[], count, first component, number of components), ft(residual group) -> 0 on success
template <typename FI, typename FT>
static inline int recon_for_each_launch(int nb_levels, const uint32_t *level_start, int has_chroma, int has_cclm, FI fi, FT ft)
{
    const int split = has_cclm ? 2 : 1;
    int rc;
    for (int s = 0; s < split; s++) if ((rc = ft(s)) != 0) return rc;
    for (int l = 0; l < nb_levels; l++) {
        const int start = (int)level_start[l], n = (int)level_start[l + 1] - start;
        if (n <= 0) continue;
        if (has_cclm) {
            if ((rc = fi(start, n, 0, 1)) != 0 || (rc = ft((l + 1) * 2)) != 0 ||
                (rc = fi(start, n, 1, 2)) != 0 || (rc = ft((l + 1) * 2 + 1)) != 0) return rc;
        } else if ((rc = fi(start, n, 0, has_chroma ? 3 : 1)) != 0 || (rc = ft(l + 1)) != 0) return rc;
    }
    return 0;
}
