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
    if (!(c->flags & (OVGPU_INTRA_LUMA | OVGPU_INTRA_CHROMA))) return -1;
    if (c->flags & OVGPU_INTRA_IBC) {       /* block copy: nothing else applies; the source block lies in the picture and apart from the block */
        if ((c->flags & ~(OVGPU_INTRA_IBC | OVGPU_INTRA_LUMA | OVGPU_INTRA_CHROMA)) || c->mrl_idx) return -1;
        if ((c->x0 & 3) || (c->y0 & 3) || c->x0 + (1 << c->log2_w) > pic_w || c->y0 + (1 << c->log2_h) > pic_h) return -1;
        const int bvx = (int16_t)(c->pad[0] | (c->pad[1] << 8)), bvy = (int16_t)(c->pad[2] | (c->pad[3] << 8));
        const int sx = c->x0 + bvx, sy = c->y0 + bvy, w = 1 << c->log2_w, h = 1 << c->log2_h;
        if (sx < 0 || sy < 0 || sx + w > pic_w || sy + h > pic_h) return -1;
        if (sx < c->x0 + w && sx + w > c->x0 && sy < c->y0 + h && sy + h > c->y0) return -1;
        return 0;
    }
    if (c->flags & OVGPU_INTRA_ISP) {       /* whole luma coding block of 32 samples or more, plain modes only */
        if (!(c->flags & OVGPU_INTRA_LUMA) || (c->flags & (OVGPU_INTRA_CIIP | OVGPU_INTRA_MIP)) || c->mrl_idx || c->log2_w + c->log2_h < 5) return -1;
    } else if (c->flags & OVGPU_INTRA_ISP_VER) return -1;
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

// Dependency levels, per plane.  Luma never reads chroma, so the whole luma plane can be reconstructed before any chroma
// (and CCLM / the LMCS chroma scale then read final luma): the luma blocks and the chroma blocks of a picture get
// independent level numbers.  level = 1 + the highest level among earlier blocks of the same plane that cover one of the
// block's reference samples (left column of 2h + 1 and above row of 2w + 1 samples, on the 4x4 luma-unit grid).
// Per plane p (0 luma, 1 chroma): order[p][] = indices of the commands that have that plane, sorted by level (stable);
// start[p][l] .. start[p][l + 1] = the blocks of level l (0-based); nb[p] = number of levels.
// pos_l / pos_c (optional, w/4 x h/4 entries each) receive 1 + the position in order[plane] of the command covering each
// unit (0 = none): the residual schedule (tr_host.h) attaches transform blocks to commands with them.
struct IntraLevels {
    int nb[2];
    uint32_t *order[2];      // n entries each
    uint32_t *start[2];      // n + 2 entries each
    // optional caller-owned scratch (a per-picture malloc / free of megabytes costs page faults and serialises host threads
    // on the process's memory-map lock): level map of w/4 x h/4 uint16, 2 * n uint16, n + 3 uint32
    uint16_t *ws_map; uint16_t *ws_lvl; uint32_t *ws_fill;
    // ws_clean: ws_map and the position maps handed to intra_build_levels are all zero on entry (and ws_map is left all zero):
    // clearing megabytes of maps per picture costs more than everything else in here when only a tenth of the picture is intra
    int ws_clean;
    // (CTU size >> 2) - 1: with it the availability words of every command are checked against the queue order (0: no check)
    int units_mask;
};

// zeroes the units of the commands' blocks in a position / level map (the inverse of what intra_build_levels writes)
template <typename T>
static inline void intra_clear_units(T *map, int w_u, const OVGPUIntraCmd *cmds, int n, int bit)
{
    for (int i = 0; i < n; i++) {
        const OVGPUIntraCmd *c = &cmds[i];
        if (!(c->flags & bit)) continue;
        const int xu = c->x0 >> 2, yu = c->y0 >> 2, nw = (1 << c->log2_w) >> 2, nh = (1 << c->log2_h) >> 2;
        for (int y = yu; y < yu + nh; y++) memset(map + (size_t)y * w_u + xu, 0, sizeof(T) * nw);
    }
}

static inline int intra_neigh_level(const uint16_t *m, int w_u, int h_u, int xu, int yu, int nw, int nh)
{
    int lv = 0;
    if (xu > 0)
        for (int y = yu - 1; y < yu + 2 * nh && y < h_u; y++) if (y >= 0 && m[y * w_u + xu - 1] > lv) lv = m[y * w_u + xu - 1];
    if (yu > 0)
        for (int x = xu - 1; x < xu + 2 * nw && x < w_u; x++) if (x >= 0 && m[(yu - 1) * w_u + x] > lv) lv = m[(yu - 1) * w_u + x];
    return lv;
}

// Returns 0, -1 (out of memory / more than max_levels levels) or -2 (a block copy reads samples that are not reconstructed before it).
static inline int intra_build_levels(IntraLevels *L, int max_levels, const OVGPUIntraCmd *cmds, int n,
                                     int pic_w, int pic_h, uint32_t *pos_l = NULL, uint32_t *pos_c = NULL)
{
    const int w_u = pic_w >> 2, h_u = pic_h >> 2;
    uint32_t *pos[2] = { pos_l, pos_c };
    const int own = !L->ws_map;
    uint16_t *map1 = own ? (uint16_t *)malloc((size_t)w_u * h_u * sizeof(uint16_t)) : L->ws_map;       // one plane at a time
    uint16_t *map[2] = { map1, map1 };
    uint16_t *lvl = own ? (uint16_t *)malloc(sizeof(uint16_t) * 2 * (n > 0 ? n : 1)) : L->ws_lvl;
    uint32_t *fill = own ? (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + 3)) : L->ws_fill;
    int ret = -1;
    L->nb[0] = L->nb[1] = 0;
    if (map[0] && map[1] && lvl && fill) {
        ret = 0;
        for (int p = 0; p < 2 && !ret; p++) {
            const int bit = p ? OVGPU_INTRA_CHROMA : OVGPU_INTRA_LUMA;
            uint16_t *m = map[p], *lv = lvl + (size_t)p * n;
            int nb_levels = 0;
            const int clean = !own && L->ws_clean;
            if (!clean) memset(m, 0, (size_t)w_u * h_u * sizeof(uint16_t));
            for (int i = 0; i < n; i++) {
                const OVGPUIntraCmd *c = &cmds[i];
                lv[i] = 0;
                if (!(c->flags & bit)) continue;
                const int xu = c->x0 >> 2, yu = c->y0 >> 2, nw = (1 << c->log2_w) >> 2, nh = (1 << c->log2_h) >> 2;
                int l;
                if (c->flags & OVGPU_INTRA_IBC) {       // block copy: depends on the commands covering its source block
                    if (!pos[p]) { ret = -2; break; }   // (validated against the queue below, which needs the position map)
                    int bvx = (int16_t)(c->pad[0] | (c->pad[1] << 8)), bvy = (int16_t)(c->pad[2] | (c->pad[3] << 8));
                    l = 0;
                    for (int y = (c->y0 + bvy) >> 2; y <= (c->y0 + bvy + (nh << 2) - 1) >> 2; y++)
                        for (int x = (c->x0 + bvx) >> 2; x <= (c->x0 + bvx + (nw << 2) - 1) >> 2; x++) if (m[y * w_u + x] > l) l = m[y * w_u + x];
                    l++;
                } else l = intra_neigh_level(m, w_u, h_u, xu, yu, nw, nh) + 1;
                if (l > 65535 || l > max_levels) { ret = -1; break; }
                lv[i] = (uint16_t)l;
                if (l > nb_levels) nb_levels = l;
                for (int y = yu; y < yu + nh; y++) for (int x = xu; x < xu + nw; x++) m[y * w_u + x] = (uint16_t)l;
            }
            if (clean) intra_clear_units(m, w_u, cmds, n, bit);
            if (ret) break;
            // counting sort by level (level 0 = the command does not have this plane)
            uint32_t *start = L->start[p];
            memset(start, 0, sizeof(uint32_t) * (nb_levels + 2));
            for (int i = 0; i < n; i++) if (lv[i]) start[lv[i]]++;            // counts at index level (1-based)
            uint32_t acc = 0;
            for (int l = 1; l <= nb_levels; l++) { uint32_t cnt = start[l]; start[l] = acc; acc += cnt; }
            start[nb_levels + 1] = acc;
            memcpy(fill, start, sizeof(uint32_t) * (nb_levels + 2));
            for (int i = 0; i < n; i++) if (lv[i]) L->order[p][fill[lv[i]]++] = (uint32_t)i;
            for (int l = 0; l <= nb_levels; l++) start[l] = start[l + 1];       // 0-based
            L->nb[p] = nb_levels;
            if (pos[p]) {
                if (!clean) memset(pos[p], 0, (size_t)w_u * h_u * sizeof(uint32_t));
                for (uint32_t q = 0; q < start[nb_levels]; q++) {
                    const OVGPUIntraCmd *c = &cmds[L->order[p][q]];
                    const int xu = c->x0 >> 2, yu = c->y0 >> 2, nw = (1 << c->log2_w) >> 2, nh = (1 << c->log2_h) >> 2;
                    for (int y = yu; y < yu + nh; y++) for (int x = xu; x < xu + nw; x++) pos[p][y * w_u + x] = q + 1;
                }
                // a block copy may only read units that no command covers or that a command EARLIER in the queue covers: the
                // dataflow kernel waits for them, and a wait on a later command (or on itself) would never end
                for (uint32_t q = 0; q < start[nb_levels] && !ret; q++) {
                    const OVGPUIntraCmd *c = &cmds[L->order[p][q]];
                    if (!(c->flags & OVGPU_INTRA_IBC)) continue;
                    int bvx = (int16_t)(c->pad[0] | (c->pad[1] << 8)), bvy = (int16_t)(c->pad[2] | (c->pad[3] << 8));
                    for (int y = (c->y0 + bvy) >> 2; y <= (c->y0 + bvy + (1 << c->log2_h) - 1) >> 2; y++)
                        for (int x = (c->x0 + bvx) >> 2; x <= (c->x0 + bvx + (1 << c->log2_w) - 1) >> 2; x++) if (pos[p][y * w_u + x] > q) ret = -2;
                }
                // the same for ordinary commands: the kernel waits on every neighbour unit the caller's availability words flag
                // (recon_kernels.h: left column from the corner down to 2h, above row over 2w).  The levels above come from geometry
                // and input order only, so a word that flags a unit covered by a LATER command of the queue (only an inconsistent
                // recorder produces one) would make that wait cyclic: refuse it here instead of hanging the GPU.
                const int um = L->units_mask;
                for (uint32_t q = 0; q < start[nb_levels] && !ret && um; q++) {
                    const OVGPUIntraCmd *c = &cmds[L->order[p][q]];
                    if (c->flags & OVGPU_INTRA_IBC) continue;
                    const int xu = c->x0 >> 2, yu = c->y0 >> 2, nw = (1 << c->log2_w) >> 2, nh = (1 << c->log2_h) >> 2;
                    const int xl = xu & um, yl = yu & um;
                    if (xu > 0)
                        for (int k = 0; k < 2 * nh + 1; k++) {
                            const int uy = yu - 1 + k, bit = yl + k;
                            if (uy < 0 || uy >= h_u || bit >= 64 || !((c->avail_cols >> bit) & 1)) continue;
                            if (pos[p][uy * w_u + xu - 1] > q) ret = -2;
                        }
                    if (yu > 0)
                        for (int k = 0; k < 2 * nw; k++) {
                            const int ux = xu + k, bit = xl + k + 1;
                            if (ux >= w_u || bit >= 64 || !((c->avail_rows >> bit) & 1)) continue;
                            if (pos[p][(yu - 1) * w_u + ux] > q) ret = -2;
                        }
                }
            }
        }
    }
    if (own) { free(fill); free(lvl); free(map1); }
    return ret;
}

// Launch sequence of a picture's intra predictions alone: the luma levels, then the chroma levels (Cb and Cr together).
// f(plane, first entry in order[plane], number of commands) -> 0 on success
template <typename F>
static inline int intra_for_each_launch(const IntraLevels *L, F f)
{
    for (int p = 0; p < 2; p++)
        for (int l = 0; l < L->nb[p]; l++) {
            const int start = (int)L->start[p][l], n = (int)L->start[p][l + 1] - start;
            int rc;
            if (n > 0 && (rc = f(p, start, n)) != 0) return rc;
        }
    return 0;
}

