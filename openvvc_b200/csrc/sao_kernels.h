// sao_kernels.h -- whole-picture SAO kernel (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: libovvc/rcn_sao.c (band :46-76, edge :79-117, border rules :162-177, CTU-line
// drivers :190-293).  One thread filters 8 horizontally adjacent samples (one 16-byte vector) of one
// component; src = deblocked picture, dst = SAO output, so neighbours are never read after modification
// (the reference needs its padded filter-region copy for that).
#pragma once
#include "dev.h"
#include "tile_grid.h"

struct SaoCtuDev {                // == OVGPUSaoCtu (40 bytes)
    int16_t offset_val[3][5];
    uint8_t band_position[3];
    uint8_t eo_class[3];
    uint8_t type_idx[3];
    uint8_t pad;
};

struct SaoK {
#ifndef SAO_MIN_BLOCKS
#define SAO_MIN_BLOCKS 6             /* 40 registers: the kernel waits on memory, not on the issue slots (4: 24.8 us, 6: 22.8 us per 4K picture) */
#endif
    static const int THREADS = 256, MIN_BLOCKS = SAO_MIN_BLOCKS, PHASES = 1;
    static const int BX = 32, BY = 8;      // 32 vectors (256 samples) x 8 rows per CTA

    struct Args {
        const uint16_t *src[3];
        uint16_t *dst[3];
        int src_pitch[3], dst_pitch[3];    // samples
        int pic_w, pic_h, log2_ctu, bitdepth, nb_ctu_w, nb_ctu_h;
        const SaoCtuDev *ctu;
        const TileGridDev *tiles;          // device memory, or NULL
    };
    struct State { int dummy; };

    OVD ovd_u4 ld_vec(const uint16_t *p, bool full)
    {
        if (full) return ovd_ld16(p);
        ovd_u2 h = ovd_ld8(p);
        ovd_u4 r; r.x = h.x; r.y = h.y; r.z = 0; r.w = 0;
        return r;
    }

    // All arithmetic runs on PACKED PAIRS of 16-bit samples (two samples per instruction):
    //   sign(n - c)          = max(min(n + (-c), 1), -1)                          VIADDMNMX.S16x2 + VIMNMX.S16x2
    //   offset look-up       = one PRMT per pair: the five offsets are sign-extendable bytes of two words and the selector
    //                          (idx_lo, idx_lo | 8, idx_hi, idx_hi | 8) fetches both as int16 halves
    //   clip(c + offset)     = max(min(c + off, maxv), 0)                         VIADDMNMX.S16x2.RELU
    // selector of the pair of indices W = idx_lo | idx_hi << 16 (each 0..4)
    OVD uint32_t sel2(uint32_t W) { const uint32_t t = W * 0x11u; return t | (t >> 8) | 0x8080u; }
    // pair of samples starting one sample to the LEFT of word `cur`: (hi half of prev, lo half of cur)
    OVD uint32_t odd_pair(uint32_t prev, uint32_t cur) { return ovd_byte_perm(prev, cur, 0x5432); }

    // edge offset with compile-time neighbour offsets (rcn_sao.c:79-117): idx = 2 + sign(c - a) + sign(c - b),
    // neighbour a = (XA, DYA), b = (-XA, -DYA)
    template <int XA, int DYA>
    OVD void edge(const uint16_t *src, int sp, int x, int y, int w, bool full, const uint32_t *c, uint32_t *o,
                  uint32_t lut_lo, uint32_t lut_hi, uint32_t maxv2, int x_in_tile, int tile_w)
    {
        const uint16_t *ra = src + (size_t)(y + DYA) * sp, *rb = src + (size_t)(y - DYA) * sp;
        uint32_t wa[4], wb[4];
        if (DYA == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) wa[i] = wb[i] = c[i];
        } else {
            const ovd_u4 va = ld_vec(ra + x, full), vb = ld_vec(rb + x, full);
            wa[0] = va.x; wa[1] = va.y; wa[2] = va.z; wa[3] = va.w;
            wb[0] = vb.x; wb[1] = vb.y; wb[2] = vb.z; wb[3] = vb.w;
        }
        uint32_t na[4], nb[4];
        if (XA == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) { na[i] = wa[i]; nb[i] = wb[i]; }
        } else {
            // the pairs that start one sample left (XA < 0: neighbour a) or right of the thread's pairs
            const uint16_t *rl = XA < 0 ? ra : rb, *rr = XA < 0 ? rb : ra;
            const uint32_t *wl = XA < 0 ? wa : wb, *wr = XA < 0 ? wb : wa;
            const uint32_t left = (uint32_t)rl[ovd_max(x - 1, 0)] << 16, right = rr[ovd_min(x + 8, w - 1)];
            uint32_t L[4], R[4];
            L[0] = odd_pair(left, wl[0]);
#pragma unroll
            for (int i = 1; i < 4; i++) L[i] = odd_pair(wl[i - 1], wl[i]);
#pragma unroll
            for (int i = 0; i < 3; i++) R[i] = odd_pair(wr[i], wr[i + 1]);
            R[3] = odd_pair(wr[3], right);
#pragma unroll
            for (int i = 0; i < 4; i++) { na[i] = XA < 0 ? L[i] : R[i]; nb[i] = XA < 0 ? R[i] : L[i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t negc = ovd_vsub_16x2(0u, c[i]);
            const uint32_t sa = ovd_vmax_s16x2(ovd_viaddmin_s16x2(na[i], negc, 0x00010001u), 0xffffffffu);      // sign(a - c)
            const uint32_t sb = ovd_vmax_s16x2(ovd_viaddmin_s16x2(nb[i], negc, 0x00010001u), 0xffffffffu);
            const uint32_t idx = ovd_vsub_16x2(0x00020002u, ovd_vadd_16x2(sa, sb));                              // 2 + sign(c - a) + sign(c - b)
            o[i] = ovd_viaddmin_s16x2_relu(c[i], ovd_prmt(lut_lo, lut_hi, sel2(idx)), maxv2);
        }
        if (XA != 0) {                     // first / last column of the picture (of the tile) keeps its value
            const int il = tile_w - 1 - x_in_tile;             // position of the tile's last column among the thread's 8 samples
            const uint32_t m = (il & 1) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
            for (int i = 0; i < 4; i++) {                       // (selects on registers: no run-time index into o[] / c[])
                uint32_t keep = il == 2 * i || il == 2 * i + 1 ? m : 0u;
                if (i == 0 && x_in_tile == 0) keep |= 0x0000ffffu;
                o[i] = (o[i] & ~keep) | (c[i] & keep);
            }
        }
    }

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st;
        const int c = bz, sh = c ? 1 : 0;
        const int w = a.pic_w >> sh, h = a.pic_h >> sh;
        const int x = (bx * BX + (tid & 31)) * 8;
        const int y = by * BY + (tid >> 5);
        if (x >= w || y >= h) return;
        // (selects, not a.src[c]: dynamic indexing would copy the parameter block to local memory)
        const uint16_t *src = c == 0 ? a.src[0] : (c == 1 ? a.src[1] : a.src[2]);
        const int sp = c == 0 ? a.src_pitch[0] : (c == 1 ? a.src_pitch[1] : a.src_pitch[2]);
        const int lc = a.log2_ctu - sh;
        const SaoCtuDev *p = &a.ctu[(y >> lc) * a.nb_ctu_w + (x >> lc)];
        const int type = p->type_idx[c];
        const int maxv = (1 << a.bitdepth) - 1;

        // the tile of the thread's samples (8 samples never straddle a CTU): its edges count as picture borders
        int tx0 = 0, tx1 = w, ty0 = 0, ty1 = h, tile_rows = a.nb_ctu_h;
        if (a.tiles) {
            int c0, c1;
            tile_span(a.tiles->col_bd, a.tiles->nb_cols, x >> lc, a.nb_ctu_w, c0, c1);
            tx0 = c0 << lc; tx1 = ovd_min(c1 << lc, w);
            tile_span(a.tiles->row_bd, a.tiles->nb_rows, y >> lc, a.nb_ctu_h, c0, c1);
            ty0 = c0 << lc; ty1 = ovd_min(c1 << lc, h); tile_rows = c1 - c0;
        }
        const bool full = (x + 8 <= w);                       // chroma widths are only multiples of 4
        ovd_u4 v = ld_vec(src + (size_t)y * sp + x, full);
        const uint32_t in[4] = { v.x, v.y, v.z, v.w };
        uint32_t o[4] = { v.x, v.y, v.z, v.w };
        const uint32_t maxv2 = (uint32_t)maxv * 0x10001u;

        // the 5 offsets (|offset| <= 31 at 10 bit, vcl_sao.c:69) as sign-extendable bytes of two words
        const uint32_t lut_lo = (uint32_t)(uint8_t)p->offset_val[c][0] | ((uint32_t)(uint8_t)p->offset_val[c][1] << 8) |
                                ((uint32_t)(uint8_t)p->offset_val[c][2] << 16) | ((uint32_t)(uint8_t)p->offset_val[c][3] << 24);
        const uint32_t lut_hi = (uint32_t)(uint8_t)p->offset_val[c][4];
        if (type == 1) {
            // band offset: 4 consecutive bands starting at band_position (rcn_sao.c:65-71); k >= 4 -> byte 4 = 0
            const uint32_t bp2 = (uint32_t)p->band_position[c] * 0x10001u;
            const int shift = a.bitdepth - 5;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t band = (in[i] >> shift) & 0x001f001fu;
                const uint32_t k = ovd_vmin_s16x2(ovd_vsub_16x2(band, bp2) & 0x001f001fu, 0x00040004u);
                o[i] = ovd_viaddmin_s16x2_relu(in[i], ovd_prmt(lut_lo, 0, sel2(k)), maxv2);
            }
        } else if (type == 2) {
            const int eo = p->eo_class[c];
            // rows/cols on the picture border are left untouched when a neighbour would be missing
            // (rcn_sao.c:162-177); single-CTU-row quirk of rcn_sao_first_pix_rows (rcn_sao.c:259-272)
            bool row_ok = true;
            if (eo != 0) {
                if (y == ty0 || y == ty1 - 1) row_ok = false;
                if (tile_rows == 1 && y == ty0 + (6 >> sh) - 1) row_ok = false;
            }
            if (row_ok) {
                // one specialised instance per edge class; columns relative to the tile: its first / last one keep their samples
                if (eo == 0)      edge<-1, 0>(src, sp, x, y, w, full, in, o, lut_lo, lut_hi, maxv2, x - tx0, tx1 - tx0);
                else if (eo == 1) edge<0, -1>(src, sp, x, y, w, full, in, o, lut_lo, lut_hi, maxv2, x - tx0, tx1 - tx0);
                else if (eo == 2) edge<-1, -1>(src, sp, x, y, w, full, in, o, lut_lo, lut_hi, maxv2, x - tx0, tx1 - tx0);
                else              edge<1, -1>(src, sp, x, y, w, full, in, o, lut_lo, lut_hi, maxv2, x - tx0, tx1 - tx0);
            }
        }
        ovd_u4 r;
        r.x = o[0]; r.y = o[1]; r.z = o[2]; r.w = o[3];
        uint16_t *d = (c == 0 ? a.dst[0] : (c == 1 ? a.dst[1] : a.dst[2])) + (size_t)y * (c == 0 ? a.dst_pitch[0] : (c == 1 ? a.dst_pitch[1] : a.dst_pitch[2])) + x;
        if (full) ovd_st16(d, r);
        else { ovd_u2 r2; r2.x = r.x; r2.y = r.y; ovd_st8(d, r2); }
    }
};
