// sao_kernels.h -- whole-picture SAO kernel (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: libovvc/rcn_sao.c (band :46-76, edge :79-117, border rules :162-177, CTU-line
// drivers :190-293).  One thread filters 8 horizontally adjacent samples (one 16-byte vector) of one
// component; src = deblocked picture, dst = SAO output, so neighbours are never read after modification
// (the reference needs its padded filter-region copy for that).
#pragma once
#include "dev.h"

struct SaoCtuDev {                // == OVGPUSaoCtu (40 bytes)
    int16_t offset_val[3][5];
    uint8_t band_position[3];
    uint8_t eo_class[3];
    uint8_t type_idx[3];
    uint8_t pad;
};

struct SaoK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 1;
    static const int BX = 32, BY = 8;      // 32 vectors (256 samples) x 8 rows per CTA

    struct Args {
        const uint16_t *src[3];
        uint16_t *dst[3];
        int src_pitch[3], dst_pitch[3];    // samples
        int pic_w, pic_h, log2_ctu, bitdepth, nb_ctu_w, nb_ctu_h;
        const SaoCtuDev *ctu;
    };
    struct State { int dummy; };

    OVD int sgn(int a, int b) { return (a > b) - (a < b); }
    OVD ovd_u4 ld_vec(const uint16_t *p, bool full)
    {
        if (full) return ovd_ld16(p);
        ovd_u2 h = ovd_ld8(p);
        ovd_u4 r; r.x = h.x; r.y = h.y; r.z = 0; r.w = 0;
        return r;
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
        const uint16_t *src = a.src[c];
        const int sp = a.src_pitch[c];
        const int lc = a.log2_ctu - sh;
        const SaoCtuDev *p = &a.ctu[(y >> lc) * a.nb_ctu_w + (x >> lc)];
        const int type = p->type_idx[c];
        const int maxv = (1 << a.bitdepth) - 1;

        const bool full = (x + 8 <= w);                       // chroma widths are only multiples of 4
        ovd_u4 v = ld_vec(src + (size_t)y * sp + x, full);
        uint32_t in[4] = { v.x, v.y, v.z, v.w };
        int s[8];
#pragma unroll
        for (int i = 0; i < 4; i++) { s[2 * i] = in[i] & 0xffff; s[2 * i + 1] = in[i] >> 16; }
        int o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = s[i];

        if (type == 1) {
            // band offset: 4 consecutive bands starting at band_position (rcn_sao.c:65-71)
            const int bp = p->band_position[c];
            const int shift = a.bitdepth - 5;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int k = ((s[i] >> shift) - bp) & 31;
                int off = k < 4 ? p->offset_val[c][k] : 0;
                o[i] = ovd_clamp(s[i] + off, 0, maxv);
            }
        } else if (type == 2) {
            const int eo = p->eo_class[c];
            // rows/cols on the picture border are left untouched when a neighbour would be missing
            // (rcn_sao.c:162-177); single-CTU-row quirk of rcn_sao_first_pix_rows (rcn_sao.c:259-272)
            bool row_ok = true;
            if (eo != 0) {
                if (y == 0 || y == h - 1) row_ok = false;
                if (a.nb_ctu_h == 1 && y == (6 >> sh) - 1) row_ok = false;
            }
            if (row_ok) {
                const int dx = (eo == 1) ? 0 : 1;
                const int dya = (eo == 0) ? 0 : -1;                 // neighbour a: (-dx or +dx for 135 deg, dya)
                const int xa = (eo == 3) ? dx : -dx;                // a = (xa, dya), b = (-xa, -dya)
                const uint16_t *ra = src + (size_t)(y + dya) * sp;
                const uint16_t *rb = src + (size_t)(y - dya) * sp;
                int na[8], nb[8];
                // neighbour rows as vectors, shifted by xa / -xa
                ovd_u4 va = ld_vec(ra + x, full), vb = ld_vec(rb + x, full);
                uint32_t wa[4] = { va.x, va.y, va.z, va.w }, wb[4] = { vb.x, vb.y, vb.z, vb.w };
                int ta[10], tb[10];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ta[1 + 2 * i] = wa[i] & 0xffff; ta[2 + 2 * i] = wa[i] >> 16;
                    tb[1 + 2 * i] = wb[i] & 0xffff; tb[2 + 2 * i] = wb[i] >> 16;
                }
                ta[0] = ta[9] = tb[0] = tb[9] = 0;
                if (dx) {
                    ta[0] = ra[ovd_max(x - 1, 0)]; ta[9] = ra[ovd_min(x + 8, w - 1)];
                    tb[0] = rb[ovd_max(x - 1, 0)]; tb[9] = rb[ovd_min(x + 8, w - 1)];
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {      // selects, not dynamic indexing: keeps everything in registers
                    na[i] = xa < 0 ? ta[i] : (xa > 0 ? ta[i + 2] : ta[i + 1]);
                    nb[i] = xa < 0 ? tb[i + 2] : (xa > 0 ? tb[i] : tb[i + 1]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    bool ok = true;
                    if (eo != 1) ok = (x + i != 0) && (x + i != w - 1);
                    if (ok) {
                        int idx = 2 + sgn(s[i], na[i]) + sgn(s[i], nb[i]);
                        o[i] = ovd_clamp(s[i] + p->offset_val[c][idx], 0, maxv);
                    }
                }
            }
        }
        ovd_u4 r;
        r.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
        r.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
        r.z = (uint32_t)o[4] | ((uint32_t)o[5] << 16);
        r.w = (uint32_t)o[6] | ((uint32_t)o[7] << 16);
        uint16_t *d = a.dst[c] + (size_t)y * a.dst_pitch[c] + x;
        if (full) ovd_st16(d, r);
        else { ovd_u2 r2; r2.x = r.x; r2.y = r.y; ovd_st8(d, r2); }
    }
};
