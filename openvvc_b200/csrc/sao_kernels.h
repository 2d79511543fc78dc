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

    // edge offset with compile-time neighbour offsets (rcn_sao.c:79-117): idx = 2 + sign(c - a) + sign(c - b)
    template <int XA, int DYA>
    OVD void edge(const uint16_t *src, int sp, int x, int y, int w, bool full, const int *s, int *o,
                  uint32_t lut_lo, uint32_t lut_hi, int maxv)
    {
        const uint16_t *ra = src + (size_t)(y + DYA) * sp, *rb = src + (size_t)(y - DYA) * sp;
        const ovd_u4 va = ld_vec(ra + x, full), vb = ld_vec(rb + x, full);
        const uint32_t wa[4] = { va.x, va.y, va.z, va.w }, wb[4] = { vb.x, vb.y, vb.z, vb.w };
        int ta[10], tb[10];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            ta[1 + 2 * i] = wa[i] & 0xffff; ta[2 + 2 * i] = wa[i] >> 16;
            tb[1 + 2 * i] = wb[i] & 0xffff; tb[2 + 2 * i] = wb[i] >> 16;
        }
        ta[0] = ta[9] = tb[0] = tb[9] = 0;
        if (XA < 0) { ta[0] = ra[ovd_max(x - 1, 0)]; tb[9] = rb[ovd_min(x + 8, w - 1)]; }
        if (XA > 0) { ta[9] = ra[ovd_min(x + 8, w - 1)]; tb[0] = rb[ovd_max(x - 1, 0)]; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int na = ta[i + 1 + XA], nb = tb[i + 1 - XA];
            const int idx = 2 + sgn(s[i], na) + sgn(s[i], nb);
            const int off = (int)ovd_prmt(lut_lo, lut_hi, (uint32_t)idx * 0x1111u + 0x8880u);
            o[i] = ovd_clamp(s[i] + off, 0, maxv);
        }
        if (XA != 0) {                     // first / last column of the picture keeps its value
            if (x == 0) o[0] = s[0];
            const int il = w - 1 - x;
#pragma unroll
            for (int i = 0; i < 8; i++) if (i == il) o[i] = s[i];
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

        const bool full = (x + 8 <= w);                       // chroma widths are only multiples of 4
        ovd_u4 v = ld_vec(src + (size_t)y * sp + x, full);
        uint32_t in[4] = { v.x, v.y, v.z, v.w };
        int s[8];
#pragma unroll
        for (int i = 0; i < 4; i++) { s[2 * i] = in[i] & 0xffff; s[2 * i + 1] = in[i] >> 16; }
        int o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = s[i];

        // the 5 offsets (|offset| <= 31 at 10 bit, vcl_sao.c:69) as sign-extendable bytes of two words: one PRMT
        // with a "replicate sign" selector (idx * 0x1111 + 0x8880) fetches offset[idx] as an int
        const uint32_t lut_lo = (uint32_t)(uint8_t)p->offset_val[c][0] | ((uint32_t)(uint8_t)p->offset_val[c][1] << 8) |
                                ((uint32_t)(uint8_t)p->offset_val[c][2] << 16) | ((uint32_t)(uint8_t)p->offset_val[c][3] << 24);
        const uint32_t lut_hi = (uint32_t)(uint8_t)p->offset_val[c][4];
        if (type == 1) {
            // band offset: 4 consecutive bands starting at band_position (rcn_sao.c:65-71)
            const int bp = p->band_position[c];
            const int shift = a.bitdepth - 5;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = ((s[i] >> shift) - bp) & 31;
                const int off = (int)ovd_prmt(lut_lo, 0, (uint32_t)ovd_min(k, 4) * 0x1111u + 0x8880u);   // k >= 4 -> byte 4 = 0
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
                // neighbour a = (XA, DYA), b = (-XA, -DYA); one specialised instance per edge class
                if (eo == 0)      edge<-1, 0>(src, sp, x, y, w, full, s, o, lut_lo, lut_hi, maxv);
                else if (eo == 1) edge<0, -1>(src, sp, x, y, w, full, s, o, lut_lo, lut_hi, maxv);
                else if (eo == 2) edge<-1, -1>(src, sp, x, y, w, full, s, o, lut_lo, lut_hi, maxv);
                else              edge<1, -1>(src, sp, x, y, w, full, s, o, lut_lo, lut_hi, maxv);
            }
        }
        ovd_u4 r;
        r.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
        r.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
        r.z = (uint32_t)o[4] | ((uint32_t)o[5] << 16);
        r.w = (uint32_t)o[6] | ((uint32_t)o[7] << 16);
        uint16_t *d = (c == 0 ? a.dst[0] : (c == 1 ? a.dst[1] : a.dst[2])) + (size_t)y * (c == 0 ? a.dst_pitch[0] : (c == 1 ? a.dst_pitch[1] : a.dst_pitch[2])) + x;
        if (full) ovd_st16(d, r);
        else { ovd_u2 r2; r2.x = r.x; r2.y = r.y; ovd_st8(d, r2); }
    }
};
