// alf_host.h -- host-side packing of ALF state into the kernels' formats and launch geometry.
// Shared by the product (alf.cu, device pointers) and the test emulation (tests/emu, host pointers).
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/ovvc_b200.h"
#include "alf_kernels.h"

// Expand the RCNALF-format luma tables into kernel operands.  Returns 1 when some coefficient does not
// fit the int8 operand of IDP.2A (then the WIDE kernel variant must be used).
static inline int alf_pack_luma_tab(uint32_t *dst /*[24][100][24]*/, const OVGPUAlfTables *t)
{
    int wide = 0;
    for (int s = 0; s < OVGPU_ALF_NUM_SETS; s++)
        for (int e = 0; e < 100; e++)
            for (int k = 0; k < 12; k++) {
                int c = t->luma_coeff[s][e * 13 + k];
                if (c < -128 || c > 127) wide = 1;
            }
    for (int s = 0; s < OVGPU_ALF_NUM_SETS; s++)
        for (int e = 0; e < 100; e++) {
            uint32_t *w = dst + ((size_t)s * 100 + e) * ALF_TAB_ENTRY_WORDS;
            for (int k = 0; k < 12; k++) {
                int c = t->luma_coeff[s][e * 13 + k];
                int cl = t->luma_clip[s][e * 13 + k];
                // IDP.2A.LO uses operand bytes 0,1 and IDP.2A.HI bytes 2,3: [c,0,0,c] multiplies each half by c
                w[k] = wide ? (uint32_t)c : ((uint32_t)(uint8_t)(int8_t)c | ((uint32_t)(uint8_t)(int8_t)c << 24));
                // a clip of 1<<bitdepth never binds; keep it inside int16 (32767) so 8- and 10-bit share the path
                int clv = cl > 32767 ? 32767 : cl;
                w[12 + k] = (uint32_t)(uint16_t)clv * 0x10001u;
            }
        }
    return wide;
}

static inline void alf_pack_chroma_tab(AlfChromaTab *d, const OVGPUAlfTables *t)
{
    for (int a = 0; a < 8; a++)
        for (int k = 0; k < 6; k++) {
            d->coeff[a][k] = t->chroma_coeff[a][k];
            d->clip[a][k] = t->chroma_clip[a][k];
        }
    memcpy(d->cc, t->cc_coeff, sizeof(d->cc));
}

static inline void alf_pack_ctu(AlfCtuDev *d, int n, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2])
{
    for (int i = 0; i < n; i++) {
        d[i].flag = ctu[i].ctb_alf_flag;
        d[i].idx = ctu[i].ctb_alf_idx;
        d[i].cb_alt = ctu[i].cb_alternative;
        d[i].cr_alt = ctu[i].cr_alternative;
        d[i].cc_cb = (cc_idx && cc_idx[0]) ? cc_idx[0][i] : 0;
        d[i].cc_cr = (cc_idx && cc_idx[1]) ? cc_idx[1][i] : 0;
        d[i].pad0 = d[i].pad1 = 0;
    }
}

struct AlfGeom {
    int nb_ctu_w, nb_ctu_h;
    int l_tw, l_th, l_tx, l_ty;      // luma tile size, tiles per CTU
    int c_tw, c_th, c_tx, c_ty;      // chroma
};

static inline AlfGeom alf_geom(int w, int h, int log2_ctu)
{
    AlfGeom g;
    int ctu_s = 1 << log2_ctu;
    g.nb_ctu_w = (w + ctu_s - 1) >> log2_ctu;
    g.nb_ctu_h = (h + ctu_s - 1) >> log2_ctu;
    g.l_tw = ctu_s < 128 ? ctu_s : 128;
    g.l_th = ctu_s < 32 ? ctu_s : 32;
    g.l_tx = ctu_s / g.l_tw;
    g.l_ty = ctu_s / g.l_th;
    int ctu_c = ctu_s >> 1;
    g.c_tw = ctu_c < 64 ? ctu_c : 64;
    g.c_th = ctu_c < 32 ? ctu_c : 32;
    g.c_tx = ctu_c / g.c_tw;
    g.c_ty = ctu_c / g.c_th;
    return g;
}
