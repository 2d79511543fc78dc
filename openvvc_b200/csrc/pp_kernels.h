// pp_kernels.h -- output stage: a decoded picture up-scaled to the sequence's maximum size (sm_100a; also compiled by tests/emu).
//
// Reference behaviour: pp_sample_rate_conv (libovvc/pp_pic_scale.c:252-377) as pp_process_frame applies it to the output frames of a
// sequence with reference picture resampling (libovvc/post_proc.c:117-128, sei->upscale_flag): position of output sample i =
// i * scale + add in 2^-13 (luma) / 2^-14 (chroma) source samples, integer part = position >> scale_bits, filter phase = the LOW
// four (five) bits of the position (as the reference has it), regular 8-tap / 4-tap interpolation filters in both directions with
// 32-bit intermediates and one rounding by 2^12.  Both passes are linear and exact, so every output sample is computed directly
// as its 8x8 (4x4) weighted sum: one thread per sample, the source (smaller than the target) stays in L1 / L2.
#pragma once
#include "dev.h"
#include "mc_rpr_tables.h"

struct UpscaleK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 1;
    struct Args {
        uint16_t *dst[3]; const uint16_t *src[3];
        int dst_pitch[3], src_pitch[3];                  // samples
        int dst_w, dst_h, src_w, src_h;                  // luma sizes
        int scale_hor[2], scale_ver[2], add_x, add_y;    // [0] luma (2^13), [1] chroma (2^14); chroma siting offsets
        int maxv;
    };
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st;
        const int c = bz, ch = c != 0, x = bx * 32 + (tid & 31), y = by * 8 + (tid >> 5);
        const int dw = a.dst_w >> ch, dh = a.dst_h >> ch, sw = a.src_w >> ch, sh = a.src_h >> ch;
        if (x >= dw || y >= dh) return;
        const int bits = ch ? 14 : 13, nt = ch ? 4 : 8, pmask = ch ? 0x1f : 0xf;
        const int px = x * a.scale_hor[ch] + (ch ? a.add_x : 0), py = y * a.scale_ver[ch] + (ch ? a.add_y : 0);
        const int xi = (px >> bits) - nt / 2 + 1, yi = (py >> bits) - nt / 2 + 1;
        const int8_t *fx = ch ? mc_rpr_c[0][px & pmask] : mc_rpr_l[0][px & pmask];
        const int8_t *fy = ch ? mc_rpr_c[0][py & pmask] : mc_rpr_l[0][py & pmask];
        int sum = 0;
        for (int j = 0; j < nt; j++) {
            const uint16_t *row = a.src[c] + (size_t)ovd_clamp(yi + j, 0, sh - 1) * a.src_pitch[c];
            int h = 0;
            for (int k = 0; k < nt; k++) h += fx[k] * (int)row[ovd_clamp(xi + k, 0, sw - 1)];
            sum += fy[j] * h;
        }
        a.dst[c][(size_t)y * a.dst_pitch[c] + x] = (uint16_t)ovd_clamp((sum + (1 << 11)) >> 12, 0, a.maxv);
    }
};

// scale factors and siting offsets as pp_sample_rate_conv derives them (pp_pic_scale.c:256-277); returns -1 for a target smaller
// than the source window (the down-sampling branch, never reached from post_proc.c)
static inline int upscale_setup(UpscaleK::Args *a, int win_l, int win_r, int win_t, int win_b, int hor_col, int ver_col)
{
    const int ew = (win_l + win_r) << 1, eh = (win_t + win_b) << 1;
    a->scale_hor[0] = ((a->src_w - (uint16_t)(ew << 1)) << 13) / a->dst_w;
    a->scale_ver[0] = ((a->src_h - (uint16_t)(eh << 1)) << 13) / a->dst_h;
    a->scale_hor[1] = (((a->src_w >> 1) - (uint16_t)ew) << 14) / (a->dst_w >> 1);
    a->scale_ver[1] = (((a->src_h >> 1) - (uint16_t)eh) << 14) / (a->dst_h >> 1);
    if (a->scale_hor[0] > (1 << 13) || a->scale_ver[0] > (1 << 13) || a->scale_hor[1] > (1 << 14) || a->scale_ver[1] > (1 << 14)) return -1;
    if (a->scale_hor[0] <= 0 || a->scale_ver[0] <= 0 || a->scale_hor[1] <= 0 || a->scale_ver[1] <= 0) return -1;
    a->add_x = ((1 - hor_col) * 8 * (a->scale_hor[1] - (1 << 14)) + (1 << 13)) >> 14;
    a->add_y = ((1 - ver_col) * 8 * (a->scale_ver[1] - (1 << 14)) + (1 << 13)) >> 14;
    return 0;
}
