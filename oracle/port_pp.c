/* oracle/port_pp.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the output stage's picture up-scaler: pp_sample_rate_conv (libovvc/pp_pic_scale.c:252-377), which
 * pp_process_frame (libovvc/post_proc.c:76-139) applies to every output frame of a sequence with reference picture
 * resampling when the application asked for full-size output (sei->upscale_flag): each plane of the decoded frame is
 * resampled to the sequence's maximum size with the regular interpolation filters (8 taps luma, 4 taps chroma), both
 * passes in 32-bit, one rounding by 2^12 at the end.
 *
 * Kept as the reference computes it, including what looks unintended: the position of output sample i is
 * i * scale + add in 2^-13 (luma) / 2^-14 (chroma) source samples and the filter phase is taken from the LOW four
 * (five) bits of that position (:318-319, :343-344), not from its top fractional bits.  Only the up-sampling branch is
 * restated (the 12-tap down-sampling filters are never reached from post_proc.c: the target is the maximum size).
 */
#include "oracle.h"
#include <stdlib.h>
#include "mc_rpr_tables.h"

int ovo_pp_upscale_plane(uint16_t *dst, int dst_stride, int dst_w, int dst_h, const uint16_t *src, int src_stride, int src_w, int src_h,
                         const OVOScalingWindow *win, int luma, int bitdepth)
{
    int extra_w = (win->left + win->right) << 1, extra_h = (win->top + win->bottom) << 1;
    if (luma) { extra_w <<= 1; extra_h <<= 1; }
    const int scale_bits = luma ? 13 : 14;
    const int scale_hor = ((src_w - (uint16_t)extra_w) << scale_bits) / dst_w, scale_ver = ((src_h - (uint16_t)extra_h) << scale_bits) / dst_h;
    if (scale_hor > (1 << scale_bits) || scale_ver > (1 << scale_bits)) return -1;      /* down-sampling: not restated */
    int add_x = 0, add_y = 0;
    if (!luma) {
        add_x = (1 - win->chroma_hor_col) * 8 * (scale_hor - (1 << scale_bits));
        add_x = (add_x + (1 << (scale_bits - 1))) >> scale_bits;
        add_y = (1 - win->chroma_ver_col) * 8 * (scale_ver - (1 << scale_bits));
        add_y = (add_y + (1 << (scale_bits - 1))) >> scale_bits;
    }
    const int nt = luma ? 8 : 4, pmask = luma ? 0xF : 0x1F, maxv = (1 << bitdepth) - 1;
    int *buf = malloc((size_t)dst_w * src_h * sizeof(int));
    for (int i = 0; i < dst_w; i++) {
        const int ref_pos = i * scale_hor + add_x, xi = ref_pos >> scale_bits, prec = ref_pos & pmask;
        const int8_t *f = luma ? mc_rpr_l[0][prec] : mc_rpr_c[0][prec];
        for (int j = 0; j < src_h; j++) {
            int sum = 0;
            for (int k = 0; k < nt; k++) {
                int x = xi + k - nt / 2 + 1;
                x = x < 0 ? 0 : x > src_w - 1 ? src_w - 1 : x;
                sum += f[k] * src[(size_t)j * src_stride + x];
            }
            buf[(size_t)j * dst_w + i] = sum;
        }
    }
    for (int j = 0; j < dst_h; j++) {
        const int ref_pos = j * scale_ver + add_y, yi = ref_pos >> scale_bits, prec = ref_pos & pmask;
        const int8_t *f = luma ? mc_rpr_l[0][prec] : mc_rpr_c[0][prec];
        for (int i = 0; i < dst_w; i++) {
            int sum = 0;
            for (int k = 0; k < nt; k++) {
                int y = yi + k - nt / 2 + 1;
                y = y < 0 ? 0 : y > src_h - 1 ? src_h - 1 : y;
                sum += f[k] * buf[(size_t)y * dst_w + i];
            }
            int v = (sum + (1 << 11)) >> 12;
            dst[(size_t)j * dst_stride + i] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
        }
    }
    free(buf);
    return 0;
}
