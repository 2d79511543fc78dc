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

/* ---------------------------------------------------------------------------------------------------
 * Film grain synthesis: fg_grain_apply_pic and its helpers (libovvc/pp_film_grain.c:569-978), the output-stage step
 * pp_process_frame runs when the stream carries a film grain characteristics SEI (post_proc.c:50-57, 100-103; always
 * with de-blocking on, bit depth 10, isIdrPic 0).  SMPTE RDD 5 frequency-filtering model:
 *   - a data base of 13 x 13 grain patterns of 64 x 64 samples (horizontal x vertical cut-off frequency), each the
 *     64-point integer inverse DCT of a band-limited block of Gaussian samples, de-blocked at its 8-row boundaries;
 *   - per 16x16 block of a component one step of a 32-bit pseudo random generator seeded per picture and component;
 *     per 8x8 block the average of the decoded samples selects an intensity interval = (scale, cut-off pair), the
 *     generator's value the 8x8 window of the pattern and the sign; the grain is scale * pattern >> (log2 scale + 6);
 *   - the grain stripe of 16 rows is smoothed across the vertical 8x8 block edges, scaled to the bit depth and added.
 * ------------------------------------------------------------------------------------------------- */
#include <string.h>
#include "fg_tables.h"

static uint32_t fg_prng(uint32_t x) { return (x << 1) + ((1 + ((x & (1u << 2)) > 0) + ((x & (1u << 30)) > 0)) % 2); }   /* :635-641 */

void ovo_fg_database(int8_t *db /* [13][13][64][64] */, int deblock)
{   /* fg_data_base_generation, :666-765 */
    static int32_t B[64][64], T[64][64];
    for (int h = 0; h < 13; h++)
        for (int v = 0; v < 13; v++) {
            memset(B, 0, sizeof(B));
            const unsigned fh = ((h + 3) << 2) - 1, fv = ((v + 3) << 2) - 1;
            uint32_t e = fg_seeds[h + v * 13];
            for (unsigned l = 0; l <= fv; l++)
                for (unsigned k = 0; k <= fh; k += 4) {
                    for (int i = 0; i < 4; i++) B[k + i][l] = fg_gaussian[(e + i) % 2048];
                    e = fg_prng(e);
                }
            B[0][0] = 0;
            for (int i = 0; i < 64; i++)
                for (int j = 0; j < 64; j++) {
                    int32_t s = 0;
                    for (int k = 0; k < 64; k++) s += fg_idct64[k * 64 + i] * B[k][j];
                    T[i][j] = (s + 128) >> 8;
                }
            int8_t *pat = db + (size_t)(h * 13 + v) * 4096;
            for (int j = 0; j < 64; j++)
                for (int i = 0; i < 64; i++) {
                    int32_t s = 0;
                    for (int k = 0; k < 64; k++) s += T[i][k] * fg_idct64[k * 64 + j];
                    s = (s + 128) >> 8;
                    pat[j * 64 + i] = (int8_t)(s < -127 ? -127 : s > 127 ? 127 : s);            /* ov_clip_intp2(., 8): symmetric */
                }
            if (deblock)
                for (int l = 0; l < 64; l += 8)
                    for (int k = 0; k < 64; k++) {
                        pat[l * 64 + k] = (int8_t)((pat[l * 64 + k] * (int)fg_deblock_factor[v]) >> 7);
                        pat[(l + 7) * 64 + k] = (int8_t)((pat[(l + 7) * 64 + k] * (int)fg_deblock_factor[v]) >> 7);
                    }
        }
}

/* every dimension must keep the reference inside its planes: its block loops compute min(8, size - position) in unsigned
 * arithmetic, which wraps when a 16-sample block row / column has fewer than 8 samples left */
int ovo_fg_size_ok(int w, int h)
{
    for (int c = 0; c < 2; c++) {
        const int cw = w >> c, chh = h >> c;
        if ((cw % 16 && cw % 16 < 8) || (chh % 16 && chh % 16 < 8)) return 0;
    }
    return 1;
}

int ovo_fg_apply(uint16_t *const dst[3], const uint16_t *const src[3], int w, int h, const OVOFilmGrain *fg, int32_t poc, const int8_t *db)
{   /* fg_grain_apply_pic, :814-977, planes of stride = width */
    static const uint8_t colour_offset[3] = { 0, 85, 170 };
    if (!ovo_fg_size_ok(w, h)) return -1;
    int32_t *stripe = malloc((size_t)w * 16 * sizeof(int32_t));
    for (int c = 0; c < 3; c++) {
        const int cw = c ? w >> 1 : w, chh = c ? h >> 1 : h;
        if (!fg->comp_model_present_flag[c]) { memcpy(dst[c], src[c], (size_t)cw * chh * 2); continue; }
        /* fg_compute_model_values, :768-807: interval of every 8-bit average, default / chroma-converted model values */
        int16_t interval[256], mv[8][3];
        memset(interval, -1, sizeof(interval));
        for (int i = 0; i < 8; i++) {
            for (int a = fg->intensity_interval_lower_bound[c][i]; a <= fg->intensity_interval_upper_bound[c][i]; a++) interval[a] = (int16_t)i;
            mv[i][0] = fg->comp_model_value[c][i][0]; mv[i][1] = fg->comp_model_value[c][i][1]; mv[i][2] = fg->comp_model_value[c][i][2];
            if (fg->num_model_values_minus1[c] + 1 == 1) { mv[i][1] = 8; mv[i][2] = 8; }
            else if (fg->num_model_values_minus1[c] + 1 == 2) mv[i][2] = mv[i][1];
            if (c) {
                mv[i][0] >>= 1;
                for (int k = 1; k < 3; k++) { int t = mv[i][k] << 1; mv[i][k] = (int16_t)(t < 2 ? 2 : t > 14 ? 14 : t); }
            }
        }
        uint32_t e = fg_seeds[((uint32_t)poc + colour_offset[c]) % 256];
        for (int y = 0; y < chh; y += 16) {
            memset(stripe, 0, (size_t)w * 16 * sizeof(int32_t));
            for (int x = 0; x < cw; x += 16) {
                for (int b = 0; b < 4; b++) {
                    const int yo = (b >> 1) * 8, xo = (b & 1) * 8;
                    const int ys = chh - y - yo < 8 ? chh - y - yo : 8, xs = cw - x - xo < 8 ? cw - x - xo : 8;
                    if (ys <= 0 || xs <= 0) continue;
                    uint32_t sum = 0;
                    for (int k = 0; k < ys; k++) for (int l = 0; l < xs; l++) sum += src[c][(size_t)(y + yo + k) * cw + x + xo + l];
                    uint32_t avg = (sum / (uint32_t)(ys * xs)) >> 2;
                    if (avg > 255) avg = 255;
                    const int it = interval[avg];
                    if (it < 0) continue;
                    const int ko = (int)(((e >> 16) % 52) & 0xfffc) + ((x + xo) & 8), lo = (int)(((e & 0xffff) % 56) & 0xfff8) + ((y + yo) & 8);
                    const int scale = (e & 1) ? -mv[it][0] : mv[it][0], hh = mv[it][1] - 2, vv = mv[it][2] - 2;
                    const int8_t *pat = db + (size_t)(hh * 13 + vv) * 4096;
                    for (int l = 0; l < 8; l++)
                        for (int k = 0; k < xs; k++)
                            stripe[(yo + l) * cw + x + xo + k] = (scale * pat[(l + lo) * 64 + k + ko]) >> (fg->log2_scale_factor + 6);
                }
                e = fg_prng(e);
            }
            for (int p = 0; p + 8 < cw; p += 8)                 /* fg_deblock_grain_stripe, :598-616 */
                for (int r = 0; r < 16; r++) {
                    int32_t *g = stripe + r * cw + p;
                    const int32_t l1 = g[6], l0 = g[7], r0 = g[8], r1 = g[9];
                    g[8] = (l0 + (r0 << 1) + r1) >> 2;
                    g[7] = (l1 + (l0 << 1) + r0) >> 2;
                }
            const int rows = chh - y < 16 ? chh - y : 16;
            for (int r = 0; r < rows; r++)
                for (int k = 0; k < cw; k++) {
                    const int v = (int)src[c][(size_t)(y + r) * cw + k] + (stripe[r * cw + k] << 2);      /* fg_blend_stripe at 10 bit */
                    dst[c][(size_t)(y + r) * cw + k] = (uint16_t)(v < 0 ? 0 : v > 1023 ? 1023 : v);
                }
        }
    }
    free(stripe);
    return 0;
}
