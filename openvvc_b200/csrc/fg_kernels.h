// fg_kernels.h -- output stage: film grain synthesis on the decoded picture (sm_100a; also compiled by tests/emu, see dev.h).
//
// Reference behaviour: fg_grain_apply_pic and its helpers (libovvc/pp_film_grain.c:569-978) as pp_process_frame runs them when
// the stream carries a film grain characteristics SEI (libovvc/post_proc.c:50-57, 100-103: de-blocking on, 10 bit).  SMPTE RDD 5
// frequency-filtering model: per 16x16 block of a component one step of a 32-bit pseudo random generator (seeded per picture
// and component), per 8x8 block the average of the decoded samples picks an intensity interval = (scale, horizontal / vertical
// cut-off) and the generator's value an 8x8 window of one of 13 x 13 pre-computed 64x64 grain patterns and a sign; the grain
// is smoothed across the vertical 8x8 edges of its 16-row stripe and added to the samples.
//
// The reference walks the picture serially (the generator advances block by block, the stripe is de-blocked in place).  Here:
//   FgSeedK      one thread per row of 16x16 blocks steps the generator to its row and writes the row's states: the only serial
//                part, 4 instructions per step;
//   FilmGrainK   one CTA per 16 rows x 128 columns: 8-sample partial sums (phase 0), the parameters of its 2 x 16 8x8 blocks and
//                of the block column either side (phase 1), then every thread builds its 8 grain samples straight from the
//                pattern, takes the two samples it needs from the neighbouring blocks for the edge smoothing (the de-blocking
//                reads only ORIGINAL grain values: every edge is independent of the others), blends and stores (phase 2).
#pragma once
#include <stdlib.h>
#include "dev.h"
#include "fg_tables.h"

#define FG_DB_BYTES (13 * 13 * 64 * 64)

OVHD uint32_t fg_prng(uint32_t x) { return (x << 1) + ((1 + ((x & (1u << 2)) > 0) + ((x & (1u << 30)) > 0)) % 2); }      // pp_film_grain.c:635-641

// the pattern data base (fg_data_base_generation, :666-765), built once on the host
static inline void fg_build_database(int8_t *db)
{
    int32_t (*B)[64] = (int32_t (*)[64])malloc(2 * 64 * 64 * sizeof(int32_t)), (*T)[64] = B + 64;      // per call: thread-safe
    for (int h = 0; h < 13; h++)
        for (int v = 0; v < 13; v++) {
            memset(B, 0, 64 * 64 * sizeof(int32_t));
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
                    pat[j * 64 + i] = (int8_t)(s < -127 ? -127 : s > 127 ? 127 : s);
                }
            for (int l = 0; l < 64; l += 8)          // de-blocking of the pattern's 8-row boundaries (always on in post_proc.c)
                for (int k = 0; k < 64; k++) {
                    pat[l * 64 + k] = (int8_t)((pat[l * 64 + k] * (int)fg_deblock_factor[v]) >> 7);
                    pat[(l + 7) * 64 + k] = (int8_t)((pat[(l + 7) * 64 + k] * (int)fg_deblock_factor[v]) >> 7);
                }
        }
    free(B);
}

// every dimension must keep the reference inside its planes: its block loops compute min(8, size - position) in unsigned
// arithmetic, which wraps when a 16-sample block row / column has fewer than 8 samples left
static inline int fg_size_ok(int w, int h)
{
    for (int c = 0; c < 2; c++) {
        const int cw = w >> c, ch = h >> c;
        if ((cw % 16 && cw % 16 < 8) || (ch % 16 && ch % 16 < 8)) return 0;
    }
    return 1;
}

struct FgParamsDev {                 // per component: fg_compute_model_values (:768-807) already applied
    int8_t interval[3][256];         // intensity interval of an 8-bit block average, -1: none
    int16_t scale[3][8];             // comp_model_value[.][.][0] (chroma: >> 1)
    uint8_t hv[3][8][2];             // cut-off pair - 2: the pattern's (h, v)
    uint8_t present[3];
    uint8_t shift;                   // log2_scale_factor + 6
};

// OVGPUFilmGrain (== struct OVSEIFGrain's fields) -> device parameters; host side
template <class SEI>
static inline void fg_make_params(FgParamsDev *p, const SEI *fg)
{
    memset(p, 0, sizeof(*p));
    memset(p->interval, -1, sizeof(p->interval));
    p->shift = (uint8_t)(fg->log2_scale_factor + 6);
    for (int c = 0; c < 3; c++) {
        p->present[c] = fg->comp_model_present_flag[c] != 0;
        for (int i = 0; i < 8; i++) {
            for (int a = fg->intensity_interval_lower_bound[c][i]; a <= fg->intensity_interval_upper_bound[c][i]; a++) p->interval[c][a] = (int8_t)i;
            int m0 = fg->comp_model_value[c][i][0], m1 = fg->comp_model_value[c][i][1], m2 = fg->comp_model_value[c][i][2];
            if (fg->num_model_values_minus1[c] + 1 == 1) m1 = m2 = 8;
            else if (fg->num_model_values_minus1[c] + 1 == 2) m2 = m1;
            if (c) {
                m0 = (int16_t)m0 >> 1;
                m1 = (int16_t)(m1 << 1); m1 = m1 < 2 ? 2 : m1 > 14 ? 14 : m1;
                m2 = (int16_t)(m2 << 1); m2 = m2 < 2 ? 2 : m2 > 14 ? 14 : m2;
            }
            p->scale[c][i] = (int16_t)m0;
            // (a cut-off outside 2..14 would index outside the data base in the reference too; clamped here)
            p->hv[c][i][0] = (uint8_t)((m1 < 2 ? 2 : m1 > 14 ? 14 : m1) - 2); p->hv[c][i][1] = (uint8_t)((m2 < 2 ? 2 : m2 > 14 ? 14 : m2) - 2);
        }
    }
}

struct FgSeedK {
    static const int THREADS = 256, MIN_BLOCKS = 1, PHASES = 1;
    struct Args { uint32_t *seeds[3]; int nb16w[3], nb16h[3]; uint32_t seed0[3]; };
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st; (void)bx; (void)by;
        const int c = bz, nw = c == 0 ? a.nb16w[0] : (c == 1 ? a.nb16w[1] : a.nb16w[2]), nh = c == 0 ? a.nb16h[0] : (c == 1 ? a.nb16h[1] : a.nb16h[2]);
        uint32_t *out = c == 0 ? a.seeds[0] : (c == 1 ? a.seeds[1] : a.seeds[2]);
        const uint32_t s0 = c == 0 ? a.seed0[0] : (c == 1 ? a.seed0[1] : a.seed0[2]);
        for (int r = tid; r < nh; r += THREADS) {
            uint32_t e = s0;
            for (int n = 0; n < r * nw; n++) e = fg_prng(e);
            for (int n = 0; n < nw; n++) { out[r * nw + n] = e; e = fg_prng(e); }
        }
    }
};

struct FilmGrainK {
    static const int THREADS = 256, MIN_BLOCKS = 4, PHASES = 3;
    static const int COLS = 128;                                 // 16 chunks of 8 samples, 16 rows
    struct Blk { int base, scale, xs; };
    static const int SMEM_BYTES = 16 * 18 * 4 + 2 * 18 * (int)sizeof(Blk);

    struct Args {
        uint16_t *dst[3]; const uint16_t *src[3];
        int dst_pitch[3], src_pitch[3];                          // samples
        int w, h;                                                // luma size
        const int8_t *db;
        const uint32_t *seeds[3];
        FgParamsDev p;
    };
    struct State { int dummy; };

    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)st;
        const int c = bz, ch_ = c != 0, cw = a.w >> ch_, chh = a.h >> ch_;
        if (bx * COLS >= cw || by * 16 >= chh) return;
        const uint16_t *src = c == 0 ? a.src[0] : (c == 1 ? a.src[1] : a.src[2]);
        uint16_t *dst = c == 0 ? a.dst[0] : (c == 1 ? a.dst[1] : a.dst[2]);
        const int sp = c == 0 ? a.src_pitch[0] : (c == 1 ? a.src_pitch[1] : a.src_pitch[2]);
        const int dp = c == 0 ? a.dst_pitch[0] : (c == 1 ? a.dst_pitch[1] : a.dst_pitch[2]);
        const int present = c == 0 ? a.p.present[0] : (c == 1 ? a.p.present[1] : a.p.present[2]);
        int *psum = (int *)smem;                                 // [row 0..15][chunk -1 .. 16]
        Blk *blk = (Blk *)(smem + 16 * 18 * 4);                  // [block row 0..1][block column -1 .. 16]
        const int r = tid >> 4, ck = tid & 15;

        if (P == 0) {
            if (!present) return;
            // sums of 8 horizontally adjacent samples: the CTA's own 16 chunks per row, and the chunk left / right of them
            for (int i = tid; i < 16 * 18; i += THREADS) {
                const int rr = i / 18, cc = i - rr * 18, x = bx * COLS + (cc - 1) * 8, y = by * 16 + rr;
                int s = 0;
                if (y < chh && x >= 0 && x < cw) {
                    const uint16_t *q = src + (size_t)y * sp + x;
                    for (int k = 0; k < 8; k++) if (x + k < cw) s += q[k];
                }
                psum[i] = s;
            }
            return;
        }
        if (P == 1) {
            if (!present || tid >= 36) return;
            const int brow = tid / 18, bcol = tid - brow * 18, gx = bx * COLS + (bcol - 1) * 8, gy = by * 16 + brow * 8;
            Blk b; b.base = 0; b.scale = 0; b.xs = 0;
            if (gx >= 0 && gx < cw && gy < chh) {
                const int ys = ovd_min(8, chh - gy), xs = ovd_min(8, cw - gx);
                unsigned sum = 0;
                for (int k = 0; k < ys; k++) sum += (unsigned)psum[(brow * 8 + k) * 18 + bcol];
                const unsigned avg = ovd_umin((sum / (unsigned)(ys * xs)) >> 2, 255u);           // fg_compute_block_avg at 10 bit
                const int it = c == 0 ? a.p.interval[0][avg] : (c == 1 ? a.p.interval[1][avg] : a.p.interval[2][avg]);
                if (it >= 0) {
                    const uint32_t *sd = c == 0 ? a.seeds[0] : (c == 1 ? a.seeds[1] : a.seeds[2]);
                    const uint32_t e = sd[(gy >> 4) * ((cw + 15) >> 4) + (gx >> 4)];
                    const int ko = (int)(((e >> 16) % 52) & 0xfffc) + (gx & 8), lo = (int)(((e & 0xffff) % 56) & 0xfff8) + (gy & 8);
                    const int sc = c == 0 ? a.p.scale[0][it] : (c == 1 ? a.p.scale[1][it] : a.p.scale[2][it]);
                    const int hh = c == 0 ? a.p.hv[0][it][0] : (c == 1 ? a.p.hv[1][it][0] : a.p.hv[2][it][0]);
                    const int vv = c == 0 ? a.p.hv[0][it][1] : (c == 1 ? a.p.hv[1][it][1] : a.p.hv[2][it][1]);
                    b.base = (hh * 13 + vv) * 4096 + lo * 64 + ko; b.scale = (e & 1) ? -sc : sc; b.xs = xs;
                }
            }
            blk[tid] = b;
            return;
        }
        const int x = bx * COLS + ck * 8, y = by * 16 + r;
        if (x >= cw || y >= chh) return;
        const int n = ovd_min(8, cw - x);
        const uint16_t *q = src + (size_t)y * sp + x;
        uint16_t *d = dst + (size_t)y * dp + x;
        if (!present) { for (int k = 0; k < n; k++) d[k] = q[k]; return; }
        const int sh = a.p.shift, l = r & 7;
        const Blk me = blk[(r >> 3) * 18 + ck + 1], lf = blk[(r >> 3) * 18 + ck], rt = blk[(r >> 3) * 18 + ck + 2];
        int g[8];
#pragma unroll
        for (int k = 0; k < 8; k++) g[k] = k < me.xs ? (me.scale * (int)a.db[me.base + l * 64 + k]) >> sh : 0;
        // smoothing across the vertical 8x8 edges of the stripe (fg_deblock_grain_stripe, :598-616): the edge left of this block
        // exists when there is a block there, the edge right of it when the next block starts inside the picture
        const int g0 = g[0], g1 = g[1], g6 = g[6], g7 = g[7];
        if (x >= 8) { const int p7 = lf.xs > 7 ? (lf.scale * (int)a.db[lf.base + l * 64 + 7]) >> sh : 0; g[0] = (p7 + (g0 << 1) + g1) >> 2; }
        if (x + 8 < cw) { const int n0 = rt.xs > 0 ? (rt.scale * (int)a.db[rt.base + l * 64]) >> sh : 0; g[7] = (g6 + (g7 << 1) + n0) >> 2; }
        for (int k = 0; k < n; k++) d[k] = (uint16_t)ovd_clamp((int)q[k] + (g[k] << 2), 0, 1023);           // fg_blend_stripe at 10 bit
    }
};
