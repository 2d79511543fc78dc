/* oracle/port_alf.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Whole-picture restatement of OpenVVC's ALF / CC-ALF (libovvc/rcn_alf.c).  The reference filters CTU by
 * CTU from a padded copy ("filter region", rcn_ctu.c:316-510) whose margins hold the *pre-ALF* samples of
 * the neighbouring CTUs, or the replicated border sample at picture edges.  Here that copy is replaced by
 * a separate source picture read with coordinates clamped to the picture -- the same samples.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---------------------------------------------------------------------------------------------------
 * Coefficient tables  (rcn_alf.c:61-281)
 * The two constant tables below are the normative VVC fixed ALF filters (H.266 eq. 89/90); values
 * checked against the reference build by tests/test_oracle_alf.py::test_tables_match_reference.
 * ------------------------------------------------------------------------------------------------- */
static const int8_t alf_fixed_coeff[64][12] = {
    { 0, 0, 2, -3, 1, -4, 1, 7, -1, 1, -1, 5 },        { 0, 0, 0, 0, 0, -1, 0, 1, 0, 0, -1, 2 },
    { 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0 },            { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 1 },
    { 2, 2, -7, -3, 0, -5, 13, 22, 12, -3, -3, 17 },   { -1, 0, 6, -8, 1, -5, 1, 23, 0, 2, -5, 10 },
    { 0, 0, -1, -1, 0, -1, 2, 1, 0, 0, -1, 4 },        { 0, 0, 3, -11, 1, 0, -1, 35, 5, 2, -9, 9 },
    { 0, 0, 8, -8, -2, -7, 4, 4, 2, 1, -1, 25 },       { 0, 0, 1, -1, 0, -3, 1, 3, -1, 1, -1, 3 },
    { 0, 0, 3, -3, 0, -6, 5, -1, 2, 1, -4, 21 },       { -7, 1, 5, 4, -3, 5, 11, 13, 12, -8, 11, 12 },
    { -5, -3, 6, -2, -3, 8, 14, 15, 2, -7, 11, 16 },   { 2, -1, -6, -5, -2, -2, 20, 14, -4, 0, -3, 25 },
    { 3, 1, -8, -4, 0, -8, 22, 5, -3, 2, -10, 29 },    { 2, 1, -7, -1, 2, -11, 23, -5, 0, 2, -10, 29 },
    { -6, -3, 8, 9, -4, 8, 9, 7, 14, -2, 8, 9 },       { 2, 1, -4, -7, 0, -8, 17, 22, 1, -1, -4, 23 },
    { 3, 0, -5, -7, 0, -7, 15, 18, -5, 0, -5, 27 },    { 2, 0, 0, -7, 1, -10, 13, 13, -4, 2, -7, 24 },
    { 3, 3, -13, 4, -2, -5, 9, 21, 25, -2, -3, 12 },   { -5, -2, 7, -3, -7, 9, 8, 9, 16, -2, 15, 12 },
    { 0, -1, 0, -7, -5, 4, 11, 11, 8, -6, 12, 21 },    { 3, -2, -3, -8, -4, -1, 16, 15, -2, -3, 3, 26 },
    { 2, 1, -5, -4, -1, -8, 16, 4, -2, 1, -7, 33 },    { 2, 1, -4, -2, 1, -10, 17, -2, 0, 2, -11, 33 },
    { 1, -2, 7, -15, -16, 10, 8, 8, 20, 11, 14, 11 },  { 2, 2, 3, -13, -13, 4, 8, 12, 2, -3, 16, 24 },
    { 1, 4, 0, -7, -8, -4, 9, 9, -2, -2, 8, 29 },      { 1, 1, 2, -4, -1, -6, 6, 3, -1, -1, -3, 30 },
    { -7, 3, 2, 10, -2, 3, 7, 11, 19, -7, 8, 10 },     { 0, -2, -5, -3, -2, 4, 20, 15, -1, -3, -1, 22 },
    { 3, -1, -8, -4, -1, -4, 22, 8, -4, 2, -8, 28 },   { 0, 3, -14, 3, 0, 1, 19, 17, 8, -3, -7, 20 },
    { 0, 2, -1, -8, 3, -6, 5, 21, 1, 1, -9, 13 },      { -4, -2, 8, 20, -2, 2, 3, 5, 21, 4, 6, 1 },
    { 2, -2, -3, -9, -4, 2, 14, 16, 3, -6, 8, 24 },    { 2, 1, 5, -16, -7, 2, 3, 11, 15, -3, 11, 22 },
    { 1, 2, 3, -11, -2, -5, 4, 8, 9, -3, -2, 26 },     { 0, -1, 10, -9, -1, -8, 2, 3, 4, 0, 0, 29 },
    { 1, 2, 0, -5, 1, -9, 9, 3, 0, 1, -7, 20 },        { -2, 8, -6, -4, 3, -9, -8, 45, 14, 2, -13, 7 },
    { 1, -1, 16, -19, -8, -4, -3, 2, 19, 0, 4, 30 },   { 1, 1, -3, 0, 2, -11, 15, -5, 1, 2, -9, 24 },
    { 0, 1, -2, 0, 1, -4, 4, 0, 0, 1, -4, 7 },         { 0, 1, 2, -5, 1, -6, 4, 10, -2, 1, -4, 10 },
    { 3, 0, -3, -6, -2, -6, 14, 8, -1, -1, -3, 31 },   { 0, 1, 0, -2, 1, -6, 5, 1, 0, 1, -5, 13 },
    { 3, 1, 9, -19, -21, 9, 7, 6, 13, 5, 15, 21 },     { 2, 4, 3, -12, -13, 1, 7, 8, 3, 0, 12, 26 },
    { 3, 1, -8, -2, 0, -6, 18, 2, -2, 3, -10, 23 },    { 1, 1, -4, -1, 1, -5, 8, 1, -1, 2, -5, 10 },
    { 0, 1, -1, 0, 0, -2, 2, 0, 0, 1, -2, 3 },         { 1, 1, -2, -7, 1, -7, 14, 18, 0, 0, -7, 21 },
    { 0, 1, 0, -2, 0, -7, 8, 1, -2, 0, -3, 24 },       { 0, 1, 1, -2, 2, -10, 10, 0, -2, 1, -7, 23 },
    { 0, 2, 2, -11, 2, -4, -3, 39, 7, 1, -10, 9 },     { 1, 0, 13, -16, -5, -6, -1, 8, 6, 0, 6, 29 },
    { 1, 3, 1, -6, -4, -7, 9, 6, -3, -2, 3, 33 },      { 4, 0, -17, -1, -1, 5, 26, 8, -2, 3, -15, 30 },
    { 0, 1, -2, 0, 2, -8, 12, -6, 1, 1, -6, 16 },      { 0, 0, 0, -1, 1, -4, 4, 0, 0, 0, -3, 11 },
    { 0, 1, 2, -8, 2, -6, 5, 15, 0, 2, -7, 9 },        { 1, -1, 12, -15, -7, -2, 3, 6, 6, -1, 7, 30 },
};

static const uint8_t alf_class_to_filter[16][25] = {
    { 8, 2, 2, 2, 3, 4, 53, 9, 9, 52, 4, 4, 5, 9, 2, 8, 10, 9, 1, 3, 39, 39, 10, 9, 52 },
    { 11, 12, 13, 14, 15, 30, 11, 17, 18, 19, 16, 20, 20, 4, 53, 21, 22, 23, 14, 25, 26, 26, 27, 28, 10 },
    { 16, 12, 31, 32, 14, 16, 30, 33, 53, 34, 35, 16, 20, 4, 7, 16, 21, 36, 18, 19, 21, 26, 37, 38, 39 },
    { 35, 11, 13, 14, 43, 35, 16, 4, 34, 62, 35, 35, 30, 56, 7, 35, 21, 38, 24, 40, 16, 21, 48, 57, 39 },
    { 11, 31, 32, 43, 44, 16, 4, 17, 34, 45, 30, 20, 20, 7, 5, 21, 22, 46, 40, 47, 26, 48, 63, 58, 10 },
    { 12, 13, 50, 51, 52, 11, 17, 53, 45, 9, 30, 4, 53, 19, 0, 22, 23, 25, 43, 44, 37, 27, 28, 10, 55 },
    { 30, 33, 62, 51, 44, 20, 41, 56, 34, 45, 20, 41, 41, 56, 5, 30, 56, 38, 40, 47, 11, 37, 42, 57, 8 },
    { 35, 11, 23, 32, 14, 35, 20, 4, 17, 18, 21, 20, 20, 20, 4, 16, 21, 36, 46, 25, 41, 26, 48, 49, 58 },
    { 12, 31, 59, 59, 3, 33, 33, 59, 59, 52, 4, 33, 17, 59, 55, 22, 36, 59, 59, 60, 22, 36, 59, 25, 55 },
    { 31, 25, 15, 60, 60, 22, 17, 19, 55, 55, 20, 20, 53, 19, 55, 22, 46, 25, 43, 60, 37, 28, 10, 55, 52 },
    { 12, 31, 32, 50, 51, 11, 33, 53, 19, 45, 16, 4, 4, 53, 5, 22, 36, 18, 25, 43, 26, 27, 27, 28, 10 },
    { 5, 2, 44, 52, 3, 4, 53, 45, 9, 3, 4, 56, 5, 0, 2, 5, 10, 47, 52, 3, 63, 39, 10, 9, 52 },
    { 12, 34, 44, 44, 3, 56, 56, 62, 45, 9, 56, 56, 7, 5, 0, 22, 38, 40, 47, 52, 48, 57, 39, 10, 9 },
    { 35, 11, 23, 14, 51, 35, 20, 41, 56, 62, 16, 20, 41, 56, 7, 16, 21, 38, 24, 40, 26, 26, 42, 57, 39 },
    { 33, 34, 51, 51, 52, 41, 41, 34, 62, 0, 41, 41, 56, 7, 5, 56, 38, 38, 40, 44, 37, 42, 57, 39, 10 },
    { 16, 31, 32, 15, 60, 30, 4, 17, 19, 25, 22, 20, 4, 53, 19, 21, 22, 46, 25, 55, 26, 48, 63, 58, 55 },
};

/* Geometric transposes of the 7x7 diamond expressed as a permutation of the 12 coefficient slots
 * (rcn_alf.c:149-155): 0 identity, 1 diagonal flip, 2 vertical flip, 3 rotation. */
static const uint8_t alf_transpose_perm[4][13] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12 },
    { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12 },
    { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12 },
    { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12 },
};

static void alf_clip_values(int16_t lut[4], int bitdepth)
{   /* rcn_alf.c:157-163 */
    lut[0] = (int16_t)(1 << bitdepth);
    lut[1] = (int16_t)(1 << (bitdepth - 3));
    lut[2] = (int16_t)(1 << (bitdepth - 5));
    lut[3] = (int16_t)(1 << (bitdepth - 7));
}

void ovo_alf_build_tables(OVOAlfTables *t, const OVOAlfAps *luma_aps, int nb_luma_aps,
                          const OVOAlfAps *chroma_aps, int bitdepth)
{
    int16_t clip_lut[4];
    alf_clip_values(clip_lut, bitdepth);

    /* fixed sets: rcn_alf_init_fixed_filter_sets, rcn_alf.c:174-194 */
    for (int s = 0; s < 16; s++)
        for (int tr = 0; tr < 4; tr++)
            for (int c = 0; c < 25; c++) {
                int16_t *co = &t->luma_coeff[s][(tr * 25 + c) * 13];
                int16_t *cl = &t->luma_clip [s][(tr * 25 + c) * 13];
                for (int k = 0; k < 12; k++) {
                    co[k] = alf_fixed_coeff[alf_class_to_filter[s][c]][alf_transpose_perm[tr][k]];
                    cl[k] = clip_lut[0];
                }
                co[12] = 128;
                cl[12] = clip_lut[0];
            }

    /* APS sets: alf_init_filter_l, rcn_alf.c:196-238 */
    for (int a = 0; a < nb_luma_aps; a++) {
        const OVOAlfAps *aps = &luma_aps[a];
        for (int c = 0; c < 25; c++) {
            int f = aps->luma_coeff_delta_idx[c];
            int16_t co[13], cl[13];
            for (int k = 0; k < 12; k++) {
                co[k] = aps->luma_coeff[f][k];
                cl[k] = clip_lut[aps->luma_clip_flag ? aps->luma_clip_idx[f][k] : 0];
            }
            co[12] = 128;
            cl[12] = clip_lut[0];
            for (int tr = 0; tr < 4; tr++)
                for (int k = 0; k < 13; k++) {
                    t->luma_coeff[16 + a][(tr * 25 + c) * 13 + k] = co[alf_transpose_perm[tr][k]];
                    t->luma_clip [16 + a][(tr * 25 + c) * 13 + k] = cl[alf_transpose_perm[tr][k]];
                }
        }
    }

    /* chroma alternatives: alf_init_filter_c, rcn_alf.c:240-262 */
    if (chroma_aps) {
        for (int alt = 0; alt < chroma_aps->num_chroma_alts; alt++) {
            for (int k = 0; k < 6; k++) {
                t->chroma_coeff[alt][k] = chroma_aps->chroma_coeff[alt][k];
                t->chroma_clip [alt][k] = clip_lut[chroma_aps->chroma_clip_flag ? chroma_aps->chroma_clip_idx[alt][k] : 0];
            }
            t->chroma_coeff[alt][6] = 128;
            t->chroma_clip [alt][6] = clip_lut[0];
        }
    }
}

/* ---------------------------------------------------------------------------------------------------
 * Classification  (rcn_alf.c:283-738)
 * ------------------------------------------------------------------------------------------------- */
/* (x0, y0) .. (x1, y1): the rect entry (tile) the samples are read for -- the whole plane without tiles */
typedef struct { const uint16_t *p; int stride, w, h, x0, y0, x1, y1; } Plane;

static inline int px(const Plane *pl, int x, int y)
{   /* replicate padding == the filter-region margins at the borders of the picture / rect entry, rcn_ctu.c:363-508 */
    x = clipi(x, pl->x0, pl->x1 - 1);
    y = clipi(y, pl->y0, pl->y1 - 1);
    return pl->p[(size_t)y * pl->stride + x];
}

/* first / one-past-last CTU of the tile that holds CTU `c` along one axis */
static void alf_tile_span(const uint16_t *bd, int nb, int c, int nb_ctu, int *c0, int *c1)
{
    *c0 = 0; *c1 = nb_ctu;
    for (int i = 0; bd && i < nb; i++) if (c >= bd[i] && c < bd[i + 1]) { *c0 = bd[i]; *c1 = bd[i + 1]; }
}
/* the plane restricted to the tile of luma CTU (cx, cy); sh = 1 for the chroma planes */
static Plane alf_tile_plane(Plane pl, const OVOTileGrid *tg, int cx, int cy, int nb_w, int nb_h, int log2_ctu_s, int sh)
{
    int c0, c1, r0, r1;
    alf_tile_span(tg ? tg->col_bd : NULL, tg ? tg->nb_cols : 0, cx, nb_w, &c0, &c1);
    alf_tile_span(tg ? tg->row_bd : NULL, tg ? tg->nb_rows : 0, cy, nb_h, &r0, &r1);
    pl.x0 = (c0 << log2_ctu_s) >> sh; pl.y0 = (r0 << log2_ctu_s) >> sh;
    pl.x1 = (c1 << log2_ctu_s) >> sh; pl.y1 = (r1 << log2_ctu_s) >> sh;
    if (pl.x1 > pl.w) pl.x1 = pl.w;
    if (pl.y1 > pl.h) pl.y1 = pl.h;
    return pl;
}

/* The four 1-D Laplacians at one sample.  y_up / y_dn are the rows used as vertical neighbours: normally
 * y-1 / y+1, but the row across the ALF virtual boundary is replaced (rcn_alf.c:464-468). */
static inline void lap4(const Plane *pl, int x, int y, int y_up, int y_dn, int s[4])
{
    int c2 = px(pl, x, y) << 1;
    s[0] += abs(c2 - px(pl, x, y_up) - px(pl, x, y_dn));            /* VER   */
    s[1] += abs(c2 - px(pl, x - 1, y) - px(pl, x + 1, y));          /* HOR   */
    s[2] += abs(c2 - px(pl, x - 1, y_up) - px(pl, x + 1, y_dn));    /* DIAG0 */
    s[3] += abs(c2 - px(pl, x + 1, y_up) - px(pl, x - 1, y_dn));    /* DIAG1 */
}

/* alf_derive_filter_idx, rcn_alf.c:283-345 */
static uint8_t alf_class_from_sums(uint32_t sum_v, uint32_t sum_h, uint32_t sum_d, uint32_t sum_b,
                                   int shift, int near_vb)
{
    static const uint8_t th[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
    static const uint8_t tr_lut[8] = { 0, 1, 0, 2, 2, 3, 1, 3 };
    uint32_t scale = near_vb ? 96 : 64;
    int act = clipi((int)(((sum_h + sum_v) * scale) >> shift), 0, 15);
    int cls = th[act];
    uint32_t max_hv, min_hv, max_db, min_db, max_dir, min_dir;
    int dir_hv, dir_db, main_dir, sec_dir;

    if (sum_v > sum_h) { max_hv = sum_v; min_hv = sum_h; dir_hv = 1; }
    else               { max_hv = sum_h; min_hv = sum_v; dir_hv = 3; }
    if (sum_d > sum_b) { max_db = sum_d; min_db = sum_b; dir_db = 0; }
    else               { max_db = sum_b; min_db = sum_d; dir_db = 2; }

    if (max_db * min_hv > max_hv * min_db) { max_dir = max_db; min_dir = min_db; main_dir = dir_db; sec_dir = dir_hv; }
    else                                   { max_dir = max_hv; min_dir = min_hv; main_dir = dir_hv; sec_dir = dir_db; }

    if (max_dir * 2 > 9 * min_dir)      cls += (((main_dir & 1) << 1) + 2) * 5;
    else if (max_dir > 2 * min_dir)     cls += (((main_dir & 1) << 1) + 1) * 5;

    return (uint8_t)(cls | (tr_lut[(main_dir << 1) + (sec_dir >> 1)] << 5));
}

/* Classify the 4x4 block whose top-left luma sample is (bx, by).
 * ctu_y0 / ctu_h: top row and height of the CTU row holding the block. */
static uint8_t alf_classify_4x4(const Plane *pl, int bx, int by, int ctu_y0, int ctu_h, int ctu_s, int bitdepth)
{
    /* virtual boundary, CTU relative (rcn_alf.c:724) */
    int vb = (ctu_h < ctu_s) ? pl->h : ctu_h - 4;
    int y_rel = by - ctu_y0;
    /* the reference classifies in 32x32 blocks and picks the VB-aware routine for a whole 32-row block
     * (rcn_alf.c:695); inside it only the 4-row groups touching the boundary differ */
    int blk_y = y_rel & ~31;
    int blk_h = (ctu_h - blk_y) < 32 ? (ctu_h - blk_y) : 32;
    int vb_blk = (blk_y + blk_h >= vb);
    int first_pair = 0, last_pair = 3, near_vb = 0;
    int s[4] = { 0, 0, 0, 0 };

    if (vb_blk) {
        if (y_rel == vb - 4)  { last_pair = 2;  near_vb = 1; }   /* rcn_alf.c:517-547 */
        else if (y_rel == vb) { first_pair = 1; near_vb = 1; }   /* rcn_alf.c:548-580 */
        else if (y_rel + 4 >= vb) return 0;                      /* never written by the reference loops */
    }

    /* row pair k covers rows by-2+2k, by-1+2k; columns bx-2 .. bx+5; samples on a checkerboard
     * (rcn_alf.c:602-657: even rows take columns bx-2+2m, odd rows bx-1+2m) */
    for (int k = first_pair; k <= last_pair; k++) {
        int y0 = by - 2 + 2 * k, y1 = y0 + 1;
        int vb_abs = ctu_y0 + vb;
        for (int m = 0; m < 4; m++) {
            int y0_up = y0 - 1, y1_dn = y1 + 1;
            if (vb_blk) {
                if (y1 == vb_abs - 1) y1_dn = y1;   /* "src3 = src2", rcn_alf.c:464 */
                if (y0 == vb_abs)     y0_up = y0;   /* "src0 = src1", rcn_alf.c:466 */
            }
            lap4(pl, bx - 2 + 2 * m, y0, y0_up, y0 + 1, s);
            lap4(pl, bx - 1 + 2 * m, y1, y1 - 1, y1_dn, s);
        }
    }
    return alf_class_from_sums((uint32_t)s[0], (uint32_t)s[1], (uint32_t)s[2], (uint32_t)s[3],
                               bitdepth + 4, near_vb);
}

static void alf_classify_picture_tiles(uint8_t *cls, const uint16_t *src, int stride,
                                       int pic_w, int pic_h, int log2_ctu_s, int bitdepth, const OVOTileGrid *tg)
{
    Plane pl = { src, stride, pic_w, pic_h, 0, 0, pic_w, pic_h };
    int ctu_s = 1 << log2_ctu_s;
    int w4 = pic_w >> 2;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s, nb_h = (pic_h + ctu_s - 1) >> log2_ctu_s;
    for (int by = 0; by < pic_h; by += 4) {
        int ctu_y0 = by & ~(ctu_s - 1);
        int ctu_h = (ctu_y0 + ctu_s > pic_h) ? pic_h - ctu_y0 : ctu_s;
        for (int bx = 0; bx < pic_w; bx += 4) {
            Plane tp = tg ? alf_tile_plane(pl, tg, bx >> log2_ctu_s, by >> log2_ctu_s, nb_w, nb_h, log2_ctu_s, 0) : pl;
            cls[(by >> 2) * w4 + (bx >> 2)] = alf_classify_4x4(&tp, bx, by, ctu_y0, ctu_h, ctu_s, bitdepth);
        }
    }
}

void ovo_alf_classify_picture(uint8_t *cls, const uint16_t *src, int stride,
                              int pic_w, int pic_h, int log2_ctu_s, int bitdepth)
{
    alf_classify_picture_tiles(cls, src, stride, pic_w, pic_h, log2_ctu_s, bitdepth, NULL);
}

/* ---------------------------------------------------------------------------------------------------
 * Filters  (rcn_alf.c:740-1272) and the per-CTU driver (rcn_alf.c:1285-1433)
 * ------------------------------------------------------------------------------------------------- */
static inline int clip_pair(int clip, int ref, int a, int b)
{   /* alf_clip, rcn_alf.c:165-171 */
    return clipi(a - ref, -clip, clip) + clipi(b - ref, -clip, clip);
}

/* Vertical reach allowed for the sample row y_vb (CTU relative) next to the virtual boundary vb, and
 * whether it takes the coarse rounding (rcn_alf.c:1200-1221 luma reach 3/window 4; :1005-1025 chroma
 * reach 2/window 2). */
static inline int vb_reach(int y_vb, int vb, int window, int *coarse)
{
    int reach = 3;
    *coarse = (y_vb == vb - 1) || (y_vb == vb);
    if (y_vb < vb && y_vb >= vb - window)            reach = vb - 1 - y_vb;
    else if (y_vb >= vb && y_vb <= vb + window - 1)  reach = y_vb - vb;
    return reach;
}

static void alf_luma_ctu(uint16_t *dst, int dst_stride, const Plane *src, const uint8_t *cls, int cls_stride,
                         int x0, int y0, int w, int h, int ctu_s, int vb,
                         const int16_t *coeff_set, const int16_t *clip_set, int bitdepth)
{
    int maxv = (1 << bitdepth) - 1;
    for (int y = y0; y < y0 + h; y++) {
        int coarse;
        int reach = vb_reach(y & (ctu_s - 1), vb, 4, &coarse);
        int d1 = reach < 1 ? reach : 1, d2 = reach < 2 ? reach : 2, d3 = reach < 3 ? reach : 3;
        for (int x = x0; x < x0 + w; x++) {
            int ci = cls[(y >> 2) * cls_stride + (x >> 2)];
            const int16_t *c  = coeff_set + ((ci >> 5) * 25 + (ci & 31)) * 13;
            const int16_t *cl = clip_set  + ((ci >> 5) * 25 + (ci & 31)) * 13;
            int cur = (int16_t)px(src, x, y);
            int sum = 0;
#define P(dx, dy) px(src, x + (dx), y + (dy))
            sum += c[0]  * clip_pair(cl[0],  cur, P( 0,  d3), P( 0, -d3));
            sum += c[1]  * clip_pair(cl[1],  cur, P( 1,  d2), P(-1, -d2));
            sum += c[2]  * clip_pair(cl[2],  cur, P( 0,  d2), P( 0, -d2));
            sum += c[3]  * clip_pair(cl[3],  cur, P(-1,  d2), P( 1, -d2));
            sum += c[4]  * clip_pair(cl[4],  cur, P( 2,  d1), P(-2, -d1));
            sum += c[5]  * clip_pair(cl[5],  cur, P( 1,  d1), P(-1, -d1));
            sum += c[6]  * clip_pair(cl[6],  cur, P( 0,  d1), P( 0, -d1));
            sum += c[7]  * clip_pair(cl[7],  cur, P(-1,  d1), P( 1, -d1));
            sum += c[8]  * clip_pair(cl[8],  cur, P(-2,  d1), P( 2, -d1));
            sum += c[9]  * clip_pair(cl[9],  cur, P( 3,  0),  P(-3,  0));
            sum += c[10] * clip_pair(cl[10], cur, P( 2,  0),  P(-2,  0));
            sum += c[11] * clip_pair(cl[11], cur, P( 1,  0),  P(-1,  0));
#undef P
            sum = coarse ? (sum + 512) >> 10 : (sum + 64) >> 7;
            dst[(size_t)y * dst_stride + x] = (uint16_t)clipi(sum + cur, 0, maxv);
        }
    }
}

static void alf_chroma_ctu(uint16_t *dst, int dst_stride, const Plane *src,
                           int x0, int y0, int w, int h, int ctu_s_c, int vb,
                           const int16_t *c, const int16_t *cl, int bitdepth)
{
    int maxv = (1 << bitdepth) - 1;
    for (int y = y0; y < y0 + h; y++) {
        int coarse;
        int reach = vb_reach(y & (ctu_s_c - 1), vb, 2, &coarse);
        int d1 = reach < 1 ? reach : 1, d2 = reach < 2 ? reach : 2;
        for (int x = x0; x < x0 + w; x++) {
            int cur = (int16_t)px(src, x, y);
            int sum = 0;
#define P(dx, dy) px(src, x + (dx), y + (dy))
            sum += c[0] * clip_pair(cl[0], cur, P( 0,  d2), P( 0, -d2));
            sum += c[1] * clip_pair(cl[1], cur, P( 1,  d1), P(-1, -d1));
            sum += c[2] * clip_pair(cl[2], cur, P( 0,  d1), P( 0, -d1));
            sum += c[3] * clip_pair(cl[3], cur, P(-1,  d1), P( 1, -d1));
            sum += c[4] * clip_pair(cl[4], cur, P( 2,  0),  P(-2,  0));
            sum += c[5] * clip_pair(cl[5], cur, P( 1,  0),  P(-1,  0));
#undef P
            sum = coarse ? (sum + 512) >> 10 : (sum + 64) >> 7;
            dst[(size_t)y * dst_stride + x] = (uint16_t)clipi(sum + cur, 0, maxv);
        }
    }
}

/* cc_alf_filterBlk(VB), rcn_alf.c:740-872: refines the (already ALF-filtered) chroma in dst from the
 * pre-ALF luma. vb_pos is compared with the LUMA row modulo the CTU size. */
static void alf_cc_ctu(uint16_t *dst, int dst_stride, const Plane *luma,
                       int x0, int y0, int w, int h, int ctu_s, int vb_pos,
                       const int16_t *c, int bitdepth)
{
    int maxv = (1 << bitdepth) - 1;
    int half = 1 << (bitdepth - 1);
    for (int y = y0; y < y0 + h; y++) {
        int pos = (y << 1) & (ctu_s - 1);
        int o1 = 1, o2 = -1, o3 = 2;
        if (pos == vb_pos - 2 || pos == vb_pos + 1) o3 = o1;
        else if (pos == vb_pos - 1 || pos == vb_pos) o1 = o2 = o3 = 0;
        for (int x = x0; x < x0 + w; x++) {
            int lx = x << 1, ly = y << 1;
            int cur = (int16_t)px(luma, lx, ly);
            int sum = 0;
            sum += c[0] * (px(luma, lx,     ly + o2) - cur);
            sum += c[1] * (px(luma, lx - 1, ly)      - cur);
            sum += c[2] * (px(luma, lx + 1, ly)      - cur);
            sum += c[3] * (px(luma, lx - 1, ly + o1) - cur);
            sum += c[4] * (px(luma, lx,     ly + o1) - cur);
            sum += c[5] * (px(luma, lx + 1, ly + o1) - cur);
            sum += c[6] * (px(luma, lx,     ly + o3) - cur);
            sum = (sum + 64) >> 7;
            sum = clipi(sum + half, 0, maxv);
            sum += dst[(size_t)y * dst_stride + x] - half;
            dst[(size_t)y * dst_stride + x] = (uint16_t)clipi(sum, 0, maxv);
        }
    }
}

void ovo_alf_picture(uint16_t *const dst[3], const int dst_stride[3],
                     const uint16_t *const src[3], const int src_stride[3],
                     int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                     const OVOAlfTables *tab, const OVOAlfCtuParams *ctu)
{
    ovo_alf_picture_tiles(dst, dst_stride, src, src_stride, pic_w, pic_h, log2_ctu_s, bitdepth, tab, ctu, NULL);
}

/* tg != NULL: the edges of a tile (rect entry) are padded like picture borders (rcn_alf_filter_line's is_border flags ->
 * rcn_extend_filter_region, rcn_alf.c:1316-1330, rcn_ctu.c:316-510): samples are fetched clamped to the CTU's tile */
void ovo_alf_picture_tiles(uint16_t *const dst[3], const int dst_stride[3],
                           const uint16_t *const src[3], const int src_stride[3],
                           int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                           const OVOAlfTables *tab, const OVOAlfCtuParams *ctu, const OVOTileGrid *tg)
{
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    int nb_h = (pic_h + ctu_s - 1) >> log2_ctu_s;
    Plane pl_pic[3] = { { src[0], src_stride[0], pic_w, pic_h, 0, 0, pic_w, pic_h },
                        { src[1], src_stride[1], pic_w >> 1, pic_h >> 1, 0, 0, pic_w >> 1, pic_h >> 1 },
                        { src[2], src_stride[2], pic_w >> 1, pic_h >> 1, 0, 0, pic_w >> 1, pic_h >> 1 } };
    Plane pl[3] = { pl_pic[0], pl_pic[1], pl_pic[2] };
    int w4 = pic_w >> 2;
    uint8_t *cls = malloc((size_t)w4 * (pic_h >> 2));

    for (int c = 0; c < 3; c++)
        for (int y = 0; y < pl[c].h; y++)
            memcpy(dst[c] + (size_t)y * dst_stride[c], src[c] + (size_t)y * src_stride[c], pl[c].w * sizeof(uint16_t));

    alf_classify_picture_tiles(cls, src[0], src_stride[0], pic_w, pic_h, log2_ctu_s, bitdepth, tg);

    for (int cy = 0; cy < nb_h; cy++) {
        for (int cx = 0; cx < nb_w; cx++) {
            int rs = cy * nb_w + cx;
            if (tg) for (int c = 0; c < 3; c++) pl[c] = alf_tile_plane(pl_pic[c], tg, cx, cy, nb_w, nb_h, log2_ctu_s, c != 0);
            int x0 = cx << log2_ctu_s, y0 = cy << log2_ctu_s;
            int w = (x0 + ctu_s > pic_w) ? pic_w - x0 : ctu_s;
            int h = (y0 + ctu_s > pic_h) ? pic_h - y0 : ctu_s;
            int partial_v = (y0 + ctu_s > pic_h);
            uint8_t flag = ctu->ctb_alf_flag[rs];

            if (flag & 4) {   /* rcn_alf.c:1330-1362 */
                int set = ctu->ctb_alf_idx[rs];
                int vb = partial_v ? pic_h : h - 4;
                alf_luma_ctu(dst[0], dst_stride[0], &pl[0], cls, w4, x0, y0, w, h, ctu_s, vb,
                             tab->luma_coeff[set], tab->luma_clip[set], bitdepth);
            }
            for (int c = 1; c < 3; c++) {
                if ((c == 1 && (flag & 2)) || (c == 2 && (flag & 1))) {   /* rcn_alf.c:1367-1394 */
                    int alt = (c == 1) ? ctu->cb_alternative[rs] : ctu->cr_alternative[rs];
                    int vb = partial_v ? pic_h / 2 : (ctu_s - 4) / 2;
                    alf_chroma_ctu(dst[c], dst_stride[c], &pl[c], x0 / 2, y0 / 2, w / 2, h / 2, ctu_s / 2, vb,
                                   tab->chroma_coeff[alt], tab->chroma_clip[alt], bitdepth);
                }
                if ((c == 1 && ctu->cc_cb_enabled) || (c == 2 && ctu->cc_cr_enabled)) {   /* rcn_alf.c:1396-1427 */
                    int idx = (c == 1) ? ctu->cc_idx_cb[rs] : ctu->cc_idx_cr[rs];
                    if (idx) {
                        int vb_pos = partial_v ? pic_h / 2 : ctu_s - 4;   /* sic: chroma rows vs luma rows, rcn_alf.c:1417 */
                        alf_cc_ctu(dst[c], dst_stride[c], &pl[0], x0 / 2, y0 / 2, w / 2, h / 2, ctu_s, vb_pos,
                                   tab->cc_coeff[c - 1][idx - 1], bitdepth);
                    }
                }
            }
        }
    }
    free(cls);
}
