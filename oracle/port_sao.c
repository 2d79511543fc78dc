/* oracle/port_sao.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Whole-picture restatement of OpenVVC's SAO (libovvc/rcn_sao.c).  The reference runs it per CTU with a
 * six-row downward shift from a padded copy of the deblocked samples (rcn_sao.c:190-293); since every
 * sample is filtered exactly once with the parameters of the CTU it belongs to and reads only deblocked
 * neighbours, this is the same as one pass src -> dst over the picture.
 */
#include "oracle.h"
#include <string.h>

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int sgn(int a, int b) { return a > b ? 1 : (a == b ? 0 : -1); }

void ovo_sao_picture(uint16_t *const dst[3], const int dst_stride[3],
                     const uint16_t *const src[3], const int src_stride[3],
                     int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                     const OVOSaoCtu *ctu)
{
    ovo_sao_picture_tiles(dst, dst_stride, src, src_stride, pic_w, pic_h, log2_ctu_s, bitdepth, ctu, NULL);
}

/* first / one-past-last CTU of the tile that holds CTU `c` along one axis */
static void tile_span(const uint16_t *bd, int nb, int c, int nb_ctu, int *c0, int *c1)
{
    *c0 = 0; *c1 = nb_ctu;
    for (int i = 0; bd && i < nb; i++) if (c >= bd[i] && c < bd[i + 1]) { *c0 = bd[i]; *c1 = bd[i + 1]; }
}

void ovo_sao_picture_tiles(uint16_t *const dst[3], const int dst_stride[3],
                           const uint16_t *const src[3], const int src_stride[3],
                           int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                           const OVOSaoCtu *ctu, const OVOTileGrid *tiles)
{
    /* neighbour offsets per eo_class: horizontal, vertical, 45, 135 (rcn_sao.c:86-91) */
    static const int8_t pos[4][2][2] = {
        { { -1, 0 }, { 1, 0 } }, { { 0, -1 }, { 0, 1 } }, { { -1, -1 }, { 1, 1 } }, { { 1, -1 }, { -1, 1 } } };
    int ctu_s = 1 << log2_ctu_s;
    int nb_w = (pic_w + ctu_s - 1) >> log2_ctu_s;
    int nb_h = (pic_h + ctu_s - 1) >> log2_ctu_s;
    int maxv = (1 << bitdepth) - 1;

    for (int c = 0; c < 3; c++) {
        int sh = c ? 1 : 0;
        int w = pic_w >> sh, h = pic_h >> sh, cs = ctu_s >> sh;
        const uint16_t *s = src[c];
        uint16_t *d = dst[c];
        int ss = src_stride[c], ds = dst_stride[c];

        for (int y = 0; y < h; y++)
            memcpy(d + (size_t)y * ds, s + (size_t)y * ss, w * sizeof(uint16_t));

        for (int cy = 0; cy < nb_h; cy++)
            for (int cx = 0; cx < nb_w; cx++) {
                const OVOSaoCtu *p = &ctu[cy * nb_w + cx];
                int x0 = cx * cs, y0 = cy * cs;
                int x1 = x0 + cs < w ? x0 + cs : w;
                int y1 = y0 + cs < h ? y0 + cs : h;
                /* the rect entry (tile) of this CTU: its edges are borders (rcn_sao.c:211-214, 241-245) */
                int tcx0, tcx1, tcy0, tcy1;
                tile_span(tiles ? tiles->col_bd : NULL, tiles ? tiles->nb_cols : 0, cx, nb_w, &tcx0, &tcx1);
                tile_span(tiles ? tiles->row_bd : NULL, tiles ? tiles->nb_rows : 0, cy, nb_h, &tcy0, &tcy1);
                const int tx0 = tcx0 * cs, tx1 = tcx1 * cs < w ? tcx1 * cs : w, ty0 = tcy0 * cs, ty1 = tcy1 * cs < h ? tcy1 * cs : h;

                if (p->type_idx[c] == 1) {            /* band, rcn_sao.c:46-76 */
                    int table[32] = { 0 };
                    for (int k = 0; k < 4; k++)
                        table[(k + p->band_position[c]) & 31] = p->offset_val[c][k];
                    for (int y = y0; y < y1; y++)
                        for (int x = x0; x < x1; x++) {
                            int v = s[(size_t)y * ss + x];
                            d[(size_t)y * ds + x] = (uint16_t)clipi(v + table[v >> (bitdepth - 5)], 0, maxv);
                        }
                } else if (p->type_idx[c] == 2) {     /* edge, rcn_sao.c:79-117 + border rules :162-177 */
                    int eo = p->eo_class[c];
                    int xs = x0, xe = x1, ys = y0, ye = y1;
                    if (eo != 1) { if (xs == tx0) xs = tx0 + 1; if (xe == tx1) xe = tx1 - 1; }
                    if (eo != 0) { if (ys == ty0) ys = ty0 + 1; if (ye == ty1) ye = ty1 - 1; }
                    for (int y = ys; y < ye; y++) {
                        /* reference quirk: in an entry of a single CTU row, rcn_sao_first_pix_rows (rcn_sao.c:259-272)
                         * carries the BOTTOM border flag, so the last of its 6 (3 chroma) rows is dropped */
                        if (tcy1 - tcy0 == 1 && eo != 0 && y == ty0 + (6 >> sh) - 1) continue;
                        for (int x = xs; x < xe; x++) {
                            int v = s[(size_t)y * ss + x];
                            int a = s[(size_t)(y + pos[eo][0][1]) * ss + x + pos[eo][0][0]];
                            int b = s[(size_t)(y + pos[eo][1][1]) * ss + x + pos[eo][1][0]];
                            int idx = 2 + sgn(v, a) + sgn(v, b);
                            d[(size_t)y * ds + x] = (uint16_t)clipi(v + p->offset_val[c][idx], 0, maxv);
                        }
                    }
                }
            }
    }
}
