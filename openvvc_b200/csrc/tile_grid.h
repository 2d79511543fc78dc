// tile_grid.h -- the tile grid of a picture as the in-loop filter kernels see it (sm_100a; also compiled by tests/emu).
#pragma once
#include "dev.h"

// Tile grid of the picture (rect entries): the reference filters entry by entry and treats an entry's edges like picture borders
// (SAO: rcn_sao.c:211-214, 241-245; ALF: the is_border flags of rcn_alf_filter_line -> rcn_extend_filter_region,
// rcn_alf.c:1311-1330, rcn_ctu.c:316-510, i.e. replicated samples).  Boundaries in CTUs.  Kernels get a POINTER to it in device memory (NULL: one tile, the lookup is
// skipped): a run-time index into a by-value kernel argument would move the whole argument block to local memory.
struct TileGridDev { int nb_cols, nb_rows; uint16_t col_bd[22], row_bd[22]; };
// first / one-past-last CTU of the tile holding CTU c along one axis
OVD void tile_span(const uint16_t *bd, int nb, int c, int nb_ctu, int &c0, int &c1)
{
    c0 = 0; c1 = nb_ctu;
    for (int i = 0; i < nb; i++) if (c >= bd[i] && c < bd[i + 1]) { c0 = bd[i]; c1 = bd[i + 1]; }
}

