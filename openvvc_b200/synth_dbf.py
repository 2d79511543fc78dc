"""Seeded synthetic deblocking state (SURVEY.md §8(d) "DBF", seed 1003) and its picture-level map layout.

A block list (what the parser's side effects would describe: TU edges, intra / coded-residual / MV-mismatch
boundary strengths, affine sub-block edges, QPs) is turned into the picture-level bit planes of OVGPUDbfMaps
(include/ovvc_b200.h) -- the same content the reference accumulates per CTU in struct DBFInfo (ctudec.h:130-170)
through fill_ctb_bound / fill_bs_map / dbf_fill_qp_map / dbf_fill_aff_map (dbf_utils.h:41-148) and dbf_set_sb_edges
(drv_affine_mvp.c:3051-3070)."""
import numpy as np

from . import synth

DBF_BLK = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8), ("flags", np.uint8),
                    ("qp_y", np.uint8), ("qp_cb", np.uint8), ("qp_cr", np.uint8), ("pad", np.uint16)])
F_INTRA, F_CBF_L, F_CBF_CB, F_CBF_CR, F_MVEDGE, F_AFFINE = 1, 2, 4, 8, 16, 32
# map indices of OVGPUDbfMaps.ver / .hor
M_EDGE_TU, M_EDGE_SB, M_BS1, M_BS2, M_AFFINE, M_EDGE_C, M_BS1_CB, M_BS1_CR, M_BS2_C = range(9)
NB_MAPS = 9


def dbf_blocks(w, h, seed=1003, log2_ctu=7, min_log2=2, p_intra=0.15, p_affine=0.1, qp_lo=22, qp_hi=37):
    rng = np.random.default_rng(seed)
    part = synth.partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(6, log2_ctu), p_split=0.5)
    blks = np.zeros(len(part), dtype=DBF_BLK)
    for i, (x, y, lw, lh) in enumerate(part):
        b = blks[i]
        b["x0"], b["y0"], b["log2_w"], b["log2_h"] = x, y, lw, lh
        f = 0
        if rng.random() < p_intra:
            f |= F_INTRA
        else:
            if lw >= 3 and lh >= 3 and rng.random() < p_affine:
                f |= F_AFFINE
            if rng.random() < 0.4:
                f |= F_MVEDGE
        r = rng.random(3)
        if r[0] < 0.4:
            f |= F_CBF_L
        if r[1] < 0.25:
            f |= F_CBF_CB
        if r[2] < 0.25:
            f |= F_CBF_CR
        b["flags"] = f
        qp = int(rng.integers(qp_lo, qp_hi + 1))
        b["qp_y"] = qp
        b["qp_cb"] = max(0, qp - int(rng.integers(0, 4)))
        b["qp_cr"] = max(0, qp - int(rng.integers(0, 4)))
    return blks


def _pack(bits, n_u):
    """bits[k, n] (k = position along the edge direction, padded to a multiple of n_u; n = edge index) ->
    words[group, n] with bit j of group g = bits[g * n_u + j, n]."""
    k, n = bits.shape
    g = k // n_u
    b = bits.reshape(g, n_u, n).astype(np.uint64)
    wts = (np.uint64(1) << np.arange(n_u, dtype=np.uint64)).reshape(1, n_u, 1)
    return (b * wts).sum(axis=1).astype(np.uint32)


def dbf_maps(w, h, blks, log2_ctu=7, beta_offset=0, tc_offset=0):
    """Picture-level maps. ver[m][cy, X]: bit j = unit row cy * n_u + j of vertical edge column X (0..w_u);
    hor[m][cx, Y]: bit j = unit column cx * n_u + j of horizontal edge row Y (0..h_u); qp[c][Y, X] per 4x4 unit."""
    n_u = (1 << log2_ctu) >> 2
    w_u, h_u = w >> 2, h >> 2
    ncw, nch = -(-w_u // n_u), -(-h_u // n_u)
    V = np.zeros((NB_MAPS, nch * n_u, w_u + 1), dtype=bool)     # [map][unit row][edge column]
    H = np.zeros((NB_MAPS, ncw * n_u, h_u + 1), dtype=bool)     # [map][unit column][edge row]
    qp = np.zeros((3, h_u, w_u), dtype=np.uint8)
    for b in blks:
        x, y = int(b["x0"]) >> 2, int(b["y0"]) >> 2
        nw, nh = (1 << int(b["log2_w"])) >> 2, (1 << int(b["log2_h"])) >> 2
        f = int(b["flags"])

        def sides(m):
            V[m, y:y + nh, x] = True; V[m, y:y + nh, x + nw] = True
            H[m, x:x + nw, y] = True; H[m, x:x + nw, y + nh] = True
        sides(M_EDGE_TU); sides(M_EDGE_C)
        if f & F_INTRA:
            sides(M_BS2); sides(M_BS2_C)
        if f & (F_CBF_L | F_MVEDGE):
            sides(M_BS1)
        if f & F_CBF_CB:
            sides(M_BS1_CB)
        if f & F_CBF_CR:
            sides(M_BS1_CR)
        if f & F_AFFINE:
            V[M_AFFINE, y:y + nh, x:x + nw] = True
            H[M_AFFINE, x:x + nw, y:y + nh] = True
            for i in range(2, nw, 2):
                V[M_EDGE_SB, y:y + nh, x + i] = True
            for i in range(2, nh, 2):
                H[M_EDGE_SB, x:x + nw, y + i] = True
        qp[0, y:y + nh, x:x + nw] = b["qp_y"]; qp[1, y:y + nh, x:x + nw] = b["qp_cb"]; qp[2, y:y + nh, x:x + nw] = b["qp_cr"]
    ver = np.stack([_pack(V[m], n_u) for m in range(NB_MAPS)])      # [map][nch][w_u + 1]
    hor = np.stack([_pack(H[m], n_u) for m in range(NB_MAPS)])      # [map][ncw][h_u + 1]
    return dict(width=w, height=h, log2_ctu=log2_ctu, beta_offset=beta_offset, tc_offset=tc_offset,
                ver=np.ascontiguousarray(ver), hor=np.ascontiguousarray(hor), qp=np.ascontiguousarray(qp))
