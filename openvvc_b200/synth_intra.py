"""Seeded synthetic intra coding units (SURVEY.md §8(d) "Intra", seed 1006): luma blocks in decode order (CTU by CTU,
z-order inside a CTU) with modes U[0,66].  The availability words of OVGPUIntraCmd are filled by the recorder
(the reference harness in tests; `availability()` below reproduces them for the bench)."""
import numpy as np

from . import synth

INTRA_CMD = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8), ("mode", np.uint8),
                      ("mode_c", np.uint8), ("flags", np.uint8), ("mrl_idx", np.uint8), ("ciip_wt", np.uint8), ("pad", np.uint8, 5),
                      ("avail_cols", np.uint64), ("avail_rows", np.uint64)])
assert INTRA_CMD.itemsize == 32

INTRA_LUMA, INTRA_CHROMA, INTRA_CIIP, INTRA_CCLM_COLLOCATED, INTRA_MIP = 1, 2, 4, 8, 16


def intra_commands(w, h, seed=1006, log2_ctu=7, min_log2=2, max_log2=6, chroma=False, p_ciip=0.0, p_mrl=0.0, p_cclm=0.0,
                   collocated=False, p_inter=0.0, blocks=None, p_mip=0.0):
    """p_inter: fraction of coding units that are NOT intra (flags = 0 placeholders): they keep the picture's samples -- the
    already reconstructed inter CUs of a real picture -- and only take part in the availability replay.  Without them an
    all-intra picture without residual degenerates to the mid value everywhere.  The placeholders are for the oracle
    harness (which has to mark them in the reference's progress field); drop them before handing commands to the GPU."""
    rng = np.random.default_rng(seed)
    part = blocks if blocks is not None else synth.partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(max_log2, log2_ctu), p_split=0.5)
    cmds = np.zeros(len(part), dtype=INTRA_CMD)
    for i, (x, y, lw, lh) in enumerate(part):
        c = cmds[i]
        c["x0"], c["y0"], c["log2_w"], c["log2_h"] = x, y, lw, lh
        r = rng.random()
        if p_inter and rng.random() < p_inter:
            continue
        c["flags"] = INTRA_LUMA
        c["mode"] = 0 if r < 0.15 else (1 if r < 0.3 else (int(rng.choice([18, 50, 2, 34, 66])) if r < 0.4 else int(rng.integers(2, 67))))
        if chroma:
            # single tree: Cb / Cr blocks of half the size, mode = DM (the luma mode) or an explicit planar / DC / ver / hor / random
            r = rng.random()
            mc = int(c["mode"]) if r < 0.4 else (0 if r < 0.5 else (1 if r < 0.6 else (int(rng.choice([18, 50, 2, 34, 66])) if r < 0.7 else int(rng.integers(2, 67)))))
            c["flags"] |= INTRA_CHROMA
            c["mode_c"] = mc
            if p_cclm and rng.random() < p_cclm:                   # CCLM / MDLM left / MDLM top (30 % of chroma)
                c["mode_c"] = int(rng.integers(67, 70))
                if collocated:
                    c["flags"] |= INTRA_CCLM_COLLOCATED
        if p_ciip and lw + lh >= 6 and rng.random() < p_ciip:       # CIIP: CUs of >= 64 samples (4x16 has the narrow-chroma rule)
            c["flags"] |= INTRA_CIIP
            c["ciip_wt"] = int(rng.integers(1, 4))
        if p_mrl and (y & ((1 << log2_ctu) - 1)) and not (c["flags"] & INTRA_CIIP) and rng.random() < p_mrl:
            c["mrl_idx"] = int(rng.integers(1, 3))                 # not on the first row of a CTU (vcl_coding_unit.c:1377)
        if p_mip and not (c["flags"] & INTRA_CIIP) and not c["mrl_idx"] and rng.random() < p_mip:
            # MIP (10 % of intra CUs): `mode` carries cu_opaque = transposed << 7 | matrix index
            nm = 16 if (lw == 2 and lh == 2) else (8 if (lw == 2 or lh == 2 or (lw <= 3 and lh <= 3)) else 6)
            c["flags"] |= INTRA_MIP
            c["mode"] = (int(rng.integers(0, 2)) << 7) | int(rng.integers(0, nm))
    return cmds


def availability(w, h, cmds, log2_ctu=7):
    """avail_cols / avail_rows of every command = progress_field.vfield[x0 >> 2] / .hfield[y0 >> 2] at call time
    (init_ctu_bitfield(_border), drv_utils.h:110-260; ctu_field_set_rect_bitfield :46-72), replayed from the decode order."""
    ctu = 1 << log2_ctu
    n_u = ctu >> 2
    ncw, nch = -(-w // ctu), -(-h // ctu)
    out = cmds.copy()
    cur = None
    for c in out:
        cx, cy = int(c["x0"]) >> log2_ctu, int(c["y0"]) >> log2_ctu
        if cur != (cx, cy):
            cur = (cx, cy)
            cw_u, ch_u = min(ctu, w - cx * ctu) >> 2, min(ctu, h - cy * ctu) >> 2
            lft, abv, tr = int(cx > 0), int(cy > 0), int(cy > 0 and cx + 1 < ncw)
            hf = [0] * (n_u + 2)
            vf = [0] * (n_u + 2)
            full = cw_u == n_u and ch_u == n_u
            if full:
                for i in range(1, n_u + 1):
                    hf[i] = lft
                    vf[i] = abv
                inner = ((1 << n_u) - 1) << 1
                hf[0] = inner if abv else 0
                vf[0] = inner if lft else 0
            else:
                for i in range(1, ch_u + 1):
                    hf[i] = lft
                for i in range(1, cw_u + 1):
                    vf[i] = abv
                hf[0] = (((1 << cw_u) - 1) << 1) if abv else 0
                vf[0] = (((1 << ch_u) - 1) << 1) if lft else 0
            hf[0] |= abv & lft
            vf[0] |= abv & lft
            if tr:
                hf[0] |= ((1 << n_u) - 1) << (n_u + 1)
            if cx + 2 == ncw and (w & (ctu - 1)):
                hf[0] &= (1 << ((((w & (ctu - 1)) + ctu) >> 2) + 1)) - 1
        x0, y0 = int(c["x0"]) & (ctu - 1), int(c["y0"]) & (ctu - 1)
        c["avail_cols"] = np.uint64(vf[x0 >> 2] & 0xFFFFFFFFFFFFFFFF)
        c["avail_rows"] = np.uint64(hf[y0 >> 2] & 0xFFFFFFFFFFFFFFFF)
        xu, yu, nw, nh = x0 >> 2, y0 >> 2, (1 << int(c["log2_w"])) >> 2, (1 << int(c["log2_h"])) >> 2
        for i in range(1, nh + 1):
            hf[yu + i] |= ((1 << nw) - 1) << (1 + xu)
        for i in range(1, nw + 1):
            vf[xu + i] |= ((1 << nh) - 1) << (1 + yu)
    return out
