"""Seeded synthetic mixed pictures for the reconstruction order (ovgpu_recon_picture): one partition whose coding units are
intra (luma + chroma, optional CCLM / MRL), CIIP or inter (their prediction is already in the picture), each with an
optional transform unit -- so intra blocks read neighbours that hold prediction + residual, like in a real picture."""
import numpy as np

from . import synth, synth_intra, synth_tu


def recon_case(w, h, seed, log2_ctu=7, min_log2=2, p_inter=0.4, p_ciip=0.1, p_mrl=0.15, p_cclm=0.3, collocated=False, ict_type=0,
               p_cbf=0.6, chroma=True, p_mip=0.0, p_isp=0.0, p_bdpcm=0.0, p_ibc=0.0):
    """Returns (intra cmds incl. flags == 0 placeholders, tu cmds, arena); cmds[i] and tus[i] are the same coding unit.
    With p_isp: also (isp words, isp transform blocks [n_isp][4]); the luma residual of an ISP coding unit lives there (its
    tus[i] keeps the chroma blocks only) and the coefficients share the arena."""
    rng = np.random.default_rng(seed)
    blocks = synth.partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(6, log2_ctu), p_split=0.5)
    cmds = synth_intra.intra_commands(w, h, seed + 1, log2_ctu, chroma=chroma, p_ciip=p_ciip, p_mrl=p_mrl, p_cclm=p_cclm,
                                      collocated=collocated, p_inter=p_inter, blocks=blocks, p_mip=p_mip, p_isp=p_isp, p_ibc=p_ibc)
    words = isp_tus = isp_arena = None
    if p_isp:
        cmds, words, isp_tus, isp_arena = synth_intra.isp_tables(cmds, seed + 3)
    cmds = synth_intra.availability(w, h, cmds, log2_ctu, isp_words=words)
    is_intra = (cmds["flags"] != 0) & ((cmds["flags"] & synth_intra.INTRA_CIIP) == 0)
    tus, arena = synth_tu.tu_commands(w, h, seed + 2, log2_ctu, ict_type=ict_type, p_cbf=p_cbf, legal=True, blocks=blocks,
                                      intra_mask=is_intra, p_bdpcm=p_bdpcm)
    assert len(tus) == len(cmds)
    if p_isp:
        tus["cbf_mask"][(cmds["flags"] & synth_intra.INTRA_ISP) != 0] &= 0x0f
        isp_tus["coeff_off"][..., 2] += len(arena)
        return np.ascontiguousarray(cmds), tus, np.concatenate([arena, isp_arena]), words, isp_tus
    return np.ascontiguousarray(cmds), tus, arena


LMCS_ANCHOR = np.dtype([("x0", np.int16), ("y0", np.int16), ("abv_mask", np.uint16), ("lft_mask", np.uint16)])   # == OVGPULmcsAnchor


def lmcs_anchors(cmds, tus, log2_ctu=7):
    """What the recorder lists for the on-device LMCS chroma scale: the coding units that start on a 64-aligned position
    (coding_unit, vcl_coding_unit.c:725-730) with the two 16-bit availability masks rcn_lmcs_compute_chroma_scale reads there
    (rcn_lmcs.c:325-331), and for every transform unit the anchor in force (flag OVGPU_TU_SCALE_ANCHOR + scale_anchor)."""
    cm = (1 << log2_ctu) - 1
    anchors = []
    tus = tus.copy()
    for i, c in enumerate(cmds):
        x0, y0 = int(c["x0"]), int(c["y0"])
        if not (x0 & 63) and not (y0 & 63):
            abv = (int(c["avail_rows"]) >> (((x0 & cm) >> 2) + 1)) & 0xffff
            lft = (int(c["avail_cols"]) >> (((y0 & cm) >> 2) + 1)) & 0xffff
            anchors.append((x0, y0, abv, lft))
        assert anchors, "the first coding unit of a picture starts at (0, 0)"
        tus["flags"][i] |= synth_tu.TU_SCALE_ANCHOR
        tus["scale_anchor"][i] = len(anchors) - 1
    return np.array(anchors, dtype=LMCS_ANCHOR), tus
