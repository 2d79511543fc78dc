"""Seeded synthetic workloads for the rcn_* hot path.

No VVC bitstreams exist offline, so every stage is exercised with decoder state that a parser *could*
have produced: planes, per-CTU SAO/ALF parameters, APS filter data, PU/MV lists, coefficient lists.
Everything is generated with numpy.random.default_rng(seed) so tests, bench.py and the oracle see the
same bytes.
"""
import numpy as np


def plane_dims(w, h):
    return [(h, w), (h // 2, w // 2), (h // 2, w // 2)]


def nb_ctus(w, h, log2_ctu=7):
    s = 1 << log2_ctu
    return (w + s - 1) >> log2_ctu, (h + s - 1) >> log2_ctu


def _smooth(rng, shape, scale):
    """Low-passed random field: random values on a coarse grid, bilinearly up-sampled."""
    h, w = shape
    gh, gw = h // scale + 2, w // scale + 2
    g = rng.uniform(0, 1023, size=(gh, gw))
    ys = np.arange(h) / scale
    xs = np.arange(w) / scale
    y0 = ys.astype(int); x0 = xs.astype(int)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def planes(w, h, seed, kind="natural", bitdepth=10):
    """3 planar 4:2:0 uint16 planes. kind 'noise' = uniform; 'natural' = smooth field + sigma-4 noise
    (keeps DBF/ALF decisions non-degenerate)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bitdepth) - 1
    out = []
    for (ph, pw) in plane_dims(w, h):
        if kind == "noise":
            p = rng.integers(0, maxv + 1, size=(ph, pw), dtype=np.uint16)
        else:
            f = _smooth(rng, (ph, pw), 16) * (maxv / 1023.0)
            f = f + rng.normal(0, 4, size=(ph, pw))
            p = np.clip(np.rint(f), 0, maxv).astype(np.uint16)
        out.append(np.ascontiguousarray(p))
    return out


def alf_aps(rng, clip=True):
    """One ALF APS worth of data, field names after OVALFData (nvcl_structures.h:649-669)."""
    nf = int(rng.integers(1, 26))
    return dict(
        luma_clip_flag=int(clip and rng.integers(0, 2)),
        num_luma_filters=nf,
        luma_coeff_delta_idx=rng.integers(0, nf, size=25),
        luma_coeff=rng.integers(-64, 64, size=(25, 13)),
        luma_clip_idx=rng.integers(0, 4, size=(25, 13)),
        chroma_clip_flag=int(clip and rng.integers(0, 2)),
        num_chroma_alts=int(rng.integers(1, 9)),
        chroma_coeff=rng.integers(-64, 64, size=(8, 7)),
        chroma_clip_idx=rng.integers(0, 4, size=(8, 7)),
    )


def alf_params(w, h, seed=1001, log2_ctu=7, p_on=0.9):
    """Per-CTU ALF/CC-ALF parameters + APS list (seed 1001 per SURVEY §8(d))."""
    rng = np.random.default_rng(seed)
    nw, nh = nb_ctus(w, h, log2_ctu)
    n = nw * nh
    n_aps = int(rng.integers(1, 9))
    luma_aps = [alf_aps(rng) for _ in range(n_aps)]
    chroma_aps = alf_aps(rng)
    flag = ((rng.random(n) < p_on) * 4 + (rng.random(n) < p_on) * 2 + (rng.random(n) < p_on)).astype(np.uint8)
    ctu = dict(
        flag=flag,
        idx=rng.integers(0, 16 + n_aps, size=n).astype(np.uint8),
        cb_alt=rng.integers(0, chroma_aps["num_chroma_alts"], size=n).astype(np.uint8),
        cr_alt=rng.integers(0, chroma_aps["num_chroma_alts"], size=n).astype(np.uint8),
        cc_cb=rng.integers(0, 5, size=n).astype(np.uint8),
        cc_cr=rng.integers(0, 5, size=n).astype(np.uint8),
        cc_cb_enabled=1, cc_cr_enabled=1,
    )
    cc_coeff = np.zeros((2, 4, 8), dtype=np.int16)
    # CC-ALF coefficients are signed powers of two up to 64 (nvcl_nal_aps.c:128-155)
    mag = rng.integers(0, 8, size=(2, 4, 7))
    sign = rng.integers(0, 2, size=(2, 4, 7))
    cc_coeff[:, :, :7] = np.where(mag == 0, 0, (1 << np.maximum(mag - 1, 0)) * np.where(sign, -1, 1))
    return dict(luma_aps=luma_aps, chroma_aps=chroma_aps, ctu=ctu, cc_coeff=cc_coeff)


def sao_params(w, h, seed=1002, log2_ctu=7, sao_dtype=None):
    """Per-CTU SAO parameters (SAOParamsCtu, dec_structures.h:263-274); seed 1002."""
    rng = np.random.default_rng(seed)
    nw, nh = nb_ctus(w, h, log2_ctu)
    n = nw * nh
    dt = sao_dtype or np.dtype([("offset_val", np.int16, (3, 5)), ("band_position", np.uint8, 3),
                                ("eo_class", np.uint8, 3), ("type_idx", np.uint8, 3), ("pad", np.uint8)])
    p = np.zeros(n, dtype=dt)
    for c in range(3):
        t = rng.choice([0, 1, 2], p=[0.3, 0.2, 0.5], size=n)
        if c == 2:
            t = p["type_idx"][:, 1].copy()          # Cb/Cr share the type (vcl_sao.c:118)
        p["type_idx"][:, c] = t
        p["band_position"][:, c] = rng.integers(0, 32, size=n)
        p["eo_class"][:, c] = rng.integers(0, 4, size=n) if c < 2 else p["eo_class"][:, 1]
        a = rng.integers(0, 32, size=(n, 4))
        band = np.where(rng.integers(0, 2, size=(n, 4)) == 1, -a, a)
        edge = np.stack([a[:, 0], a[:, 1], np.zeros(n, dtype=a.dtype), -a[:, 2], -a[:, 3]], axis=1)
        ov = np.zeros((n, 5), dtype=np.int16)
        ov[:, :4] = band
        ov = np.where((t == 2)[:, None], edge, ov)
        p["offset_val"][:, c, :] = ov
    return p
