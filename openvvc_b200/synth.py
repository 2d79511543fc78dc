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
    """3 planar 4:2:0 uint16 planes. kind 'noise' = uniform; 'blocky' = smooth + block steps; 'natural' = smooth field + sigma-4 noise
    (keeps DBF/ALF decisions non-degenerate)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bitdepth) - 1
    out = []
    for (ph, pw) in plane_dims(w, h):
        if kind == "noise":
            p = rng.integers(0, maxv + 1, size=(ph, pw), dtype=np.uint16)
        elif kind == "blocky":
            # very smooth field + a DC step per 8x8 block and sigma-0.7 noise: what a coarsely quantised picture looks
            # like before deblocking (reaches the strong and long deblocking filters)
            f = _smooth(rng, (ph, pw), 48) * (maxv / 1023.0)
            steps = rng.integers(-6, 7, size=((ph + 7) // 8, (pw + 7) // 8)).astype(np.float64)
            f = f + np.kron(steps, np.ones((8, 8)))[:ph, :pw] + rng.normal(0, 0.7, size=(ph, pw))
            p = np.clip(np.rint(f), 0, maxv).astype(np.uint16)
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


# ---- inter prediction -----------------------------------------------------------------------------------
MC_CMD = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8),
                   ("inter_dir", np.uint8), ("flags", np.uint8), ("ref_idx0", np.int8), ("ref_idx1", np.int8),
                   ("bcw_idx_plus1", np.uint8), ("tool", np.uint8),
                   ("mv0_x", np.int32), ("mv0_y", np.int32), ("mv1_x", np.int32), ("mv1_y", np.int32),
                   ("aux", np.uint32)])


def partition(w, h, rng, log2_ctu=7, min_log2=2, max_log2=7, p_split=0.55):
    """Random QTBT-like partition of the picture into power-of-two blocks (x, y, log2_w, log2_h) that lie
    inside the picture.  Blocks never straddle a CTU; picture dimensions must be multiples of 8."""
    out = []
    ctu = 1 << log2_ctu

    def rec(x, y, lw, lh):
        bw, bh = 1 << lw, 1 << lh
        if x >= w or y >= h:
            return
        must = x + bw > w or y + bh > h or lw > max_log2 or lh > max_log2
        can_h = lw > min_log2
        can_v = lh > min_log2
        if must or ((can_h or can_v) and rng.random() < p_split):
            if must:
                sh = x + bw > w or lw > max_log2
                sv = y + bh > h or lh > max_log2
            else:
                mode = rng.integers(0, 3)
                sh = can_h and mode != 1
                sv = can_v and mode != 0
                if not sh and not sv:
                    sh, sv = can_h, can_v
            nlw, nlh = lw - (1 if sh else 0), lh - (1 if sv else 0)
            for dy in range(2 if sv else 1):
                for dx in range(2 if sh else 1):
                    rec(x + (dx << nlw), y + (dy << nlh), nlw, nlh)
        else:
            out.append((x, y, lw, lh))

    for cy in range(0, h, ctu):
        for cx in range(0, w, ctu):
            rec(cx, cy, log2_ctu, log2_ctu)
    return out


def mc_commands(w, h, seed=1004, log2_ctu=7, n_refs=2, p_bi=0.7, mv_range=64, min_log2=3):
    """PU list for one inter picture (seed 1004): random partition, MVs with an integer part in
    +-mv_range samples (some pointing outside the picture so border replication is exercised), uniformly
    random 1/16 fractions, 70 % bi-prediction, 5 % BCW != default, 5 % half-pel AMVR."""
    rng = np.random.default_rng(seed)
    blocks = partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(7, log2_ctu))
    cmds = np.zeros(len(blocks), dtype=MC_CMD)
    for i, (x, y, lw, lh) in enumerate(blocks):
        c = cmds[i]
        c["x0"], c["y0"], c["log2_w"], c["log2_h"] = x, y, lw, lh
        bi = rng.random() < p_bi and not (lw + lh < 6)       # bi-prediction needs >= 8x8
        c["inter_dir"] = 3 if bi else int(rng.integers(1, 3))
        # half-pel AMVR (not on 4x4 blocks: those only exist as affine sub-blocks, where it cannot be signalled)
        c["flags"] = 3 | (4 if (rng.random() < 0.05 and lw + lh > 4) else 0)
        c["ref_idx0"] = rng.integers(0, n_refs); c["ref_idx1"] = rng.integers(0, n_refs)
        if bi and rng.random() < 0.05:
            c["bcw_idx_plus1"] = int(rng.choice([1, 2, 4, 5]))
        kind = rng.random()
        for k in ("mv0", "mv1"):
            if kind < 0.1:        # far outside: exercises clip_mv and full border replication
                mx = int(rng.integers(-(w + 200), w + 200)) * 16; my = int(rng.integers(-(h + 200), h + 200)) * 16
            else:
                mx = int(rng.integers(-mv_range * 16, mv_range * 16 + 1)); my = int(rng.integers(-mv_range * 16, mv_range * 16 + 1))
                if rng.random() < 0.25:
                    mx &= ~15
                if rng.random() < 0.25:
                    my &= ~15
            if c["flags"] & 4:    # half-pel AMVR: MVs are multiples of 8
                mx &= ~7; my &= ~7
            c[k + "_x"], c[k + "_y"] = mx, my
        if bi and rng.random() < 0.03:   # identical motion -> falls back to uni-prediction from list 1
            c["mv1_x"], c["mv1_y"] = c["mv0_x"], c["mv0_y"]
    return cmds


MC_TOOL_PLAIN, MC_TOOL_BDOF, MC_TOOL_PROF, MC_TOOL_DMVR, MC_TOOL_GPM = 0, 1, 2, 3, 4


def mc_commands_tools(w, h, seed=1004, log2_ctu=7, n_refs=2, mv_range=32, min_log2=3, p_bdof=0.3, p_affine=0.2, p_dmvr=0.0, p_gpm=0.0):
    """PU list with BDOF and affine/PROF coding units mixed in (25 % of eligible bi blocks BDOF, 15 %
    affine 4x4).  Returns (cmds, prof_table): BDOF CUs become <=16x16 luma sub-block commands (tool 1) + one chroma
    command; affine CUs become 4x4 luma commands (tool 2 with PROF, or plain 6-tap) + 8x8-luma chroma commands with the
    averaged MVs (drv_affine_mvp.c:3380-3410); prof_table[n] = (dmv_scale_h_0, v_0, h_1, v_1) of affine CU n."""
    rng = np.random.default_rng(seed)
    blocks = partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(7, log2_ctu))
    out, prof = [], []

    def new(x, y, lw, lh, inter_dir, flags, r0, r1, mv0, mv1, tool=0, bcw=0, aux=0):
        c = np.zeros((), dtype=MC_CMD)
        c["x0"], c["y0"], c["log2_w"], c["log2_h"] = x, y, lw, lh
        c["inter_dir"], c["flags"], c["ref_idx0"], c["ref_idx1"], c["tool"], c["bcw_idx_plus1"], c["aux"] = inter_dir, flags, r0, r1, tool, bcw, aux
        c["mv0_x"], c["mv0_y"], c["mv1_x"], c["mv1_y"] = mv0[0], mv0[1], mv1[0], mv1[1]
        out.append(c)

    def rmv(far=False):
        if far:
            return int(rng.integers(-(w + 100), w + 100)) * 16 + int(rng.integers(0, 16)), int(rng.integers(-(h + 100), h + 100)) * 16 + int(rng.integers(0, 16))
        mx, my = int(rng.integers(-mv_range * 16, mv_range * 16 + 1)), int(rng.integers(-mv_range * 16, mv_range * 16 + 1))
        if rng.random() < 0.2:
            mx &= ~15
        if rng.random() < 0.2:
            my &= ~15
        return mx, my

    for (x, y, lw, lh) in blocks:
        r0, r1 = int(rng.integers(0, n_refs)), int(rng.integers(0, n_refs))
        far = rng.random() < 0.06
        mv0, mv1 = rmv(far), rmv(far)
        hpel = 4 if rng.random() < 0.05 else 0
        if hpel:
            mv0 = (mv0[0] & ~7, mv0[1] & ~7); mv1 = (mv1[0] & ~7, mv1[1] & ~7)
        kind = rng.random()
        if lw >= 3 and lh >= 3 and lw + lh >= 7 and kind >= 1.0 - p_dmvr:
            # DMVR CU (vcl_coding_unit.c:2600-2647): luma + chroma per <= 16x16 sub-block, the refinement starts from the
            # CU's MVs for every sub-block; MVs of the two lists roughly mirrored like true bi-prediction
            sw, sh = min(lw, 4), min(lh, 4)
            if rng.random() < 0.7:
                mv1 = (-mv0[0] + int(rng.integers(-40, 41)), -mv0[1] + int(rng.integers(-40, 41)))
                if hpel:
                    mv1 = (mv1[0] & ~7, mv1[1] & ~7)
            bd = int(rng.random() < 0.6)
            for oy in range(0, 1 << lh, 1 << sh):
                for ox in range(0, 1 << lw, 1 << sw):
                    new(x + ox, y + oy, sw, sh, 3, 3 | hpel, r0, r1, mv0, mv1, tool=MC_TOOL_DMVR, aux=bd)
        elif lw >= 3 and lh >= 3 and lw + lh >= 7 and kind < p_bdof:
            # BDOF CU (vcl_coding_unit.c:2648-2672): luma per <= 16x16 sub-block, chroma for the whole CU
            sw, sh = min(lw, 4), min(lh, 4)
            for oy in range(0, 1 << lh, 1 << sh):
                for ox in range(0, 1 << lw, 1 << sw):
                    new(x + ox, y + oy, sw, sh, 3, 1 | hpel, r0, r1, mv0, mv1, tool=MC_TOOL_BDOF)
            new(x, y, lw, lh, 3, 2 | hpel, r0, r1, mv0, mv1)
        elif lw >= 4 and lh >= 4 and kind < p_bdof + p_affine:
            # affine CU: one MV per 4x4 from a random 6-parameter model, PROF on the lists flagged in prof_dir
            inter_dir = int(rng.choice([1, 2, 3]))
            prof_dir = 0 if rng.random() < 0.25 else (inter_dir if rng.random() < 0.7 else int(rng.integers(1, 4)) & inter_dir)
            if inter_dir != 3 and prof_dir:
                prof_dir = inter_dir
            bcw = int(rng.choice([1, 2, 4, 5])) if (inter_dir == 3 and rng.random() < 0.1) else 0
            mdl = [(rmv(far), rng.integers(-6, 7, size=4)) for _ in range(2)]
            prof.append(rng.integers(-32, 32, size=(4, 16)).astype(np.int16))
            nw, nh = (1 << lw) >> 2, (1 << lh) >> 2
            mvs = [[[(int(b[0] + a[0] * i + a[1] * j), int(b[1] + a[2] * i + a[3] * j)) for i in range(nw)] for j in range(nh)] for (b, a) in mdl]
            for j in range(nh):
                for i in range(nw):
                    new(x + 4 * i, y + 4 * j, 2, 2, inter_dir, 1, r0, r1, mvs[0][j][i], mvs[1][j][i],
                        tool=MC_TOOL_PROF if prof_dir else MC_TOOL_PLAIN, bcw=bcw, aux=(prof_dir | ((len(prof) - 1) << 8)) if prof_dir else 0)
            for j in range(0, nh, 2):
                for i in range(0, nw, 2):
                    avg = []
                    for l in range(2):
                        ax = mvs[l][j][i][0] + mvs[l][j + 1][i + 1][0]; ay = mvs[l][j][i][1] + mvs[l][j + 1][i + 1][1]
                        ax += ax < 0; ay += ay < 0
                        avg.append((ax >> 1, ay >> 1))
                    new(x + 4 * i, y + 4 * j, 3, 3, inter_dir, 2, r0, r1, avg[0], avg[1], bcw=bcw)
        elif p_gpm and 3 <= lw <= 6 and 3 <= lh <= 6 and abs(lw - lh) < 3 and rng.random() < p_gpm:
            # GPM CU (rcn_gpm_b): two uni-predictions from any list, one of the 64 splits; MVs are merge candidates, no half-pel AMVR
            new(x, y, lw, lh, 3, 3, r0, r1, mv0, mv1, tool=MC_TOOL_GPM,
                aux=int(rng.integers(0, 64)) | (int(rng.integers(0, 2)) << 8) | (int(rng.integers(0, 2)) << 9))
        else:
            bi = rng.random() < 0.7 and lw + lh >= 6
            new(x, y, lw, lh, 3 if bi else int(rng.integers(1, 3)), 3 | hpel, r0, r1, mv0, mv1,
                bcw=int(rng.choice([1, 2, 4, 5])) if (bi and rng.random() < 0.05) else 0)
    cmds = np.array(out, dtype=MC_CMD)
    prof_tab = np.stack(prof) if prof else np.zeros((1, 4, 16), np.int16)
    return cmds, np.ascontiguousarray(prof_tab, dtype=np.int16)


def lmcs_data(seed=1007):
    rng = np.random.default_rng(seed)
    min_idx = int(rng.integers(0, 4)); dmax = int(rng.integers(0, 4))
    return dict(min_bin_idx=min_idx, delta_max_bin_idx=dmax,
                delta_abs_cw=rng.integers(0, 24, size=16), delta_sign_cw=rng.integers(0, 2, size=16),
                delta_abs_crs=int(rng.integers(0, 8)), delta_sign_crs=int(rng.integers(0, 2)))
