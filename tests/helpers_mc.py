"""Test helpers for motion compensation / LMCS cases."""
import ctypes as C

import numpy as np

from openvvc_b200 import synth


def mc_case(po, w, h, seed, log2_ctu=7, lmcs=False, min_log2=3, n_refs=2):
    refs0 = [synth.planes(w, h, seed + 10 + i) for i in range(n_refs)]
    refs1 = [synth.planes(w, h, seed + 20 + i) for i in range(n_refs)]
    pocs = [[8 * (i + 1) for i in range(n_refs)], [8 * (i + 2) for i in range(n_refs)]]   # poc0[1] == poc1[0]
    cmds = synth.mc_commands(w, h, seed, log2_ctu, n_refs=n_refs, min_log2=min_log2).astype(po.McCmd)
    luts = po.lmcs_build_luts(synth.lmcs_data(seed), impl="port") if lmcs else None
    return refs0, refs1, pocs, cmds, luts


def blank(w, h):
    return [np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16)]


def mc_tools_case(po, w, h, seed, log2_ctu=7, lmcs=False, min_log2=3, n_refs=2, p_gpm=0.0):
    """like mc_case, with BDOF and affine/PROF coding units mixed in; returns the PROF table as well"""
    refs0 = [synth.planes(w, h, seed + 10 + i) for i in range(n_refs)]
    refs1 = [synth.planes(w, h, seed + 20 + i) for i in range(n_refs)]
    pocs = [[8 * (i + 1) for i in range(n_refs)], [8 * (i + 2) for i in range(n_refs)]]
    cmds, prof = synth.mc_commands_tools(w, h, seed, log2_ctu, n_refs=n_refs, min_log2=min_log2, p_gpm=p_gpm)
    luts = po.lmcs_build_luts(synth.lmcs_data(seed), impl="port") if lmcs else None
    return refs0, refs1, pocs, cmds.astype(po.McCmd), luts, prof


def emu_mc(emu, po, w, h, refs0, refs1, pocs, cmds, luts, bitdepth=10, prof=None, dmvr_out=None):
    out = blank(w, h)
    refs, keep = po.make_mc_refs([refs0, refs1], pocs)
    dst, ds = po._planes_arg(out)
    lut = luts.np("fwd_lut").ctypes.data_as(po.u16p) if luts is not None else None
    prof = np.ascontiguousarray(prof, dtype=np.int16) if prof is not None else None
    n = emu.emu_mc_picture(dst, ds, w, h, bitdepth, C.byref(refs), cmds.ctypes.data_as(C.c_void_p), len(cmds), lut,
                           prof.ctypes.data_as(C.c_void_p) if prof is not None else None, len(prof) if prof is not None else 0,
                           dmvr_out.ctypes.data_as(C.c_void_p) if dmvr_out is not None else None)
    assert n >= 0, "emu rejected command %d" % (-n - 1)
    return out


def mc_dmvr_case(po, w, h, seed, log2_ctu=7, lmcs=False, kind="natural"):
    """PU list with DMVR / BDOF / affine CUs; list-1 references are shifted noisy copies of list 0 so that the SAD search
    finds real minima.  Returns (refs0, refs1, pocs, cmds, luts, prof)."""
    refs0 = [synth.planes(w, h, seed + 10 + i, kind) for i in range(2)]
    rng = np.random.default_rng(seed)
    refs1 = [[np.clip(np.roll(p, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), axis=(0, 1)).astype(int) + rng.integers(-3, 4, size=p.shape),
                      0, 1023).astype(np.uint16) for p in refs0[i]] for i in range(2)]
    pocs = [[8, 16], [16, 24]]
    cmds, prof = synth.mc_commands_tools(w, h, seed, log2_ctu, p_dmvr=0.4)
    luts = po.lmcs_build_luts(synth.lmcs_data(seed), impl="port") if lmcs else None
    return refs0, refs1, pocs, cmds.astype(po.McCmd), luts, prof


def gpu_mc(ctx, w, h, refs0, refs1, pocs, cmds, luts, prof=None, dmvr_out=None):
    from openvvc_b200 import api
    dst = ctx.picture(w, h)
    dst.upload(blank(w, h))
    pics = [[], []]
    for l, lst in enumerate((refs0, refs1)):
        for planes in lst:
            p = ctx.picture(w, h); p.upload(planes); pics[l].append(p)
    batch = api.McBatch(ctx, len(cmds), api.mc_nb_items(cmds))
    batch.upload(cmds.astype(api.MC_CMD), prof=prof)
    lm = None
    if luts is not None:
        gl = api.LmcsLuts()
        C.memmove(C.byref(gl), C.byref(luts), C.sizeof(gl))
        lm = api.Lmcs(ctx, gl)
    batch.run(dst, pics, pocs, lm)
    batch.status()                      # the command walk runs on the device: this is where a refused batch is reported
    if dmvr_out is not None:
        dmvr_out[:] = batch.read_dmvr(len(dmvr_out))
    out = dst.download()
    batch.destroy(); dst.destroy()
    if lm:
        lm.destroy()
    for lst in pics:
        for p in lst:
            p.destroy()
    return out


# reference picture resampling: reference sizes as (num, den) of the picture size, per list -- 2x, 1.5x (the two ratios VTM's
# RPR configurations use), same size, and small references that are up-sampled
RPR_RATIOS = [[(2, 1), (1, 1), (3, 2)], [(1, 2), (1, 1), (1, 4)]]


def rpr_case(po, w, h, seed, log2_ctu=7, lmcs=False, kind="natural", hor_col=1, ver_col=0, p_gpm=0.15, p_affine=0.2):
    """PU list (plain uni / bi / BCW, affine 4x4 with PROF, GPM) over references of their own sizes.
    Returns (refs0, refs1, pocs, cmds, luts, prof, scaling)."""
    def size(r):
        return (w * r[0] // r[1]) & ~7, (h * r[0] // r[1]) & ~7
    refs0 = [synth.planes(*size(r), seed + 10 + i, kind) for i, r in enumerate(RPR_RATIOS[0])]
    refs1 = [synth.planes(*size(r), seed + 20 + i, kind) for i, r in enumerate(RPR_RATIOS[1])]
    pocs = [[8, 16, 24], [16, 24, 32]]
    cmds, prof = synth.mc_commands_tools(w, h, seed, log2_ctu, n_refs=3, p_bdof=0.0, p_affine=p_affine, p_gpm=p_gpm)
    luts = po.lmcs_build_luts(synth.lmcs_data(seed), impl="port") if lmcs else None
    return refs0, refs1, pocs, cmds.astype(po.McCmd), luts, prof, po.make_mc_scaling(w, h, [refs0, refs1], hor_col, ver_col)


def emu_mc_rpr(emu, po, w, h, refs0, refs1, pocs, cmds, luts, sc, bitdepth=10, prof=None):
    out = blank(w, h)
    refs, keep = po.make_mc_refs([refs0, refs1], pocs)
    dst, ds = po._planes_arg(out)
    lut = luts.np("fwd_lut").ctypes.data_as(po.u16p) if luts is not None else None
    prof = np.ascontiguousarray(prof, dtype=np.int16) if prof is not None else None
    n = emu.emu_mc_picture_rpr(dst, ds, w, h, bitdepth, C.byref(refs), C.byref(sc), cmds.ctypes.data_as(C.c_void_p), len(cmds), lut,
                               prof.ctypes.data_as(C.c_void_p) if prof is not None else None, len(prof) if prof is not None else 0, None)
    assert n >= 0, "emu rejected the batch (%d)" % n
    return out


def gpu_mc_rpr(ctx, w, h, refs0, refs1, pocs, cmds, luts, sc, prof=None):
    """the C ABI with references of their own sizes: ovgpu_mc_batch_set_scaling before the upload, then ovgpu_mc_picture"""
    from openvvc_b200 import api
    dst = ctx.picture(w, h)
    dst.upload(blank(w, h))
    pics = [[], []]
    for l, lst in enumerate((refs0, refs1)):
        for planes in lst:
            rh, rw = planes[0].shape
            p = ctx.picture(rw, rh); p.upload(planes); pics[l].append(p)
    batch = api.McBatch(ctx, len(cmds), api.mc_nb_items(cmds))
    batch.set_scaling([[(sc.scale[l][i][0], sc.scale[l][i][1]) for i in range(len(lst))] for l, lst in enumerate((refs0, refs1))],
                      sc.chroma_hor_col, sc.chroma_ver_col)
    batch.upload(cmds.astype(api.MC_CMD), prof=prof)
    lm = None
    if luts is not None:
        gl = api.LmcsLuts()
        C.memmove(C.byref(gl), C.byref(luts), C.sizeof(gl))
        lm = api.Lmcs(ctx, gl)
    try:
        batch.run(dst, pics, pocs, lm)
        batch.status()
        out = dst.download()
    finally:
        batch.destroy(); dst.destroy()
        if lm:
            lm.destroy()
        for lst in pics:
            for p in lst:
                p.destroy()
    return out
