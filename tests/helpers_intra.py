"""Test helpers for intra prediction."""
import ctypes as C

import numpy as np

from openvvc_b200 import synth, synth_intra


def intra_case(w, h, seed, log2_ctu=7, min_log2=2, kind="natural", chroma=False, p_ciip=0.0, p_mrl=0.0, p_cclm=0.0, collocated=False, p_inter=0.3, p_mip=0.0, p_ibc=0.0):
    """planes + intra commands in decode order with the availability words replayed from that order; commands with
    flags == 0 are non-intra placeholders (see synth_intra.intra_commands) that emu_intra / gpu_intra drop"""
    cmds = synth_intra.availability(w, h, synth_intra.intra_commands(w, h, seed, log2_ctu, min_log2=min_log2, chroma=chroma, p_ciip=p_ciip, p_mrl=p_mrl, p_cclm=p_cclm, collocated=collocated, p_inter=p_inter, p_mip=p_mip, p_ibc=p_ibc), log2_ctu)
    return synth.planes(w, h, seed + 1, kind), np.ascontiguousarray(cmds)


def emu_intra(emu, planes, cmds, log2_ctu=7, bitdepth=10, keep_placeholders=False):
    if not keep_placeholders:
        cmds = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    out = [p.copy() for p in planes]
    h, w = out[0].shape
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    nl = emu.emu_intra_picture(ptrs, strides, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds))
    assert nl >= 0, "emu rejected command %d" % (-nl - 1)
    return out, nl


def gpu_intra(ctx, planes, cmds, log2_ctu=7):
    cmds = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    from openvvc_b200 import api
    h, w = planes[0].shape
    pic = ctx.picture(w, h)
    pic.upload(planes)
    b = api.IntraBatch(ctx, max(1, len(cmds)))
    b.upload(cmds, w, h, log2_ctu)
    nl = b.levels()
    b.run(pic)
    out = pic.download()
    b.destroy(); pic.destroy()
    return out, nl


def isp_case(w, h, seed, log2_ctu=7, p_isp=0.5, p_inter=0.3):
    """planes, intra commands (single tree, ISP on half of the eligible luma CUs) with availability words, the ISP table words,
    the partitions' transform blocks [n_isp][4] and the coefficient arena"""
    cmds = synth_intra.intra_commands(w, h, seed, log2_ctu, chroma=True, p_inter=p_inter, p_isp=p_isp, p_mrl=0.2, p_mip=0.1)
    cmds, words, tus, arena = synth_intra.isp_tables(cmds, seed + 7)
    cmds = synth_intra.availability(w, h, cmds, log2_ctu, isp_words=words)
    return synth.planes(w, h, seed + 1), np.ascontiguousarray(cmds), words, tus, arena


def ibc_case(w, h, seed, log2_ctu=7, chroma=True, p_ibc=0.35):
    """planes + single-tree commands in decode order, a third of them intra block copies (block vector in pad[0..3])"""
    return intra_case(w, h, seed, log2_ctu, chroma=chroma, p_cclm=0.3 if chroma else 0.0, p_mrl=0.1, p_inter=0.3, p_ibc=p_ibc)
