"""Test helpers for the residual path (dequantisation, LFNST, inverse transforms, residual add)."""
import ctypes as C

import numpy as np

from openvvc_b200 import synth, synth_tu


def tr_case(w, h, seed, log2_ctu=7, ict_type=0, sh_ts_disabled=0, dep_quant=1, legal=False, p_intra=0.3, min_log2=2, p_bdpcm=0.0):
    """prediction planes + TU commands + coefficient arena for one seeded picture"""
    cmds, arena = synth_tu.tu_commands(w, h, seed, log2_ctu, ict_type=ict_type, sh_ts_disabled=sh_ts_disabled,
                                       dep_quant=dep_quant, legal=legal, p_intra=p_intra, min_log2=min_log2, p_bdpcm=p_bdpcm)
    return synth.planes(w, h, seed + 3), cmds, arena


def emu_residual(emu, planes, cmds, arena, dep_quant=1, sh_ts_disabled=0, ict_type=0, bitdepth=10):
    out = [p.copy() for p in planes]
    u16p = C.POINTER(C.c_uint16)
    dst = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in out])
    ds = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    n = emu.emu_residual_picture(dst, ds, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds),
                                 arena.ctypes.data_as(C.c_void_p), C.c_size_t(len(arena)), dep_quant, sh_ts_disabled, ict_type)
    assert n >= 0, "emu rejected command %d" % (-n - 1)
    return out


def gpu_residual(ctx, planes, cmds, arena, dep_quant=1, sh_ts_disabled=0, ict_type=0):
    from openvvc_b200 import api
    h, w = planes[0].shape
    pic = ctx.picture(w, h)
    pic.upload(planes)
    batch = api.TuBatch(ctx, max(1, len(cmds)), max(64, len(arena)))
    batch.upload(cmds, arena, dep_quant, sh_ts_disabled, ict_type)
    batch.run(pic)
    out = pic.download()
    batch.destroy(); pic.destroy()
    return out
