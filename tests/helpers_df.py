"""Test helpers for the deblocking filter."""
import ctypes as C

import numpy as np

from openvvc_b200 import synth, synth_dbf


def dbf_case(w, h, seed, log2_ctu=7, kind="blocky", beta_offset=0, tc_offset=0, qp_lo=26, qp_hi=45, min_log2=2):
    blks = synth_dbf.dbf_blocks(w, h, seed, log2_ctu, min_log2=min_log2, qp_lo=qp_lo, qp_hi=qp_hi)
    maps = synth_dbf.dbf_maps(w, h, blks, log2_ctu, beta_offset, tc_offset)
    return synth.planes(w, h, seed + 1, kind), blks, maps


def emu_dbf(emu, planes, maps, bitdepth=10):
    from openvvc_b200 import api
    out = [p.copy() for p in planes]
    u16p = C.POINTER(C.c_uint16)
    dst = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in out])
    ds = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    st = api.DbfMaps.from_dict(maps)
    emu.emu_dbf_picture(dst, ds, bitdepth, C.byref(st))
    return out


def gpu_dbf(ctx, planes, maps, log2_ctu=7):
    from openvvc_b200 import api
    h, w = planes[0].shape
    pic = ctx.picture(w, h)
    pic.upload(planes)
    prm = api.DbfParams(ctx, w, h, log2_ctu)
    prm.upload(maps)
    prm.run(pic)
    out = pic.download()
    prm.destroy(); pic.destroy()
    return out
