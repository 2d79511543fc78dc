"""ctypes bindings of the deblocking oracle (TEST INFRASTRUCTURE ONLY, see pyoracle.py)."""
import ctypes as C

import numpy as np

from . import pyoracle as po
from openvvc_b200.synth_dbf import DBF_BLK, NB_MAPS

u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class DbfMaps(C.Structure):
    """OVODbfMaps == OVGPUDbfMaps"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("log2_ctu_s", C.c_int32),
                ("beta_offset", C.c_int16), ("tc_offset", C.c_int16),
                ("ver", u32p * NB_MAPS), ("hor", u32p * NB_MAPS), ("qp", u8p * 3)]


def maps_struct(m):
    """dict from synth_dbf.dbf_maps (or empty_maps) -> (struct, keepalive)"""
    s = DbfMaps()
    s.width, s.height, s.log2_ctu_s = m["width"], m["height"], m["log2_ctu"]
    s.beta_offset, s.tc_offset = m["beta_offset"], m["tc_offset"]
    for k in range(NB_MAPS):
        s.ver[k] = m["ver"][k].ctypes.data_as(u32p)
        s.hor[k] = m["hor"][k].ctypes.data_as(u32p)
    for k in range(3):
        s.qp[k] = m["qp"][k].ctypes.data_as(u8p)
    return s, m


def empty_maps(w, h, log2_ctu=7, beta_offset=0, tc_offset=0):
    n_u = (1 << log2_ctu) >> 2
    w_u, h_u = w >> 2, h >> 2
    ncw, nch = -(-w_u // n_u), -(-h_u // n_u)
    return dict(width=w, height=h, log2_ctu=log2_ctu, beta_offset=beta_offset, tc_offset=tc_offset,
                ver=np.zeros((NB_MAPS, nch, w_u + 1), np.uint32), hor=np.zeros((NB_MAPS, ncw, h_u + 1), np.uint32),
                qp=np.zeros((3, h_u, w_u), np.uint8))


def dbf_picture_ref(planes, blks, log2_ctu=7, beta_offset=0, tc_offset=0, simd=1, record=False, bitdepth=10):
    """The reference's rcn_dbf_ctu over the whole picture, in place. Returns the recorded picture-level maps if asked."""
    h, w = planes[0].shape
    blks = np.ascontiguousarray(blks, dtype=DBF_BLK)
    dst, ds = po._planes_arg(planes)
    rec = empty_maps(w, h, log2_ctu, beta_offset, tc_offset) if record else None
    st = maps_struct(rec)[0] if record else None
    po.ref().ovref_dbf_picture(dst, ds, w, h, log2_ctu, bitdepth, simd, blks.ctypes.data_as(C.c_void_p), len(blks),
                               beta_offset, tc_offset, C.byref(st) if record else None)
    return rec


def dbf_picture_port(planes, maps, bitdepth=10):
    dst, ds = po._planes_arg(planes)
    st, _ = maps_struct(maps)
    po.port().ovo_dbf_picture(dst, ds, bitdepth, C.byref(st))
    return planes
