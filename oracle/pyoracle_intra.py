"""ctypes bindings of the intra-prediction oracle (TEST INFRASTRUCTURE ONLY, see pyoracle.py)."""
import ctypes as C

import numpy as np

from . import pyoracle as po
from openvvc_b200.synth_intra import INTRA_CMD


def intra_picture(planes, cmds, log2_ctu=7, bitdepth=10, impl="port", simd=1):
    """Predicts every block in decode order, in place.  impl="ref" also fills the commands' availability words."""
    h, w = planes[0].shape
    assert cmds.dtype == INTRA_CMD and cmds.flags["C_CONTIGUOUS"]
    dst, ds = po._planes_arg(planes)
    if impl == "port":
        po.port().ovo_intra_picture(dst, ds, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds))
    else:
        po.ref().ovref_intra_picture(dst, ds, w, h, log2_ctu, bitdepth, simd, cmds.ctypes.data_as(C.c_void_p), len(cmds))
    return planes


from openvvc_b200.synth_tu import TU_CMD


def isp_info_array(words, tus):
    """OVOIspInfo[n]: availability words + the four transform blocks of every ISP coding unit"""
    dt = np.dtype([("common", np.uint64), ("part", np.uint64, 4), ("tu", TU_CMD, 4)])
    assert dt.itemsize == 40 + 4 * 72
    out = np.zeros(len(words), dtype=dt)
    out["common"], out["part"], out["tu"] = words["common"], words["part"], tus
    return out


def intra_picture_isp(planes, cmds, isp, arena, log2_ctu=7, bitdepth=10, impl="port", simd=1, dep_quant=1, mts_enabled=1):
    """like intra_picture; ISP commands are reconstructed (prediction + residual per partition).  impl="ref" also fills the
    availability words of cmds and isp in place."""
    h, w = planes[0].shape
    assert cmds.dtype == INTRA_CMD and cmds.flags["C_CONTIGUOUS"] and isp.flags["C_CONTIGUOUS"]
    arena = np.ascontiguousarray(arena, dtype=np.int16)
    dst, ds = po._planes_arg(planes)
    if impl == "port":
        po.port().ovo_intra_picture2(dst, ds, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds),
                                     isp.ctypes.data_as(C.c_void_p), arena.ctypes.data_as(C.c_void_p), dep_quant)
    else:
        po.ref().ovref_intra_picture2(dst, ds, w, h, log2_ctu, bitdepth, simd, cmds.ctypes.data_as(C.c_void_p), len(cmds),
                                      isp.ctypes.data_as(C.c_void_p), arena.ctypes.data_as(C.c_void_p), dep_quant, mts_enabled)
    return planes
