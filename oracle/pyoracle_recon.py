"""ctypes binding of the CU-order reconstruction oracle (TEST INFRASTRUCTURE ONLY, see pyoracle.py)."""
import ctypes as C

import numpy as np

from . import pyoracle as po
from openvvc_b200.synth_intra import INTRA_CMD
from openvvc_b200.synth_tu import TU_CMD


LMCS_ANCHOR = np.dtype([("x0", np.int16), ("y0", np.int16), ("abv_mask", np.uint16), ("lft_mask", np.uint16)])


def lmcs_chroma_scale(luma, x0, y0, abv_mask, lft_mask, luts, data, bitdepth=10, impl="port"):
    """lmcs_chroma_scale for the 64x64 region at (x0, y0) of a luma plane (rcn_lmcs_compute_chroma_scale)"""
    fn = po.port().ovo_lmcs_chroma_scale if impl == "port" else po.ref().ovref_lmcs_chroma_scale
    fn.restype = C.c_int
    luma = np.ascontiguousarray(luma)
    base = luma.ctypes.data + 2 * (y0 * luma.shape[1] + x0)
    return fn(C.c_void_p(base), luma.shape[1], int(abv_mask), int(lft_mask), C.byref(luts), C.byref(data), bitdepth)


def recon_picture(planes, cmds, tus, arena, log2_ctu=7, bitdepth=10, dep_quant=1, sh_ts_disabled=0, ict_type=0,
                  anchors=None, luts=None, lmcs_data=None, isp=None):
    """prediction then residual, coding unit by coding unit in decode order (oracle/port_recon.c), in place"""
    h, w = planes[0].shape
    assert cmds.dtype == INTRA_CMD and tus.dtype == TU_CMD and len(cmds) == len(tus)
    cmds = np.ascontiguousarray(cmds); tus = np.ascontiguousarray(tus)
    arena = np.ascontiguousarray(arena, dtype=np.int16)
    dst, ds = po._planes_arg(planes)
    po.port().ovo_recon_picture2(dst, ds, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), tus.ctypes.data_as(C.c_void_p),
                                len(cmds), arena.ctypes.data_as(C.c_void_p), dep_quant, sh_ts_disabled, ict_type,
                                anchors.ctypes.data_as(C.c_void_p) if anchors is not None else None, len(anchors) if anchors is not None else 0,
                                C.byref(luts) if luts is not None else None, C.byref(lmcs_data) if lmcs_data is not None else None,
                                isp.ctypes.data_as(C.c_void_p) if isp is not None else None)
    return planes
