"""ctypes binding of the CU-order reconstruction oracle (TEST INFRASTRUCTURE ONLY, see pyoracle.py)."""
import ctypes as C

import numpy as np

from . import pyoracle as po
from openvvc_b200.synth_intra import INTRA_CMD
from openvvc_b200.synth_tu import TU_CMD


def recon_picture(planes, cmds, tus, arena, log2_ctu=7, bitdepth=10, dep_quant=1, sh_ts_disabled=0, ict_type=0):
    """prediction then residual, coding unit by coding unit in decode order (oracle/port_recon.c), in place"""
    h, w = planes[0].shape
    assert cmds.dtype == INTRA_CMD and tus.dtype == TU_CMD and len(cmds) == len(tus)
    cmds = np.ascontiguousarray(cmds); tus = np.ascontiguousarray(tus)
    arena = np.ascontiguousarray(arena, dtype=np.int16)
    dst, ds = po._planes_arg(planes)
    po.port().ovo_recon_picture(dst, ds, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), tus.ctypes.data_as(C.c_void_p),
                                len(cmds), arena.ctypes.data_as(C.c_void_p), dep_quant, sh_ts_disabled, ict_type)
    return planes
