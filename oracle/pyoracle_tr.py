"""ctypes bindings of the residual-path oracle (TEST INFRASTRUCTURE ONLY, see pyoracle.py)."""
import ctypes as C

import numpy as np

from . import pyoracle as po

from openvvc_b200.synth_tu import TU_CMD as TuCmd   # one definition of the 72-byte command layout


def residual_picture(planes, cmds, arena, log2_ctu=7, bitdepth=10, impl="port", dep_quant=1, sh_ts_disabled=0, ict_type=0, simd=1):
    """Adds the residual of every TU to `planes` in place (planes hold the prediction)."""
    h, w = planes[0].shape
    cmds = np.ascontiguousarray(cmds, dtype=TuCmd)
    arena = np.ascontiguousarray(arena, dtype=np.int16)
    dst, ds = po._planes_arg(planes)
    if impl == "port":
        po.port().ovo_residual_picture(dst, ds, w, h, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds),
                                       arena.ctypes.data_as(C.c_void_p), dep_quant, sh_ts_disabled, ict_type)
    else:
        po.ref().ovref_residual_picture(dst, ds, w, h, log2_ctu, bitdepth, cmds.ctypes.data_as(C.c_void_p), len(cmds),
                                        arena.ctypes.data_as(C.c_void_p), dep_quant, sh_ts_disabled, ict_type, simd)
    return planes
