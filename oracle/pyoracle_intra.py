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
