"""Shared test helpers: seeded cases, md5, and adapters between the oracle's and the product's formats."""
import ctypes as C
import hashlib

import numpy as np

from openvvc_b200 import synth


def md5(*arrays):
    h = hashlib.md5()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def alf_case(po, w, h, seed, log2_ctu=7, kind="natural", wide=False):
    """planes + tables (oracle struct) + per-CTU dict for one seeded ALF picture."""
    planes = synth.planes(w, h, seed, kind)
    ap = synth.alf_params(w, h, seed + 1, log2_ctu)
    if wide:
        ap["luma_aps"][0]["luma_coeff"][:, 3] = 300     # outside int8: exercises the WIDE kernel variant
        ap["chroma_aps"]["chroma_coeff"][0, 2] = 200    # ... and the scalar fall-backs of the chroma kernel (5x5 core, CC-ALF)
        ap["cc_coeff"][0, 1, 3] = 130
        ap["cc_coeff"][1, 0, 5] = -140
    tab = po.alf_build_tables(ap["luma_aps"], ap["chroma_aps"], impl="port")
    tab.np("cc_coeff")[:] = ap["cc_coeff"]
    return planes, tab, ap["ctu"], ap


def sao_case(po, w, h, seed, log2_ctu=7, kind="natural"):
    planes = synth.planes(w, h, seed, kind)
    return planes, synth.sao_params(w, h, seed + 2, log2_ctu, po.SaoCtu)


def alf_ctu_rows(ctu):
    n = len(ctu["flag"])
    a = np.zeros((n, 4), dtype=np.uint8)
    a[:, 0] = ctu["flag"]; a[:, 1] = ctu["idx"]; a[:, 2] = ctu["cb_alt"]; a[:, 3] = ctu["cr_alt"]
    return a


def planes_arg(planes):
    u16p = C.POINTER(C.c_uint16)
    arr = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in planes])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in planes])
    return arr, strides


def emu_alf(emu, planes, tab, ctu, log2_ctu=7, bitdepth=10):
    h, w = planes[0].shape
    out = [np.zeros_like(p) for p in planes]
    src, ss = planes_arg(planes); dst, ds = planes_arg(out)
    rows = alf_ctu_rows(ctu)
    u8p = C.POINTER(C.c_uint8)
    cc0 = np.ascontiguousarray(ctu["cc_cb"], dtype=np.uint8); cc1 = np.ascontiguousarray(ctu["cc_cr"], dtype=np.uint8)
    cc = (u8p * 2)(cc0.ctypes.data_as(u8p), cc1.ctypes.data_as(u8p))
    wide = emu.emu_alf_picture(dst, ds, src, ss, w, h, log2_ctu, bitdepth, C.byref(tab), C.c_void_p(rows.ctypes.data), cc,
                               int(ctu["cc_cb_enabled"]), int(ctu["cc_cr_enabled"]))
    return out, wide


def emu_sao(emu, planes, sao_ctu, log2_ctu=7, bitdepth=10):
    h, w = planes[0].shape
    out = [np.zeros_like(p) for p in planes]
    src, ss = planes_arg(planes); dst, ds = planes_arg(out)
    emu.emu_sao_picture(dst, ds, src, ss, w, h, log2_ctu, bitdepth, C.c_void_p(sao_ctu.ctypes.data))
    return out


def gpu_tables(po_tab):
    """Oracle OVOAlfTables -> product OVGPUAlfTables (byte-identical layouts)."""
    from openvvc_b200 import api
    t = api.AlfTables()
    assert C.sizeof(t) == C.sizeof(po_tab)
    C.memmove(C.byref(t), C.byref(po_tab), C.sizeof(t))
    return t


def assert_planes_equal(a, b, what=""):
    for c in range(3):
        d = a[c] != b[c]
        if d.any():
            ys, xs = np.nonzero(d)
            raise AssertionError("%s: component %d differs at %d samples, first (y=%d,x=%d): %d vs %d" % (
                what, c, d.sum(), ys[0], xs[0], a[c][ys[0], xs[0]], b[c][ys[0], xs[0]]))
