"""ctypes bindings for the oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers live behind this module:

* ``port``  -- oracle/liboracle.so, our plain-C restatement of the reference algorithms (oracle/port_*.c);
* ``ref``   -- oracle/_ref/libovref_harness.so, a driver around the UNMODIFIED OpenVVC reference compiled
  from /root/reference by oracle/Makefile (prebuilt files travel to the GPU box; sources never do).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (openvvc_b200) must never import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"

u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)


class AlfTables(C.Structure):
    """OVOAlfTables (oracle.h) == OVGPUAlfTables (include/ovvc_b200.h) == RCNALF + cc coeffs."""
    _fields_ = [("luma_coeff", C.c_int16 * (24 * 1300)),
                ("luma_clip", C.c_int16 * (24 * 1300)),
                ("chroma_coeff", C.c_int16 * (8 * 7)),
                ("chroma_clip", C.c_int16 * (8 * 7)),
                ("cc_coeff", C.c_int16 * (2 * 4 * 8))]

    def np(self, name):
        shape = {"luma_coeff": (24, 4, 25, 13), "luma_clip": (24, 4, 25, 13),
                 "chroma_coeff": (8, 7), "chroma_clip": (8, 7), "cc_coeff": (2, 4, 8)}[name]
        return np.ctypeslib.as_array(getattr(self, name)).reshape(shape)


class AlfCtuParams(C.Structure):
    _fields_ = [("ctb_alf_flag", u8p), ("ctb_alf_idx", u8p), ("cb_alternative", u8p),
                ("cr_alternative", u8p), ("cc_idx_cb", u8p), ("cc_idx_cr", u8p),
                ("cc_cb_enabled", C.c_uint8), ("cc_cr_enabled", C.c_uint8)]


class AlfAps(C.Structure):
    _fields_ = [("luma_clip_flag", C.c_uint8), ("num_luma_filters", C.c_uint8),
                ("luma_coeff_delta_idx", C.c_uint8 * 25),
                ("luma_coeff", C.c_int16 * (25 * 13)), ("luma_clip_idx", C.c_int16 * (25 * 13)),
                ("chroma_clip_flag", C.c_uint8), ("num_chroma_alts", C.c_uint8),
                ("chroma_coeff", C.c_int16 * (8 * 7)), ("chroma_clip_idx", C.c_int16 * (8 * 7))]


SaoCtu = np.dtype([("offset_val", np.int16, (3, 5)), ("band_position", np.uint8, 3),
                   ("eo_class", np.uint8, 3), ("type_idx", np.uint8, 3), ("pad", np.uint8)])
assert SaoCtu.itemsize == 40


def build(ref=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref and os.path.isdir(os.path.join(REF_SRC, "libovvc")):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _port = C.CDLL(path)
    return _port


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libovref_harness.so"))


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(HERE, "_ref", "libovref_harness.so"))
    return _ref


def _planes_arg(planes):
    arr = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in planes])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in planes])
    return arr, strides


def _ctu_params(ctu):
    """ctu: dict of uint8 arrays (flag, idx, cb_alt, cr_alt, cc_cb, cc_cr) + cc enables."""
    keep = [np.ascontiguousarray(ctu[k], dtype=np.uint8) for k in
            ("flag", "idx", "cb_alt", "cr_alt", "cc_cb", "cc_cr")]
    p = AlfCtuParams(*[a.ctypes.data_as(u8p) for a in keep],
                     int(ctu.get("cc_cb_enabled", 1)), int(ctu.get("cc_cr_enabled", 1)))
    return p, keep


def make_aps(d):
    a = AlfAps()
    a.luma_clip_flag = int(d["luma_clip_flag"])
    a.num_luma_filters = int(d["num_luma_filters"])
    a.luma_coeff_delta_idx[:] = [int(v) for v in d["luma_coeff_delta_idx"]]
    a.luma_coeff[:] = np.asarray(d["luma_coeff"], dtype=np.int16).reshape(-1).tolist()
    a.luma_clip_idx[:] = np.asarray(d["luma_clip_idx"], dtype=np.int16).reshape(-1).tolist()
    a.chroma_clip_flag = int(d["chroma_clip_flag"])
    a.num_chroma_alts = int(d["num_chroma_alts"])
    a.chroma_coeff[:] = np.asarray(d["chroma_coeff"], dtype=np.int16).reshape(-1).tolist()
    a.chroma_clip_idx[:] = np.asarray(d["chroma_clip_idx"], dtype=np.int16).reshape(-1).tolist()
    return a


def alf_build_tables(luma_aps, chroma_aps, bitdepth=10, impl="port"):
    t = AlfTables()
    arr = (AlfAps * max(1, len(luma_aps)))(*[make_aps(d) for d in luma_aps])
    ca = make_aps(chroma_aps) if chroma_aps is not None else None
    fn = port().ovo_alf_build_tables if impl == "port" else ref().ovref_alf_build_tables
    fn(C.byref(t), arr, len(luma_aps), C.byref(ca) if ca is not None else None, bitdepth)
    return t


def alf_picture(planes, tables, ctu, log2_ctu=7, bitdepth=10, impl="port", simd=1, tiles=None):
    """planes: 3 contiguous uint16 arrays (post-SAO). Returns the 3 filtered planes.  tiles: TileGrid or None (one tile)."""
    h, w = planes[0].shape
    p, keep = _ctu_params(ctu)
    tg = C.byref(tiles) if tiles is not None else None
    if impl == "port":
        out = [np.empty_like(pl) for pl in planes]
        src, ss = _planes_arg(planes)
        dst, ds = _planes_arg(out)
        f = port().ovo_alf_picture_tiles
        f.restype = None
        f(dst, ds, src, ss, w, h, log2_ctu, bitdepth, C.byref(tables), C.byref(p), tg)
        return out
    out = [pl.copy() for pl in planes]
    dst, ds = _planes_arg(out)
    if tiles is not None:
        f = ref().ovref_alf_picture_tiles
        f.restype = None
        f(dst, ds, w, h, log2_ctu, bitdepth, simd, C.byref(tables), C.byref(p), tg)
    else:
        ref().ovref_alf_picture(dst, ds, w, h, log2_ctu, bitdepth, simd, C.byref(tables), C.byref(p))
    return out


def alf_classify(luma, log2_ctu=7, bitdepth=10):
    h, w = luma.shape
    cls = np.zeros((h // 4, w // 4), dtype=np.uint8)
    port().ovo_alf_classify_picture(cls.ctypes.data_as(u8p), luma.ctypes.data_as(u16p),
                                    luma.strides[0] // 2, w, h, log2_ctu, bitdepth)
    return cls


class TileGrid(C.Structure):
    """OVOTileGrid / OVGPUTileGrid: tile boundaries in CTUs (col_bd[0] = 0 ... col_bd[nb_cols] = picture width in CTUs)"""
    _fields_ = [("nb_cols", C.c_int32), ("nb_rows", C.c_int32), ("col_bd", C.c_uint16 * 22), ("row_bd", C.c_uint16 * 22)]


def make_tile_grid(col_bd, row_bd):
    tg = TileGrid()
    tg.nb_cols, tg.nb_rows = len(col_bd) - 1, len(row_bd) - 1
    tg.col_bd[:len(col_bd)] = [int(v) for v in col_bd]
    tg.row_bd[:len(row_bd)] = [int(v) for v in row_bd]
    return tg


def sao_picture(planes, sao_ctu, log2_ctu=7, bitdepth=10, impl="port", simd=1, tiles=None):
    """tiles: TileGrid or None (one tile).  The reference is driven one rect entry per tile."""
    h, w = planes[0].shape
    sao_ctu = np.ascontiguousarray(sao_ctu, dtype=SaoCtu)
    if impl == "port":
        out = [np.empty_like(pl) for pl in planes]
        src, ss = _planes_arg(planes)
        dst, ds = _planes_arg(out)
        f = port().ovo_sao_picture_tiles
        f.restype = None
        f(dst, ds, src, ss, w, h, log2_ctu, bitdepth, sao_ctu.ctypes.data_as(C.c_void_p), C.byref(tiles) if tiles is not None else None)
        return out
    out = [pl.copy() for pl in planes]
    dst, ds = _planes_arg(out)
    if tiles is not None:
        f = ref().ovref_sao_picture_tiles
        f.restype = None
        f(dst, ds, w, h, log2_ctu, bitdepth, simd, sao_ctu.ctypes.data_as(C.c_void_p), C.byref(tiles))
    else:
        ref().ovref_sao_picture(dst, ds, w, h, log2_ctu, bitdepth, simd, sao_ctu.ctypes.data_as(C.c_void_p))
    return out


# ---- motion compensation / LMCS -------------------------------------------------------------------------
McCmd = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8),
                  ("inter_dir", np.uint8), ("flags", np.uint8), ("ref_idx0", np.int8), ("ref_idx1", np.int8),
                  ("bcw_idx_plus1", np.uint8), ("tool", np.uint8),
                  ("mv0_x", np.int32), ("mv0_y", np.int32), ("mv1_x", np.int32), ("mv1_y", np.int32),
                  ("aux", np.uint32)])
assert McCmd.itemsize == 32


class McRefs(C.Structure):
    _fields_ = [("plane", u16p * 3 * 16 * 2), ("stride", C.c_int * 3 * 16 * 2), ("poc", C.c_int32 * 16 * 2)]


class LmcsData(C.Structure):
    _fields_ = [("lmcs_min_bin_idx", C.c_uint8), ("lmcs_delta_max_bin_idx", C.c_uint8),
                ("lmcs_delta_cw_prec_minus1", C.c_uint8), ("lmcs_delta_abs_cw", C.c_uint16 * 16),
                ("lmcs_delta_sign_cw_flag", C.c_uint8 * 16), ("lmcs_delta_abs_crs", C.c_uint8),
                ("lmcs_delta_sign_crs_flag", C.c_uint8)]


class LmcsLuts(C.Structure):
    _fields_ = [("fwd_lut", C.c_uint16 * 1024), ("bwd_lut", C.c_uint16 * 1024), ("wnd_bnd", C.c_uint16 * 17)]

    def np(self, name):
        return np.ctypeslib.as_array(getattr(self, name))


def make_lmcs_data(d):
    x = LmcsData()
    x.lmcs_min_bin_idx = int(d["min_bin_idx"]); x.lmcs_delta_max_bin_idx = int(d["delta_max_bin_idx"])
    x.lmcs_delta_cw_prec_minus1 = 0
    x.lmcs_delta_abs_cw[:] = [int(v) for v in d["delta_abs_cw"]]
    x.lmcs_delta_sign_cw_flag[:] = [int(v) for v in d["delta_sign_cw"]]
    x.lmcs_delta_abs_crs = int(d["delta_abs_crs"]); x.lmcs_delta_sign_crs_flag = int(d["delta_sign_crs"])
    return x


def lmcs_build_luts(d, bitdepth=10, impl="port"):
    luts = LmcsLuts()
    data = make_lmcs_data(d)
    fn = port().ovo_lmcs_build_luts if impl == "port" else ref().ovref_lmcs_build_luts
    fn(C.byref(luts), C.byref(data), bitdepth)
    return luts


def lmcs_reshape_backward(luma, luts, log2_ctu=7, bitdepth=10, impl="port"):
    """in-place backward reshape of a luma plane: impl="ref" = the compiled reference's own lmcs_reshape_backward slot, CTU by CTU
    (oracle/ref_harness.c); "port" = the C restatement"""
    h, w = luma.shape
    if impl == "ref":
        f = ref().ovref_lmcs_reshape_backward_plane
        f.restype = None
        f(luma.ctypes.data_as(u16p), luma.strides[0] // 2, w, h, C.byref(luts), log2_ctu, bitdepth)
    else:
        port().ovo_lmcs_reshape_plane(luma.ctypes.data_as(u16p), luma.strides[0] // 2, w, h, luts.np("bwd_lut").ctypes.data_as(u16p), bitdepth)


def make_mc_refs(ref_lists, pocs):
    """ref_lists: [list0, list1], each a list of pictures (3 uint16 planes); pocs likewise."""
    r = McRefs()
    keep = []
    for l in range(2):
        for i, pic in enumerate(ref_lists[l]):
            for c in range(3):
                r.plane[l][i][c] = pic[c].ctypes.data_as(u16p)
                r.stride[l][i][c] = pic[c].strides[0] // 2
            r.poc[l][i] = int(pocs[l][i])
            keep.append(pic)
    return r, keep


class McScaling(C.Structure):
    """OVOMcScaling (oracle.h): scale[list][ref_idx][hor, ver] in 2^-14 units, the references' own sizes, the SPS's chroma siting"""
    _fields_ = [("scale", C.c_uint16 * 2 * 16 * 2), ("ref_w", C.c_int32 * 16 * 2), ("ref_h", C.c_int32 * 16 * 2),
                ("chroma_hor_col", C.c_uint8), ("chroma_ver_col", C.c_uint8)]


def make_mc_scaling(w, h, ref_lists, hor_col=1, ver_col=0):
    """scale factors as ctudec_compute_refs_scaling derives them (ctudec.c:42-84), no scaling-window offsets"""
    sc = McScaling()
    for l in range(2):
        for i in range(16):
            sc.scale[l][i][0] = sc.scale[l][i][1] = 1 << 14
            sc.ref_w[l][i] = w; sc.ref_h[l][i] = h
        for i, pic in enumerate(ref_lists[l]):
            rh, rw = pic[0].shape
            sc.ref_w[l][i] = rw; sc.ref_h[l][i] = rh
            sc.scale[l][i][0] = ((rw << 14) + (w >> 1)) // w
            sc.scale[l][i][1] = ((rh << 14) + (h >> 1)) // h
    sc.chroma_hor_col = hor_col; sc.chroma_ver_col = ver_col
    return sc


def mc_picture_rpr(dst_planes, ref_lists, pocs, cmds, sc, log2_ctu=7, bitdepth=10, impl="port", lmcs=None, prof=None, simd=0):
    """mc_picture with references of their own sizes (reference picture resampling); sc = make_mc_scaling(...)"""
    h, w = dst_planes[0].shape
    cmds = np.ascontiguousarray(cmds, dtype=McCmd)
    refs, keep = make_mc_refs(ref_lists, pocs)
    dst, ds = _planes_arg(dst_planes)
    prof = np.ascontiguousarray(prof, dtype=np.int16) if prof is not None else None
    pp = prof.ctypes.data_as(C.c_void_p) if prof is not None else None
    if impl == "port":
        lut = lmcs.np("fwd_lut").ctypes.data_as(u16p) if lmcs is not None else None
        f = port().ovo_mc_picture_rpr
        f.restype = None
        f(dst, ds, w, h, bitdepth, C.byref(refs), C.byref(sc), cmds.ctypes.data_as(C.c_void_p), len(cmds), lut, pp)
    else:
        f = ref().ovref_mc_picture_rpr
        f.restype = None
        f(dst, ds, w, h, log2_ctu, bitdepth, simd, C.byref(refs), C.byref(sc), cmds.ctypes.data_as(C.c_void_p), len(cmds), pp,
          C.byref(lmcs) if lmcs is not None else None)
    return dst_planes


DmvrMv = np.dtype([("mv0_x", np.int32), ("mv0_y", np.int32), ("mv1_x", np.int32), ("mv1_y", np.int32)])


def mc_picture(dst_planes, ref_lists, pocs, cmds, log2_ctu=7, bitdepth=10, impl="port", lmcs=None, prof=None, simd=1, dmvr_out=None):
    """Writes the predictions of all commands into dst_planes (in place) and returns them.
    prof: int16 array [n][4][16] (struct PROFInfo: dmv_scale_h_0, v_0, h_1, v_1) indexed by tool-2 commands."""
    h, w = dst_planes[0].shape
    cmds = np.ascontiguousarray(cmds, dtype=McCmd)
    refs, keep = make_mc_refs(ref_lists, pocs)
    dst, ds = _planes_arg(dst_planes)
    prof = np.ascontiguousarray(prof, dtype=np.int16) if prof is not None else None
    pp = prof.ctypes.data_as(C.c_void_p) if prof is not None else None
    # dmvr_out: DmvrMv array with one entry per tool-3 command (refined MVs, in command order)
    po_ = dmvr_out.ctypes.data_as(C.c_void_p) if dmvr_out is not None else None
    if impl == "port":
        lut = lmcs.np("fwd_lut").ctypes.data_as(u16p) if lmcs is not None else None
        port().ovo_mc_picture3(dst, ds, w, h, bitdepth, C.byref(refs), cmds.ctypes.data_as(C.c_void_p), len(cmds), lut, pp, po_)
    else:
        ref().ovref_mc_picture(dst, ds, w, h, log2_ctu, bitdepth, simd, C.byref(refs), cmds.ctypes.data_as(C.c_void_p),
                               len(cmds), pp, C.byref(lmcs) if lmcs is not None else None, po_)
    return dst_planes


# ---- output stage: up-scaling ---------------------------------------------------------------------------------
class ScalingWindow(C.Structure):
    """OVOScalingWindow: struct ScalingInfo (ovdpb.h:85-93), offsets in chroma sample units"""
    _fields_ = [("left", C.c_uint16), ("right", C.c_uint16), ("top", C.c_uint16), ("bottom", C.c_uint16),
                ("chroma_hor_col", C.c_uint8), ("chroma_ver_col", C.c_uint8)]


def pp_upscale(planes, dst_w, dst_h, window=(0, 0, 0, 0), hor_col=1, ver_col=0, impl="port", bitdepth=10):
    """pp_sample_rate_conv on the three planes of a decoded picture -> planes of dst_w x dst_h (post_proc.c:117-128)"""
    win = ScalingWindow(*[int(v) for v in window], int(hor_col), int(ver_col))
    out = []
    for c, p in enumerate(planes):
        p = np.ascontiguousarray(p)
        dw, dh = (dst_w >> 1, dst_h >> 1) if c else (dst_w, dst_h)
        d = np.zeros((dh, dw), np.uint16)
        sh, sw = p.shape
        if impl == "port":
            rc = port().ovo_pp_upscale_plane(d.ctypes.data_as(u16p), dw, dw, dh, p.ctypes.data_as(u16p), p.strides[0] // 2, sw, sh, C.byref(win), int(c == 0), bitdepth)
            assert rc == 0, "down-sampling is not restated"
        else:
            f = ref().ovref_pp_upscale_plane
            f.restype = None
            f(d.ctypes.data_as(u16p), dw, dw, dh, p.ctypes.data_as(u16p), p.strides[0] // 2, sw, sh, C.byref(win), int(c == 0))
        out.append(d)
    return out


# ---- output stage: film grain synthesis --------------------------------------------------------------------------
class FilmGrain(C.Structure):
    """OVOFilmGrain: the fields of the film grain characteristics SEI the synthesis reads (struct OVSEIFGrain)"""
    _fields_ = [("log2_scale_factor", C.c_uint8), ("comp_model_present_flag", C.c_uint8 * 3), ("num_model_values_minus1", C.c_uint8 * 3),
                ("intensity_interval_lower_bound", C.c_uint8 * 8 * 3), ("intensity_interval_upper_bound", C.c_uint8 * 8 * 3),
                ("comp_model_value", C.c_int16 * 3 * 8 * 3)]


def make_film_grain(seed, present=(1, 1, 1)):
    """a random but well-formed SEI: log2 scale 2..7, up to 8 disjoint intensity intervals per component, 1..3 model values
    (standard deviation 0..255, cut-off frequencies 2..14)"""
    rng = np.random.default_rng(seed)
    fg = FilmGrain()
    fg.log2_scale_factor = int(rng.integers(2, 8))
    for c in range(3):
        fg.comp_model_present_flag[c] = int(present[c])
        fg.num_model_values_minus1[c] = int(rng.integers(0, 3))
        n = int(rng.integers(1, 9))
        cuts = np.sort(rng.choice(np.arange(1, 255), size=2 * n, replace=False))
        for i in range(n):
            fg.intensity_interval_lower_bound[c][i] = int(cuts[2 * i]); fg.intensity_interval_upper_bound[c][i] = int(cuts[2 * i + 1])
            fg.comp_model_value[c][i][0] = int(rng.integers(0, 256))
            fg.comp_model_value[c][i][1] = int(rng.integers(2, 15)); fg.comp_model_value[c][i][2] = int(rng.integers(2, 15))
    return fg


_fg_db = {}


def fg_database(impl="port"):
    if impl not in _fg_db:
        db = np.zeros(13 * 13 * 64 * 64, np.int8)
        (port().ovo_fg_database if impl == "port" else ref().ovref_fg_database)(db.ctypes.data_as(C.c_void_p), 1)
        _fg_db[impl] = db
    return _fg_db[impl]


def fg_apply(planes, fg, poc, impl="port"):
    """fg_grain_apply_pic on the three planes of a decoded picture"""
    src = [np.ascontiguousarray(p) for p in planes]
    h, w = src[0].shape
    out = [np.zeros_like(p) for p in src]
    d3 = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in out]); s3 = (u16p * 3)(*[p.ctypes.data_as(u16p) for p in src])
    if impl == "port":
        rc = port().ovo_fg_apply(d3, s3, w, h, C.byref(fg), int(poc), fg_database("port").ctypes.data_as(C.c_void_p))
        assert rc == 0, "picture size the reference's block loops do not handle"
    else:
        f = ref().ovref_fg_apply
        f.restype = None
        f(d3, s3, w, h, C.byref(fg), int(poc))
    return out
