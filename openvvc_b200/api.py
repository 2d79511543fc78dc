"""Thin ctypes binding of the C ABI in include/ovvc_b200.h (libovvc_b200.so).

This is what tests and bench.py call; it adds nothing to the data path.  There is no CPU fallback:
if the library is missing it is built (nvcc), and if no sm_100 GPU is present Ctx() raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libovvc_b200.so")

vp = C.c_void_p
szp = C.POINTER(C.c_size_t)


class AlfTables(C.Structure):
    """OVGPUAlfTables: RCNALF (rcn_alf.h:60-66) + CC-ALF coefficients."""
    _fields_ = [("luma_coeff", C.c_int16 * (24 * 1300)), ("luma_clip", C.c_int16 * (24 * 1300)),
                ("chroma_coeff", C.c_int16 * (8 * 7)), ("chroma_clip", C.c_int16 * (8 * 7)),
                ("cc_coeff", C.c_int16 * (2 * 4 * 8))]

    def np(self, name):
        shape = {"luma_coeff": (24, 4, 25, 13), "luma_clip": (24, 4, 25, 13),
                 "chroma_coeff": (8, 7), "chroma_clip": (8, 7), "cc_coeff": (2, 4, 8)}[name]
        return np.ctypeslib.as_array(getattr(self, name)).reshape(shape)


class AlfAps(C.Structure):
    """OVGPUAlfAps: the OVALFData fields rcn_alf_reconstruct_coeff_APS reads."""
    _fields_ = [("luma_clip_flag", C.c_uint8), ("num_luma_filters", C.c_uint8),
                ("luma_coeff_delta_idx", C.c_uint8 * 25),
                ("luma_coeff", C.c_int16 * (25 * 13)), ("luma_clip_idx", C.c_int16 * (25 * 13)),
                ("chroma_clip_flag", C.c_uint8), ("num_chroma_alts", C.c_uint8),
                ("chroma_coeff", C.c_int16 * (8 * 7)), ("chroma_clip_idx", C.c_int16 * (8 * 7))]

    @classmethod
    def from_dict(cls, d):
        a = cls()
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


ALF_CTU = np.dtype([("ctb_alf_flag", np.uint8), ("ctb_alf_idx", np.uint8),
                    ("cb_alternative", np.uint8), ("cr_alternative", np.uint8)])
SAO_CTU = np.dtype([("offset_val", np.int16, (3, 5)), ("band_position", np.uint8, 3),
                    ("eo_class", np.uint8, 3), ("type_idx", np.uint8, 3), ("pad", np.uint8)])

STAGES = {"alf_luma": 0, "alf_chroma": 1, "sao": 2, "dbf_v": 3, "dbf_h": 4, "mc": 5, "residual": 6,
          "intra": 7, "lmcs": 8}

_lib = None


def lib():
    """Load (building if needed) libovvc_b200.so. Raises if it cannot be built/loaded."""
    global _lib
    if _lib is None:
        _build.build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ovgpu_last_error.restype = C.c_char_p
        _lib.ovgpu_launch_count.restype = C.c_uint64
    return _lib


class OVGPUError(RuntimeError):
    pass


def _check(ctx, ret, what):
    if ret != 0:
        msg = lib().ovgpu_last_error(ctx).decode() if ctx else ""
        raise OVGPUError("%s failed: %d %s" % (what, ret, msg))


def _host_planes(planes):
    data = (vp * 3)(*[p.ctypes.data for p in planes])
    ls = (C.c_size_t * 3)(*[p.strides[0] for p in planes])
    return data, ls


def alf_tables_from_aps(luma_aps, chroma_aps, bitdepth=10):
    """ovgpu_alf_reconstruct_coeff_aps: host-side table build (rcn_alf.c:264)."""
    t = AlfTables()
    arr = (AlfAps * max(1, len(luma_aps)))(*[AlfAps.from_dict(d) for d in luma_aps])
    ca = AlfAps.from_dict(chroma_aps) if chroma_aps is not None else None
    ret = lib().ovgpu_alf_reconstruct_coeff_aps(C.byref(t), arr, len(luma_aps),
                                                C.byref(ca) if ca is not None else None, bitdepth)
    _check(None, ret, "ovgpu_alf_reconstruct_coeff_aps")
    return t


def alf_ctu_array(ctu):
    n = len(ctu["flag"])
    a = np.zeros(n, dtype=ALF_CTU)
    a["ctb_alf_flag"] = ctu["flag"]; a["ctb_alf_idx"] = ctu["idx"]
    a["cb_alternative"] = ctu["cb_alt"]; a["cr_alternative"] = ctu["cr_alt"]
    return a


class Ctx:
    def __init__(self, device=0, stream=None):
        self.h = vp()
        ret = lib().ovgpu_ctx_create(C.byref(self.h), device, vp(stream) if stream else None)
        if ret != 0:
            raise OVGPUError("ovgpu_ctx_create failed: %d (no usable sm_100 GPU?)" % ret)

    def close(self):
        if self.h:
            lib().ovgpu_ctx_destroy(self.h)
            self.h = vp()

    def sync(self):
        _check(self.h, lib().ovgpu_ctx_sync(self.h), "sync")

    def launch_count(self):
        return int(lib().ovgpu_launch_count(self.h))

    def profile(self, on):
        lib().ovgpu_profile_enable(self.h, int(on))

    def profile_reset(self):
        lib().ovgpu_profile_reset(self.h)

    def profile_get(self, stage):
        ms = C.c_double(); n = C.c_uint64()
        _check(self.h, lib().ovgpu_profile_get(self.h, STAGES[stage], C.byref(ms), C.byref(n)), "profile_get")
        return ms.value, n.value

    def flush_l2(self):
        _check(self.h, lib().ovgpu_flush_l2(self.h), "flush_l2")

    # ---- pictures
    def picture(self, w, h, bitdepth=10):
        return Picture(self, w, h, bitdepth)

    def wrap_picture(self, w, h, dev_ptrs, pitches, bitdepth=10):
        return Picture(self, w, h, bitdepth, dev_ptrs, pitches)

    # ---- host-buffer (end-to-end) calls
    def alf_frame_host(self, planes, tables, ctu, cc_idx, cc_en=(1, 1), log2_ctu=7, bitdepth=10, out=None):
        h, w = planes[0].shape
        out = out or [np.empty_like(p) for p in planes]
        sd, sl = _host_planes(planes); dd, dl = _host_planes(out)
        ctu = np.ascontiguousarray(ctu, dtype=ALF_CTU)
        keep = [np.ascontiguousarray(c, dtype=np.uint8) for c in cc_idx]
        cc = (vp * 2)(keep[0].ctypes.data, keep[1].ctypes.data)
        ret = lib().ovgpu_alf_frame_host(self.h, dd, dl, sd, sl, w, h, log2_ctu, bitdepth, C.byref(tables),
                                         vp(ctu.ctypes.data), cc, int(cc_en[0]), int(cc_en[1]))
        _check(self.h, ret, "ovgpu_alf_frame_host")
        return out

    def sao_frame_host(self, planes, sao_ctu, log2_ctu=7, bitdepth=10, out=None):
        h, w = planes[0].shape
        out = out or [np.empty_like(p) for p in planes]
        sd, sl = _host_planes(planes); dd, dl = _host_planes(out)
        sao_ctu = np.ascontiguousarray(sao_ctu, dtype=SAO_CTU)
        ret = lib().ovgpu_sao_frame_host(self.h, dd, dl, sd, sl, w, h, log2_ctu, bitdepth, vp(sao_ctu.ctypes.data))
        _check(self.h, ret, "ovgpu_sao_frame_host")
        return out


class Picture:
    def __init__(self, ctx, w, h, bitdepth=10, dev_ptrs=None, pitches=None):
        self.ctx, self.w, self.h, self.bitdepth = ctx, w, h, bitdepth
        self.p = vp()
        if dev_ptrs is None:
            ret = lib().ovgpu_picture_create(ctx.h, w, h, bitdepth, C.byref(self.p))
        else:
            ptrs = (vp * 3)(*dev_ptrs)
            pit = (C.c_size_t * 3)(*pitches)
            ret = lib().ovgpu_picture_wrap(ctx.h, w, h, bitdepth, ptrs, pit, C.byref(self.p))
        _check(ctx.h, ret, "picture create/wrap")

    def upload(self, planes):
        d, l = _host_planes(planes)
        _check(self.ctx.h, lib().ovgpu_picture_upload(self.ctx.h, self.p, d, l), "picture_upload")

    def download(self, out=None):
        shapes = [(self.h, self.w), (self.h // 2, self.w // 2), (self.h // 2, self.w // 2)]
        out = out or [np.empty(s, dtype=np.uint16) for s in shapes]
        d, l = _host_planes(out)
        _check(self.ctx.h, lib().ovgpu_picture_download(self.ctx.h, self.p, d, l), "picture_download")
        self.ctx.sync()
        return out

    def planes(self):
        ptrs = (vp * 3)(); pit = (C.c_size_t * 3)()
        lib().ovgpu_picture_planes(self.p, ptrs, pit)
        return [int(x or 0) for x in ptrs], [int(x) for x in pit]

    def destroy(self):
        if self.p:
            lib().ovgpu_picture_destroy(self.p)
            self.p = vp()


class AlfParams:
    def __init__(self, ctx, w, h, log2_ctu=7):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_alf_params_create(ctx.h, w, h, log2_ctu, C.byref(self.p)), "alf_params_create")
        self._keep = None

    def upload(self, tables, ctu, cc_idx, cc_en=(1, 1), ctx=None):
        ctu = np.ascontiguousarray(ctu, dtype=ALF_CTU)
        keep = [np.ascontiguousarray(c, dtype=np.uint8) for c in cc_idx]
        cc = (vp * 2)(keep[0].ctypes.data, keep[1].ctypes.data)
        self._keep = (tables, ctu, keep)
        _check((ctx or self.ctx).h, lib().ovgpu_alf_params_upload((ctx or self.ctx).h, self.p, C.byref(tables), vp(ctu.ctypes.data), cc,
                                                         int(cc_en[0]), int(cc_en[1])), "alf_params_upload")

    def set_tiles(self, col_bd=None, row_bd=None, ctx=None):
        """ovgpu_alf_params_set_tiles: tile boundaries in CTUs (None: one tile)"""
        tg = None
        if col_bd is not None:
            tg = TileGrid()
            tg.nb_cols, tg.nb_rows = len(col_bd) - 1, len(row_bd) - 1
            tg.col_bd[:len(col_bd)] = [int(v) for v in col_bd]
            tg.row_bd[:len(row_bd)] = [int(v) for v in row_bd]
        _check((ctx or self.ctx).h, lib().ovgpu_alf_params_set_tiles((ctx or self.ctx).h, self.p, C.byref(tg) if tg is not None else None), "alf_params_set_tiles")

    def run(self, dst, src, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_alf_picture((ctx or self.ctx).h, dst.p, src.p, self.p), "alf_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_alf_params_destroy(self.p)
            self.p = vp()


MC_CMD = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8),
                   ("inter_dir", np.uint8), ("flags", np.uint8), ("ref_idx0", np.int8), ("ref_idx1", np.int8),
                   ("bcw_idx_plus1", np.uint8), ("tool", np.uint8),
                   ("mv0_x", np.int32), ("mv0_y", np.int32), ("mv1_x", np.int32), ("mv1_y", np.int32),
                   ("aux", np.uint32)])


DMVR_MV = np.dtype([("mv0_x", np.int32), ("mv0_y", np.int32), ("mv1_x", np.int32), ("mv1_y", np.int32)])


class McRefs(C.Structure):
    """OVGPUMcRefs: inter_ctx->rpl0/rpl1 as device pictures + POCs."""
    _fields_ = [("rpl", vp * 16 * 2), ("poc", C.c_int32 * 16 * 2)]


class McScaling(C.Structure):
    """OVGPUMcScaling: scale[list][ref_idx][hor, ver] (inter_ctx->scale_fact_rpl0 / 1, 1 << 14 = same size) + chroma siting flags"""
    _fields_ = [("scale", C.c_uint16 * 2 * 16 * 2), ("chroma_hor_col", C.c_uint8), ("chroma_ver_col", C.c_uint8), ("pad", C.c_uint8 * 2)]


class LmcsData(C.Structure):
    _fields_ = [("lmcs_min_bin_idx", C.c_uint8), ("lmcs_delta_max_bin_idx", C.c_uint8),
                ("lmcs_delta_cw_prec_minus1", C.c_uint8), ("lmcs_delta_abs_cw", C.c_uint16 * 16),
                ("lmcs_delta_sign_cw_flag", C.c_uint8 * 16), ("lmcs_delta_abs_crs", C.c_uint8),
                ("lmcs_delta_sign_crs_flag", C.c_uint8)]

    @classmethod
    def from_dict(cls, d):
        x = cls()
        x.lmcs_min_bin_idx = int(d["min_bin_idx"]); x.lmcs_delta_max_bin_idx = int(d["delta_max_bin_idx"])
        x.lmcs_delta_abs_cw[:] = [int(v) for v in d["delta_abs_cw"]]
        x.lmcs_delta_sign_cw_flag[:] = [int(v) for v in d["delta_sign_cw"]]
        x.lmcs_delta_abs_crs = int(d["delta_abs_crs"]); x.lmcs_delta_sign_crs_flag = int(d["delta_sign_crs"])
        return x


class LmcsLuts(C.Structure):
    _fields_ = [("fwd_lut", C.c_uint16 * 1024), ("bwd_lut", C.c_uint16 * 1024), ("wnd_bnd", C.c_uint16 * 17)]

    def np(self, name):
        return np.ctypeslib.as_array(getattr(self, name))


def lmcs_build_luts(d, bitdepth=10):
    luts = LmcsLuts()
    data = LmcsData.from_dict(d)
    _check(None, lib().ovgpu_lmcs_build_luts(C.byref(luts), C.byref(data), bitdepth), "ovgpu_lmcs_build_luts")
    return luts


class Lmcs:
    def __init__(self, ctx, luts=None):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_lmcs_create(ctx.h, C.byref(self.p)), "lmcs_create")
        if luts is not None:
            self.upload(luts)

    def upload(self, luts, ctx=None):
        self._keep = luts
        _check((ctx or self.ctx).h, lib().ovgpu_lmcs_upload((ctx or self.ctx).h, self.p, C.byref(luts)), "lmcs_upload")

    def set_data(self, data):
        """LmcsData of the slice (window range, chroma scaling offset) for the on-device chroma scale"""
        _check(self.ctx.h, lib().ovgpu_lmcs_set_data(self.p, C.byref(data)), "lmcs_set_data")

    def reshape(self, pic, forward=False, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_lmcs_reshape_picture((ctx or self.ctx).h, pic.p, self.p, int(forward)), "lmcs_reshape_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_lmcs_destroy(self.p)
            self.p = vp()


class Window(C.Structure):
    _fields_ = [("offset_lft", C.c_int32), ("offset_rgt", C.c_int32), ("offset_abv", C.c_int32), ("offset_blw", C.c_int32)]


class Md5:
    """ovgpu_md5_*: running MD5 of the output byte stream (dectest's YUV file), computed on the device"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_md5_create(ctx.h, C.byref(self.p)), "md5_create")

    def reset(self):
        _check(self.ctx.h, lib().ovgpu_md5_reset(self.ctx.h, self.p), "md5_reset")

    def picture(self, pic, window=None, ctx=None):
        win = Window(*window) if window is not None else None
        _check((ctx or self.ctx).h, lib().ovgpu_md5_picture((ctx or self.ctx).h, self.p, pic.p, C.byref(win) if win is not None else None), "md5_picture")

    def hexdigest(self, ctx=None):
        out = (C.c_uint8 * 16)()
        _check((ctx or self.ctx).h, lib().ovgpu_md5_final((ctx or self.ctx).h, self.p, out), "md5_final")
        return bytes(out).hex()

    def destroy(self):
        if self.p:
            lib().ovgpu_md5_destroy.restype = None
            lib().ovgpu_md5_destroy(self.p)
            self.p = vp()


def download_window(ctx, pic, window, w, h):
    """ovgpu_picture_download_window: the planes dectest would write for output window (lft, rgt, abv, blw) in chroma units"""
    l, r, a, b = window
    out = [np.zeros((h - 2 * (a + b), w - 2 * (l + r)), np.uint16), np.zeros((h // 2 - (a + b), w // 2 - (l + r)), np.uint16),
           np.zeros((h // 2 - (a + b), w // 2 - (l + r)), np.uint16)]
    win = Window(l, r, a, b)
    _check(ctx.h, lib().ovgpu_picture_download_window(ctx.h, pic.p, C.byref(win), *_host_planes(out)), "download_window")
    ctx.sync()
    return out


class ScalingWindow(C.Structure):
    """OVGPUScalingWindow: struct ScalingInfo (ovdpb.h:85-93)"""
    _fields_ = [("left", C.c_uint16), ("right", C.c_uint16), ("top", C.c_uint16), ("bottom", C.c_uint16),
                ("chroma_hor_col", C.c_uint8), ("chroma_ver_col", C.c_uint8)]


def upscale(ctx, dst, src, window=(0, 0, 0, 0), hor_col=1, ver_col=0):
    """ovgpu_picture_upscale: src resampled into dst (the sequence's maximum size), as pp_process_frame does for output frames"""
    win = ScalingWindow(*[int(v) for v in window], int(hor_col), int(ver_col))
    _check(ctx.h, lib().ovgpu_picture_upscale(ctx.h, dst.p, src.p, C.byref(win)), "picture_upscale")


class FilmGrainSei(C.Structure):
    """OVGPUFilmGrain: the fields of the film grain characteristics SEI the synthesis reads"""
    _fields_ = [("log2_scale_factor", C.c_uint8), ("comp_model_present_flag", C.c_uint8 * 3), ("num_model_values_minus1", C.c_uint8 * 3),
                ("intensity_interval_lower_bound", C.c_uint8 * 8 * 3), ("intensity_interval_upper_bound", C.c_uint8 * 8 * 3),
                ("comp_model_value", C.c_int16 * 3 * 8 * 3)]


class FilmGrainSynth:
    """ovgpu_film_grain_*: the grain patterns in HBM; picture(dst, src, sei, poc) = dst <- src + grain"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_film_grain_create(ctx.h, C.byref(self.p)), "film_grain_create")

    def picture(self, dst, src, sei, poc):
        g = FilmGrainSei()
        C.memmove(C.byref(g), C.byref(sei), C.sizeof(g))
        _check(self.ctx.h, lib().ovgpu_film_grain_picture(self.ctx.h, self.p, dst.p, src.p, C.byref(g), int(poc)), "film_grain_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_film_grain_destroy(self.p)
            self.p = vp()


class Dist:
    """ovgpu_dist_*: the reference hand-off between the ranks of one node (NCCL send / receive behind the C ABI)"""

    ID_BYTES = 128

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * Dist.ID_BYTES)()
        ret = lib().ovgpu_dist_unique_id(buf, Dist.ID_BYTES)
        if ret:
            raise OVGPUError("ovgpu_dist_unique_id failed: %d" % ret)
        return bytes(buf)

    def __init__(self, ctx, nb_ranks, rank, uid):
        self.ctx = ctx
        self.p = vp()
        buf = (C.c_uint8 * Dist.ID_BYTES).from_buffer_copy(uid)
        _check(ctx.h, lib().ovgpu_dist_create(ctx.h, int(nb_ranks), int(rank), buf, C.byref(self.p)), "dist_create")

    def publish(self, pic, consumers, ctx=None):
        arr = (C.c_int * len(consumers))(*consumers)
        _check((ctx or self.ctx).h, lib().ovgpu_picture_publish((ctx or self.ctx).h, self.p, pic.p, arr, len(consumers)), "picture_publish")

    def consume(self, pic, owner, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_picture_consume((ctx or self.ctx).h, self.p, pic.p, int(owner)), "picture_consume")

    def traffic(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().ovgpu_dist_traffic(self.p, C.byref(a), C.byref(b))
        return a.value, b.value

    def destroy(self):
        if self.p:
            lib().ovgpu_dist_destroy.restype = None
            lib().ovgpu_dist_destroy(self.p)
            self.p = vp()


class McBatch:
    """One picture's PU list on the device (ovgpu_mc_batch_*)."""

    def __init__(self, ctx, max_cmds, max_items):
        self.ctx = ctx
        self.p = vp()
        self.max_cmds = int(max_cmds)
        _check(ctx.h, lib().ovgpu_mc_batch_create(ctx.h, int(max_cmds), int(max_items), C.byref(self.p)), "mc_batch_create")

    def upload(self, cmds, ctx=None, prof=None):
        """prof: int16 [n][4][16] PROF table (OVGPUProfInfo) indexed by tool-2 commands"""
        cmds = np.ascontiguousarray(cmds, dtype=MC_CMD)
        prof = np.ascontiguousarray(prof, dtype=np.int16) if prof is not None else None
        self._keep = (cmds, prof)
        _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_upload2((ctx or self.ctx).h, self.p, vp(cmds.ctypes.data), len(cmds),
                                                                 vp(prof.ctypes.data) if prof is not None else None,
                                                                 len(prof) if prof is not None else 0), "mc_batch_upload2")

    def set_scaling(self, scale=None, hor_col=1, ver_col=0, ctx=None):
        """ovgpu_mc_batch_set_scaling: scale[list][ref_idx] = (hor, ver) factors of the NEXT upload / commit; None = no reference scaled"""
        if scale is None:
            _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_set_scaling((ctx or self.ctx).h, self.p, None), "mc_batch_set_scaling")
            return
        sc = McScaling()
        for l in range(2):
            for i, (sh, sv) in enumerate(scale[l]):
                sc.scale[l][i][0] = int(sh); sc.scale[l][i][1] = int(sv)
        sc.chroma_hor_col = int(hor_col); sc.chroma_ver_col = int(ver_col)
        _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_set_scaling((ctx or self.ctx).h, self.p, C.byref(sc)), "mc_batch_set_scaling")

    def staging(self, ctx=None):
        """numpy views of the batch's PINNED command and PROF-table buffers (ovgpu_mc_batch_staging)"""
        pc, pp = vp(), vp()
        _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_staging((ctx or self.ctx).h, self.p, C.byref(pc), C.byref(pp)), "mc_batch_staging")
        cmds = np.frombuffer((C.c_uint8 * (MC_CMD.itemsize * self.max_cmds)).from_address(pc.value), dtype=MC_CMD)
        n_prof = self.max_cmds // 16 + 1
        prof = np.frombuffer((C.c_int16 * (64 * n_prof)).from_address(pp.value), dtype=np.int16).reshape(n_prof, 4, 16)
        return cmds, prof

    def read_dmvr(self, n, ctx=None):
        """refined MVs (OVGPUDmvrMv) of the DMVR commands of the last run, in command order"""
        out = np.zeros(n, dtype=DMVR_MV)
        ret = lib().ovgpu_mc_batch_read_dmvr((ctx or self.ctx).h, self.p, vp(out.ctypes.data), int(n))
        if ret < 0:
            _check((ctx or self.ctx).h, ret, "mc_batch_read_dmvr")
        assert ret == n, "batch holds %d DMVR commands, asked for %d" % (ret, n)
        return out

    def status(self, ctx=None):
        """ovgpu_mc_batch_status: raises what the device-side walk of the last run refused (blocking)"""
        _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_status((ctx or self.ctx).h, self.p), "mc_batch_status")

    def commit(self, nb_cmds, nb_prof=0, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_mc_batch_commit((ctx or self.ctx).h, self.p, int(nb_cmds), int(nb_prof)), "mc_batch_commit")

    def run(self, dst, ref_lists, pocs, lmcs=None, ctx=None):
        """ref_lists: [rpl0, rpl1] lists of Picture; pocs likewise."""
        r = McRefs()
        for l in range(2):
            for i, pic in enumerate(ref_lists[l]):
                r.rpl[l][i] = pic.p
                r.poc[l][i] = int(pocs[l][i])
        _check((ctx or self.ctx).h, lib().ovgpu_mc_picture((ctx or self.ctx).h, dst.p, C.byref(r), self.p, lmcs.p if lmcs else None), "mc_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_mc_batch_destroy(self.p)
            self.p = vp()


class TuSlice(C.Structure):
    _fields_ = [("dep_quant", C.c_uint8), ("sh_ts_disabled", C.c_uint8), ("ict_type", C.c_uint8), ("pad", C.c_uint8)]


class TuBatch:
    """One picture's transform units + coefficient arena on the device (ovgpu_tu_batch_*)."""

    def __init__(self, ctx, max_cmds, max_coeffs):
        self.ctx = ctx
        self.p = vp()
        self.max_cmds, self.max_coeffs = int(max_cmds), int(max_coeffs)
        _check(ctx.h, lib().ovgpu_tu_batch_create(ctx.h, int(max_cmds), C.c_size_t(int(max_coeffs)), C.byref(self.p)), "tu_batch_create")

    def staging(self, ctx=None):
        """numpy views of the batch's PINNED command / coefficient buffers (ovgpu_tu_batch_staging)."""
        from .synth_tu import TU_CMD
        pc, pk = vp(), vp()
        _check((ctx or self.ctx).h, lib().ovgpu_tu_batch_staging((ctx or self.ctx).h, self.p, C.byref(pc), C.byref(pk)), "tu_batch_staging")
        cmds = np.frombuffer((C.c_uint8 * (TU_CMD.itemsize * self.max_cmds)).from_address(pc.value), dtype=TU_CMD)
        coeffs = np.frombuffer((C.c_int16 * self.max_coeffs).from_address(pk.value), dtype=np.int16)
        return cmds, coeffs

    def commit(self, nb_cmds, nb_coeffs, dep_quant=1, sh_ts_disabled=0, ict_type=0, ctx=None):
        sl = TuSlice(int(dep_quant), int(sh_ts_disabled), int(ict_type), 0)
        _check((ctx or self.ctx).h, lib().ovgpu_tu_batch_commit((ctx or self.ctx).h, self.p, int(nb_cmds), C.c_size_t(int(nb_coeffs)), C.byref(sl)), "tu_batch_commit")

    def upload(self, cmds, coeffs, dep_quant=1, sh_ts_disabled=0, ict_type=0, ctx=None):
        from .synth_tu import TU_CMD
        cmds = np.ascontiguousarray(cmds, dtype=TU_CMD)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.int16)
        sl = TuSlice(int(dep_quant), int(sh_ts_disabled), int(ict_type), 0)
        self._keep = (cmds, coeffs)
        _check((ctx or self.ctx).h, lib().ovgpu_tu_batch_upload((ctx or self.ctx).h, self.p, vp(cmds.ctypes.data), len(cmds), vp(coeffs.ctypes.data),
                                                       C.c_size_t(len(coeffs)), C.byref(sl)), "tu_batch_upload")

    def set_anchors(self, anchors, ctx=None):
        """OVGPULmcsAnchor list for the on-device LMCS chroma scale; call before the commands that name them"""
        anchors = np.ascontiguousarray(anchors)
        assert anchors.dtype.itemsize == 8
        _check((ctx or self.ctx).h, lib().ovgpu_tu_batch_set_anchors((ctx or self.ctx).h, self.p, vp(anchors.ctypes.data), len(anchors)), "tu_batch_set_anchors")

    def run(self, pic, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_residual_picture((ctx or self.ctx).h, pic.p, self.p), "residual_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_tu_batch_destroy(self.p)
            self.p = vp()


class DbfMaps(C.Structure):
    """OVGPUDbfMaps: picture-level deblocking maps (layout in include/ovvc_b200.h)."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("log2_ctu_s", C.c_int32),
                ("beta_offset", C.c_int16), ("tc_offset", C.c_int16),
                ("ver", C.POINTER(C.c_uint32) * 9), ("hor", C.POINTER(C.c_uint32) * 9), ("qp", C.POINTER(C.c_uint8) * 3)]

    @classmethod
    def from_dict(cls, m):
        """m: dict from synth_dbf.dbf_maps -- ver [9][nch][w_u+1] u32, hor [9][ncw][h_u+1] u32, qp [3][h_u][w_u] u8"""
        s = cls()
        s.width, s.height, s.log2_ctu_s = m["width"], m["height"], m["log2_ctu"]
        s.beta_offset, s.tc_offset = m["beta_offset"], m["tc_offset"]
        for k in range(9):
            s.ver[k] = m["ver"][k].ctypes.data_as(C.POINTER(C.c_uint32))
            s.hor[k] = m["hor"][k].ctypes.data_as(C.POINTER(C.c_uint32))
        for k in range(3):
            s.qp[k] = m["qp"][k].ctypes.data_as(C.POINTER(C.c_uint8))
        s._keep = m
        return s


class DbfParams:
    def __init__(self, ctx, w, h, log2_ctu=7):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_dbf_params_create(ctx.h, w, h, log2_ctu, C.byref(self.p)), "dbf_params_create")

    def upload(self, maps, ctx=None):
        st = DbfMaps.from_dict(maps) if isinstance(maps, dict) else maps
        self._keep = st
        _check((ctx or self.ctx).h, lib().ovgpu_dbf_params_upload((ctx or self.ctx).h, self.p, C.byref(st)), "dbf_params_upload")

    def staging(self, ctx=None):
        """numpy views (ver [9][nch][w_u+1], hor [9][ncw][h_u+1], qp [3][h_u][w_u]) of the PINNED map buffer."""
        st = DbfMaps()
        _check((ctx or self.ctx).h, lib().ovgpu_dbf_params_staging((ctx or self.ctx).h, self.p, C.byref(st)), "dbf_params_staging")
        n_u = (1 << st.log2_ctu_s) >> 2
        w_u, h_u = st.width >> 2, st.height >> 2
        ncw, nch = -(-w_u // n_u), -(-h_u // n_u)
        u32 = lambda ptr, n: np.frombuffer((C.c_uint32 * n).from_address(C.addressof(ptr.contents)), dtype=np.uint32)
        ver = [u32(st.ver[k], nch * (w_u + 1)).reshape(nch, w_u + 1) for k in range(9)]
        hor = [u32(st.hor[k], ncw * (h_u + 1)).reshape(ncw, h_u + 1) for k in range(9)]
        qp = [np.frombuffer((C.c_uint8 * (w_u * h_u)).from_address(C.addressof(st.qp[k].contents)), dtype=np.uint8).reshape(h_u, w_u)
              for k in range(3)]
        return dict(ver=ver, hor=hor, qp=qp)

    def commit(self, beta_offset=0, tc_offset=0, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_dbf_params_commit((ctx or self.ctx).h, self.p, int(beta_offset), int(tc_offset)), "dbf_params_commit")

    def run(self, pic, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_dbf_picture((ctx or self.ctx).h, pic.p, self.p), "dbf_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_dbf_params_destroy(self.p)
            self.p = vp()


class IntraBatch:
    """Luma intra blocks of one picture, sorted into dependency levels on the host (ovgpu_intra_batch_*)."""

    def __init__(self, ctx, max_cmds):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_intra_batch_create(ctx.h, int(max_cmds), C.byref(self.p)), "intra_batch_create")

    def upload(self, cmds, w, h, log2_ctu=7, ctx=None, isp=None):
        """isp: OVGPUIspInfo array (synth_intra.ISP_WORDS) the ISP commands index through pad[0..1]"""
        from .synth_intra import INTRA_CMD
        cmds = np.ascontiguousarray(cmds, dtype=INTRA_CMD)
        isp = np.ascontiguousarray(isp) if isp is not None and len(isp) else None
        assert isp is None or isp.dtype.itemsize == 40
        self._keep = (cmds, isp)
        _check((ctx or self.ctx).h, lib().ovgpu_intra_batch_upload2((ctx or self.ctx).h, self.p, vp(cmds.ctypes.data), len(cmds),
                                                                    vp(isp.ctypes.data) if isp is not None else None, len(isp) if isp is not None else 0,
                                                                    w, h, log2_ctu), "intra_batch_upload2")

    def levels(self):
        return int(lib().ovgpu_intra_batch_levels(self.p))

    def run(self, pic, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_intra_picture((ctx or self.ctx).h, pic.p, self.p), "intra_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_intra_batch_destroy(self.p)
            self.p = vp()


def recon_picture(ctx, pic, tu=None, intra=None, mc=None, ref_lists=None, pocs=None, lmcs=None):
    """ovgpu_recon_picture: motion compensation, then residual / intra prediction in the reference's CU order"""
    r = McRefs()
    if mc is not None:
        for l in range(2):
            for i, rp in enumerate(ref_lists[l]):
                r.rpl[l][i] = rp.p
                r.poc[l][i] = int(pocs[l][i])
    _check(ctx.h, lib().ovgpu_recon_picture(ctx.h, pic.p, C.byref(r) if mc is not None else None, mc.p if mc is not None else None,
                                            lmcs.p if lmcs else None, tu.p if tu is not None else None,
                                            intra.p if intra is not None else None), "recon_picture")


def mc_nb_items(cmds):
    """work items a PU list expands to: <=16x16 blocks of plain commands + one per BDOF / PROF command"""
    w = (1 << cmds["log2_w"].astype(np.int64)); h = (1 << cmds["log2_h"].astype(np.int64))
    n = ((w + 15) >> 4) * ((h + 15) >> 4)
    return int(np.where((cmds["tool"] == 0) | (cmds["tool"] == 4), n, 1).sum())


class TileGrid(C.Structure):
    """OVGPUTileGrid: tile boundaries in CTUs"""
    _fields_ = [("nb_cols", C.c_int32), ("nb_rows", C.c_int32), ("col_bd", C.c_uint16 * 22), ("row_bd", C.c_uint16 * 22)]


class SaoParams:
    def __init__(self, ctx, w, h, log2_ctu=7):
        self.ctx = ctx
        self.p = vp()
        _check(ctx.h, lib().ovgpu_sao_params_create(ctx.h, w, h, log2_ctu, C.byref(self.p)), "sao_params_create")

    def upload(self, sao_ctu, ctx=None):
        sao_ctu = np.ascontiguousarray(sao_ctu, dtype=SAO_CTU)
        self._keep = sao_ctu
        _check((ctx or self.ctx).h, lib().ovgpu_sao_params_upload((ctx or self.ctx).h, self.p, vp(sao_ctu.ctypes.data)), "sao_params_upload")

    def set_tiles(self, col_bd=None, row_bd=None, ctx=None):
        """ovgpu_sao_params_set_tiles: tile boundaries in CTUs (None: one tile)"""
        tg = None
        if col_bd is not None:
            tg = TileGrid()
            tg.nb_cols, tg.nb_rows = len(col_bd) - 1, len(row_bd) - 1
            tg.col_bd[:len(col_bd)] = [int(v) for v in col_bd]
            tg.row_bd[:len(row_bd)] = [int(v) for v in row_bd]
        _check((ctx or self.ctx).h, lib().ovgpu_sao_params_set_tiles((ctx or self.ctx).h, self.p, C.byref(tg) if tg is not None else None), "sao_params_set_tiles")

    def run(self, dst, src, ctx=None):
        _check((ctx or self.ctx).h, lib().ovgpu_sao_picture((ctx or self.ctx).h, dst.p, src.p, self.p), "sao_picture")

    def destroy(self):
        if self.p:
            lib().ovgpu_sao_params_destroy(self.p)
            self.p = vp()
