"""The whole hot path on one picture, as bench.py runs it (the same workload generator and the same calls): motion
compensation -> residual + intra in reconstruction order (LMCS chroma scales derived on the way) -> LMCS backward reshape ->
deblocking -> SAO -> ALF through the C ABI, against the same chain of oracle ports (each pinned to the compiled reference by
its own test file).  This is the "MD5 of the output picture" check of BASELINE.json on the benchmark's own input."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


from helpers_chain import oracle_picture  # noqa: E402,F401


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["1080p", "4k"])
def test_gpu_whole_picture_matches_oracle_chain(oracle, gpu_ctx, size):
    from openvvc_b200 import api
    w, h = bench.SIZES[size]
    wl = bench.make_workload(w, h, 1, seed0=7100)
    p = wl["pics"][0]
    assert ((p["icmds"]["flags"] & 32) != 0).any() and (p["tus"]["bdpcm"] != 0).any() and (p["cmds"]["tool"] != 0).any()
    ctx = gpu_ctx
    refs = [[ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
    for lst, src in zip(refs, wl["refs"]):
        for pic, planes in zip(lst, src):
            pic.upload(planes)
    lmcs = api.Lmcs(ctx, wl["lmcs"])
    lmcs.set_data(wl["lmcs_data"])
    d = bench.make_dev(api, ctx, p, w, h)
    for _ in range(2):                       # twice: the second pass reuses the bound schedule and the dataflow epoch
        bench.run_stages(api, ctx, d, refs, wl["pocs"], lmcs)
    got = d["dst"].download()
    want = oracle_picture(oracle, wl, p)
    helpers.assert_planes_equal(got, want, "whole picture (%s): C ABI chain vs oracle chain" % size)
    assert len({helpers.md5(*got), helpers.md5(*want)}) == 1
    for k in ("rec", "mid", "dst", "alf", "sao", "mc", "tu", "dbf", "intra"):
        d[k].destroy()
    for lst in refs:
        for pic in lst:
            pic.destroy()


@pytest.mark.parametrize("w,h", [(416, 240), (832, 480)])
def test_emulated_whole_picture_matches_oracle_chain(oracle, emu, w, h):
    """the same chain with the kernel sources emulated on the CPU (tests/emu): what the GPU test checks, minus the hardware"""
    import helpers_df
    import helpers_mc
    po = oracle
    wl = bench.make_workload(w, h, 1, seed0=7200)
    p = wl["pics"][0]
    luts = po.LmcsLuts()
    C.memmove(C.byref(luts), C.byref(wl["lmcs"]), C.sizeof(luts))
    t = po.AlfTables()
    C.memmove(C.byref(t), C.byref(p["tab"]), C.sizeof(t))
    rec = helpers_mc.emu_mc(emu, po, w, h, wl["refs"][0], wl["refs"][1], wl["pocs"], p["cmds"].astype(po.McCmd), luts, prof=p["prof"])
    ptrs = (C.c_void_p * 3)(*[x.ctypes.data for x in rec])
    strides = (C.c_int * 3)(*[x.strides[0] // 2 for x in rec])
    data = po.make_lmcs_data(wl["lmcs_dict"])
    isp = np.ascontiguousarray(p["isp_words"])
    nl = emu.emu_recon_picture(ptrs, strides, w, h, 7, 10, p["icmds"].ctypes.data_as(C.c_void_p), len(p["icmds"]),
                               p["tus"].ctypes.data_as(C.c_void_p), len(p["tus"]), p["coeffs"].ctypes.data_as(C.c_void_p),
                               C.c_size_t(len(p["coeffs"])), 1, 0, 1, p["anchors"].ctypes.data_as(C.c_void_p), len(p["anchors"]),
                               C.byref(luts), C.byref(data), isp.ctypes.data_as(C.c_void_p), len(isp))
    assert nl >= 0, nl
    helpers.assert_planes_equal(rec, oracle_picture(po, wl, p, stop_after="recon"), "emulated MC + reconstruction vs oracle")
    emu.emu_lmcs_reshape(rec[0].ctypes.data_as(po.u16p), rec[0].strides[0] // 2, w, h, 10, luts.np("bwd_lut").ctypes.data_as(po.u16p))
    out = helpers_df.emu_dbf(emu, rec, p["dbf"])
    out = helpers.emu_sao(emu, out, p["sao"].astype(po.SaoCtu))
    out, _ = helpers.emu_alf(emu, out, t, p["ctu"])
    helpers.assert_planes_equal(out, oracle_picture(po, wl, p), "emulated whole picture vs oracle chain")
