"""Motion compensation + LMCS: oracle port vs golden / compiled reference (CPU), kernel-source emulation vs
oracle (CPU), and the sm_100a kernels through the C ABI vs oracle / golden (GPU)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers
import helpers_mc

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mc_golden.json")))
IDS = lambda c: "%dx%d_ctu%d%s" % (c["w"], c["h"], 1 << c["log2_ctu"], "_lmcs" if c["lmcs"] else "")


def _case(po, c):
    refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(po, c["w"], c["h"], c["seed"], c["log2_ctu"], c["lmcs"], c["min_log2"])
    assert helpers.md5(cmds, *refs0[0], *refs1[-1]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return refs0, refs1, pocs, cmds, luts


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(oracle, c):
    refs0, refs1, pocs, cmds, luts = _case(oracle, c)
    dst = helpers_mc.blank(c["w"], c["h"])
    oracle.mc_picture(dst, [refs0, refs1], pocs, cmds, c["log2_ctu"], impl="port", lmcs=luts)
    assert [helpers.md5(p) for p in dst] == c["md5"]


@pytest.mark.parametrize("w,h,log2,lmcs,min_log2", [(416, 240, 7, False, 3), (416, 240, 7, True, 2), (832, 480, 7, True, 3),
                                                     (264, 136, 6, False, 2), (72, 40, 5, True, 2)])
def test_port_vs_compiled_reference(oracle, w, h, log2, lmcs, min_log2):
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(po, w, h, 5 * w + h, log2, lmcs, min_log2)
    a, b = helpers_mc.blank(w, h), helpers_mc.blank(w, h)
    po.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts)
    po.mc_picture(b, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts)
    helpers.assert_planes_equal(a, b, "MC port vs reference")


def test_lmcs_luts_port_vs_reference_and_product(oracle):
    from openvvc_b200 import api, synth
    for seed in range(8):
        d = synth.lmcs_data(900 + seed)
        p = oracle.lmcs_build_luts(d, impl="port")
        g = api.lmcs_build_luts(d)
        for k in ("fwd_lut", "bwd_lut", "wnd_bnd"):
            assert (p.np(k) == g.np(k)).all(), k
            if oracle.ref_available():
                assert (p.np(k) == oracle.lmcs_build_luts(d, impl="ref").np(k)).all(), k


@pytest.mark.parametrize("w,h,log2,lmcs,min_log2", [(416, 240, 7, False, 3), (416, 240, 7, True, 2), (256, 128, 6, False, 3),
                                                     (136, 72, 5, True, 2), (832, 480, 7, False, 2)])
def test_emulated_kernel_matches_oracle(oracle, emu, w, h, log2, lmcs, min_log2):
    refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(oracle, w, h, 7 * w + h, log2, lmcs, min_log2)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts)
    helpers.assert_planes_equal(helpers_mc.emu_mc(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts), a, "emulated MC kernel vs oracle")


def test_emulated_lmcs_reshape(oracle, emu):
    from openvvc_b200 import synth
    luts = oracle.lmcs_build_luts(synth.lmcs_data(3))
    pl = synth.planes(264, 136, 4)[0]
    want = pl.copy()
    oracle.port().ovo_lmcs_reshape_plane(want.ctypes.data_as(oracle.u16p), want.strides[0] // 2, 264, 136,
                                         luts.np("bwd_lut").ctypes.data_as(oracle.u16p), 10)
    got = pl.copy()
    emu.emu_lmcs_reshape(got.ctypes.data_as(oracle.u16p), got.strides[0] // 2, 264, 136, 10, luts.np("bwd_lut").ctypes.data_as(oracle.u16p))
    assert (got == want).all()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_matches_golden(oracle, gpu_ctx, c):
    refs0, refs1, pocs, cmds, luts = _case(oracle, c)
    out = helpers_mc.gpu_mc(gpu_ctx, c["w"], c["h"], refs0, refs1, pocs, cmds, luts)
    assert [helpers.md5(p) for p in out] == c["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,lmcs,min_log2", [(416, 240, 7, False, 2), (264, 136, 6, True, 2), (1920, 1080, 7, True, 3)])
def test_gpu_matches_oracle(oracle, gpu_ctx, w, h, log2, lmcs, min_log2):
    refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(oracle, w, h, 9 * w + h, log2, lmcs, min_log2)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts)
    helpers.assert_planes_equal(helpers_mc.gpu_mc(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts), a, "CUDA MC vs oracle")


@pytest.mark.gpu
def test_gpu_full_size_4k_vs_reference(oracle, gpu_ctx):
    """BASELINE.json config 2 geometry (3840x2160): CUDA vs the compiled reference (SIMD table) when present."""
    w, h = 3840, 2160
    refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(oracle, w, h, 77, 7, True, 3)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, 7, impl="ref" if oracle.ref_available() else "port", lmcs=luts)
    helpers.assert_planes_equal(helpers_mc.gpu_mc(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts), a, "CUDA MC vs reference at 4K")


@pytest.mark.gpu
def test_gpu_lmcs_backward_picture(oracle, gpu_ctx):
    from openvvc_b200 import api, synth
    w, h = 1920, 1080
    luts = oracle.lmcs_build_luts(synth.lmcs_data(11))
    planes = synth.planes(w, h, 12)
    want = planes[0].copy()
    oracle.port().ovo_lmcs_reshape_plane(want.ctypes.data_as(oracle.u16p), w, w, h, luts.np("bwd_lut").ctypes.data_as(oracle.u16p), 10)
    gl = api.LmcsLuts(); C.memmove(C.byref(gl), C.byref(luts), C.sizeof(gl))
    lm = api.Lmcs(gpu_ctx, gl)
    pic = gpu_ctx.picture(w, h); pic.upload(planes)
    lm.reshape(pic, forward=False)
    out = pic.download()
    assert (out[0] == want).all() and (out[1] == planes[1]).all()
    lm.destroy(); pic.destroy()


# ================================================================================================ BDOF / PROF
TOOLS_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mc_tools_golden.json")))


def _tools_case(po, c):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(po, c["w"], c["h"], c["seed"], c["log2_ctu"], c["lmcs"], c["min_log2"])
    assert helpers.md5(cmds, prof, *refs0[0], *refs1[-1]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return refs0, refs1, pocs, cmds, luts, prof


@pytest.mark.parametrize("c", TOOLS_GOLDEN, ids=IDS)
def test_tools_port_matches_golden(oracle, c):
    refs0, refs1, pocs, cmds, luts, prof = _tools_case(oracle, c)
    assert (cmds["tool"] == 1).sum() > 10 and (cmds["tool"] == 2).sum() > 10
    dst = helpers_mc.blank(c["w"], c["h"])
    oracle.mc_picture(dst, [refs0, refs1], pocs, cmds, c["log2_ctu"], impl="port", lmcs=luts, prof=prof)
    assert [helpers.md5(p) for p in dst] == c["md5"]


@pytest.mark.parametrize("w,h,log2,lmcs", [(416, 240, 7, False), (416, 240, 7, True), (832, 480, 7, True), (264, 136, 6, False)])
def test_tools_port_vs_compiled_reference(oracle, w, h, log2, lmcs):
    """BDOF (rcn_bdof_mcp_l) and PROF (rcn_prof_mcp_b_l): port vs the reference's scalar C and its SSE/AVX2 table"""
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(po, w, h, 5 * w + h + 1, log2, lmcs)
    a = helpers_mc.blank(w, h)
    po.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts, prof=prof)
    for simd in (0, 1):
        b = helpers_mc.blank(w, h)
        po.mc_picture(b, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts, prof=prof, simd=simd)
        helpers.assert_planes_equal(a, b, "MC tools port vs reference (simd=%d)" % simd)


@pytest.mark.parametrize("w,h,log2,lmcs", [(416, 240, 7, False), (416, 240, 7, True), (256, 128, 6, False), (832, 480, 7, True)])
def test_tools_emulated_kernels_match_oracle(oracle, emu, w, h, log2, lmcs):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, w, h, 7 * w + h + 1, log2, lmcs)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts, prof=prof)
    helpers.assert_planes_equal(helpers_mc.emu_mc(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts, prof=prof), a,
                                "emulated MC/BDOF/PROF kernels vs oracle")


def test_tools_malformed_commands_rejected(oracle, emu):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, 256, 128, 3, 7, False)
    bad = cmds.copy()
    i = int(np.nonzero(bad["tool"] == 2)[0][0])
    bad["aux"][i] = (len(prof) << 8) | 1            # PROF table index out of range
    with pytest.raises(AssertionError):
        helpers_mc.emu_mc(emu, oracle, 256, 128, refs0, refs1, pocs, bad, luts, prof=prof)
    bad = cmds.copy()
    i = int(np.nonzero(bad["tool"] == 1)[0][0])
    bad["inter_dir"][i] = 1                         # BDOF needs bi-prediction
    with pytest.raises(AssertionError):
        helpers_mc.emu_mc(emu, oracle, 256, 128, refs0, refs1, pocs, bad, luts, prof=prof)


@pytest.mark.gpu
def test_gpu_malformed_commands_rejected(oracle, gpu_ctx):
    """The command walk (validation, routing, work-item lists) runs on the GPU: a malformed command is reported, with its index,
    by the call that would run the batch, nothing is launched for it, and the batch is usable again afterwards."""
    from openvvc_b200 import api
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, 256, 128, 3, 7, False)
    good = helpers_mc.gpu_mc(gpu_ctx, 256, 128, refs0, refs1, pocs, cmds, luts, prof=prof)
    for tool, field, val in ((2, "aux", (len(prof) << 8) | 1), (1, "inter_dir", 1), (0, "log2_w", 9)):
        bad = cmds.copy()
        i = int(np.nonzero(bad["tool"] == tool)[0][-1])
        bad[field][i] = val
        with pytest.raises(api.OVGPUError, match="MC command %d malformed" % i):
            helpers_mc.gpu_mc(gpu_ctx, 256, 128, refs0, refs1, pocs, bad, luts, prof=prof)
    helpers.assert_planes_equal(helpers_mc.gpu_mc(gpu_ctx, 256, 128, refs0, refs1, pocs, cmds, luts, prof=prof), good, "after the rejected batches")


@pytest.mark.gpu
@pytest.mark.parametrize("c", TOOLS_GOLDEN, ids=IDS)
def test_tools_gpu_matches_golden(oracle, gpu_ctx, c):
    refs0, refs1, pocs, cmds, luts, prof = _tools_case(oracle, c)
    out = helpers_mc.gpu_mc(gpu_ctx, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, prof=prof)
    assert [helpers.md5(p) for p in out] == c["md5"]


@pytest.mark.gpu
def test_tools_gpu_full_size_4k_vs_reference(oracle, gpu_ctx):
    """BASELINE.json config 4 flavour (affine + BDOF heavy inter) at 3840x2160: CUDA vs the compiled reference"""
    w, h = 3840, 2160
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, w, h, 81, 7, True, 3)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, 7, impl="ref" if oracle.ref_available() else "port", lmcs=luts, prof=prof)
    helpers.assert_planes_equal(helpers_mc.gpu_mc(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts, prof=prof), a, "CUDA MC tools vs reference at 4K")


# ================================================================================================ GPM
GPM_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mc_gpm_golden.json")))


def _gpm_case(po, c):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(po, c["w"], c["h"], c["seed"], c["log2_ctu"], c["lmcs"], c["min_log2"], p_gpm=0.6)
    assert helpers.md5(cmds, prof, *refs0[0], *refs1[-1]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return refs0, refs1, pocs, cmds, luts, prof


@pytest.mark.parametrize("c", GPM_GOLDEN, ids=IDS)
def test_gpm_port_and_emulated_kernel_match_golden(oracle, emu, c):
    """rcn_gpm_b (a15): two uni-predictions from any list blended with the 64 geometric splits, luma and chroma, LMCS"""
    refs0, refs1, pocs, cmds, luts, prof = _gpm_case(oracle, c)
    g = cmds[cmds["tool"] == 4]
    assert len(g) > 10 and len(set((g["aux"] & 63).tolist())) > 10 and set(((g["aux"] >> 8) & 3).tolist()) == {0, 1, 2, 3}
    dst = helpers_mc.blank(c["w"], c["h"])
    oracle.mc_picture(dst, [refs0, refs1], pocs, cmds, c["log2_ctu"], impl="port", lmcs=luts, prof=prof)
    assert [helpers.md5(p) for p in dst] == c["md5"]
    if c["w"] <= 832:
        helpers.assert_planes_equal(helpers_mc.emu_mc(emu, oracle, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, prof=prof), dst,
                                    "emulated GPM kernel vs oracle")


@pytest.mark.parametrize("w,h,log2,lmcs", [(416, 240, 7, False), (832, 480, 7, True), (264, 136, 6, False), (200, 120, 5, True)])
def test_gpm_port_vs_compiled_reference(oracle, w, h, log2, lmcs):
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(po, w, h, 3 * w + h + 2, log2, lmcs, p_gpm=0.6)
    a = helpers_mc.blank(w, h)
    po.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts, prof=prof)
    for simd in (0, 1):
        b = helpers_mc.blank(w, h)
        po.mc_picture(b, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts, prof=prof, simd=simd)
        helpers.assert_planes_equal(a, b, "GPM port vs reference (simd=%d)" % simd)


def test_gpm_malformed_commands_rejected(oracle, emu):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, 256, 128, 3, 7, False, p_gpm=0.8)
    i = int(np.nonzero(cmds["tool"] == 4)[0][0])
    for field, val in (("inter_dir", 1), ("flags", 1), ("aux", 1 << 12), ("bcw_idx_plus1", 1), ("log2_w", 7)):
        bad = cmds.copy()
        bad[field][i] = val
        with pytest.raises(AssertionError):
            helpers_mc.emu_mc(emu, oracle, 256, 128, refs0, refs1, pocs, bad, luts, prof=prof)


@pytest.mark.gpu
@pytest.mark.parametrize("c", GPM_GOLDEN, ids=IDS)
def test_gpm_gpu_matches_golden(oracle, gpu_ctx, c):
    refs0, refs1, pocs, cmds, luts, prof = _gpm_case(oracle, c)
    out = helpers_mc.gpu_mc(gpu_ctx, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, prof=prof)
    assert [helpers.md5(p) for p in out] == c["md5"]


# ================================================================================================ DMVR
DMVR_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mc_dmvr_golden.json")))
DIDS = lambda c: "%dx%d_ctu%d_%s%s" % (c["w"], c["h"], 1 << c["log2_ctu"], c["kind"], "_lmcs" if c["lmcs"] else "")


def _dmvr_case(po, c):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_dmvr_case(po, c["w"], c["h"], c["seed"], c["log2_ctu"], c["lmcs"], c["kind"])
    assert helpers.md5(cmds, prof, *refs0[0], *refs1[-1]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return refs0, refs1, pocs, cmds, luts, prof, np.zeros(c["nb_dmvr"], oracle_dmvr_dtype(po))


def oracle_dmvr_dtype(po):
    return po.DmvrMv


@pytest.mark.parametrize("c", DMVR_GOLDEN, ids=DIDS)
def test_dmvr_port_matches_golden(oracle, c):
    refs0, refs1, pocs, cmds, luts, prof, mvs = _dmvr_case(oracle, c)
    dst = helpers_mc.blank(c["w"], c["h"])
    oracle.mc_picture(dst, [refs0, refs1], pocs, cmds, c["log2_ctu"], impl="port", lmcs=luts, prof=prof, dmvr_out=mvs)
    assert [helpers.md5(p) for p in dst] == c["md5"] and helpers.md5(mvs) == c["mv_md5"]
    dm = cmds[cmds["tool"] == 3]
    assert ((mvs["mv0_x"] != dm["mv0_x"]) | (mvs["mv0_y"] != dm["mv0_y"])).mean() > 0.3      # the search really moves MVs


@pytest.mark.parametrize("w,h,log2,lmcs,kind", [(416, 240, 7, False, "natural"), (416, 240, 7, True, "blocky"), (832, 480, 7, True, "natural"),
                                                 (264, 136, 6, False, "natural")])
def test_dmvr_port_vs_compiled_reference(oracle, w, h, log2, lmcs, kind):
    """rcn_dmvr_mv_refine: pixels (luma + chroma) and refined MVs, port vs the reference's scalar C and SSE/AVX2 table"""
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_dmvr_case(po, w, h, 5 * w + h + 2, log2, lmcs, kind)
    nd = int((cmds["tool"] == 3).sum())
    ma = np.zeros(nd, po.DmvrMv)
    a = helpers_mc.blank(w, h)
    po.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts, prof=prof, dmvr_out=ma)
    for simd in (0, 1):
        mb = np.zeros(nd, po.DmvrMv)
        b = helpers_mc.blank(w, h)
        po.mc_picture(b, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts, prof=prof, simd=simd, dmvr_out=mb)
        helpers.assert_planes_equal(a, b, "DMVR port vs reference (simd=%d)" % simd)
        assert (ma == mb).all(), "refined MVs differ (simd=%d)" % simd


@pytest.mark.parametrize("w,h,log2,lmcs,kind", [(416, 240, 7, False, "natural"), (416, 240, 7, True, "blocky"), (256, 128, 6, True, "natural"),
                                                 (832, 480, 7, False, "natural")])
def test_dmvr_emulated_kernel_matches_oracle(oracle, emu, w, h, log2, lmcs, kind):
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_dmvr_case(oracle, w, h, 7 * w + h + 2, log2, lmcs, kind)
    nd = int((cmds["tool"] == 3).sum())
    ma, mb = np.zeros(nd, oracle.DmvrMv), np.zeros(nd, oracle.DmvrMv)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, log2, impl="port", lmcs=luts, prof=prof, dmvr_out=ma)
    b = helpers_mc.emu_mc(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts, prof=prof, dmvr_out=mb)
    helpers.assert_planes_equal(b, a, "emulated DMVR kernel vs oracle")
    assert (ma == mb).all()


@pytest.mark.gpu
@pytest.mark.parametrize("c", DMVR_GOLDEN, ids=DIDS)
def test_dmvr_gpu_matches_golden(oracle, gpu_ctx, c):
    from openvvc_b200 import api
    refs0, refs1, pocs, cmds, luts, prof, _ = _dmvr_case(oracle, c)
    mvs = np.zeros(c["nb_dmvr"], api.DMVR_MV)
    out = helpers_mc.gpu_mc(gpu_ctx, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, prof=prof, dmvr_out=mvs)
    assert [helpers.md5(p) for p in out] == c["md5"] and helpers.md5(mvs) == c["mv_md5"]


@pytest.mark.gpu
def test_dmvr_gpu_full_size_4k_vs_reference(oracle, gpu_ctx):
    """BASELINE.json config 4 flavour (affine + BDOF + DMVR heavy inter) at 3840x2160: CUDA vs the compiled reference"""
    from openvvc_b200 import api
    w, h = 3840, 2160
    refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_dmvr_case(oracle, w, h, 83, 7, True)
    nd = int((cmds["tool"] == 3).sum())
    ma, mb = np.zeros(nd, oracle.DmvrMv), np.zeros(nd, api.DMVR_MV)
    a = helpers_mc.blank(w, h)
    oracle.mc_picture(a, [refs0, refs1], pocs, cmds, 7, impl="ref" if oracle.ref_available() else "port", lmcs=luts, prof=prof, dmvr_out=ma)
    helpers.assert_planes_equal(helpers_mc.gpu_mc(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts, prof=prof, dmvr_out=mb), a, "CUDA DMVR vs reference at 4K")
    assert helpers.md5(ma) == helpers.md5(mb)


# ================================================================================================ reference picture resampling (a17)
RPR_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mc_rpr_golden.json")))
RIDS = lambda c: "%dx%d_ctu%d_%s%s" % (c["w"], c["h"], 1 << c["log2_ctu"], c["kind"], "_lmcs" if c["lmcs"] else "")


def _rpr_case(po, c):
    case = helpers_mc.rpr_case(po, c["w"], c["h"], c["seed"], c["log2_ctu"], c["lmcs"], c["kind"], c["hor_col"], c["ver_col"])
    refs0, refs1, pocs, cmds, luts, prof, sc = case
    assert helpers.md5(cmds, prof, *refs0[0], *refs1[-1]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return case


@pytest.mark.parametrize("c", RPR_GOLDEN, ids=RIDS)
def test_rpr_port_matches_golden(oracle, c):
    """rcn_mcp_rpr_* / rcn_mc_rpr_b_* / rcn_mc_rpr_prof_b_l (a17): plain uni / bi / BCW, GPM and affine 4x4 (PROF on the unscaled
    list of a mixed bi-prediction) from references 2x, 1.5x, 1x, 1/2 and 1/4 the picture's size"""
    refs0, refs1, pocs, cmds, luts, prof, sc = _rpr_case(oracle, c)
    used = set()
    for m in cmds:
        if m["inter_dir"] & 1:
            used.add((0, int(m["ref_idx0"])))
        if m["inter_dir"] & 2:
            used.add((1, int(m["ref_idx1"])))
    assert len(used) == 6 and set(cmds["tool"].tolist()) >= {0, 2, 4}
    dst = helpers_mc.blank(c["w"], c["h"])
    oracle.mc_picture_rpr(dst, [refs0, refs1], pocs, cmds, sc, c["log2_ctu"], impl="port", lmcs=luts, prof=prof)
    assert [helpers.md5(p) for p in dst] == c["md5"]


@pytest.mark.parametrize("w,h,log2,lmcs,kind,hc,vc", [(416, 240, 7, False, "natural", 1, 0), (320, 192, 7, True, "noise", 0, 0),
                                                      (264, 136, 6, True, "noise", 1, 1), (200, 120, 5, False, "natural", 0, 1)])
def test_rpr_port_vs_compiled_reference(oracle, w, h, log2, lmcs, kind, hc, vc):
    """scalar C templates (arithmetic of record) and the SSE4.1 / AVX2 table.  Noise planes reach the reference's unsigned read of
    a negative intermediate in put_vvc_pel_rpr_clip (rcn_mc.c:623-642), which the port reproduces."""
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    refs0, refs1, pocs, cmds, luts, prof, sc = helpers_mc.rpr_case(po, w, h, 5 * w + h + 1, log2, lmcs, kind, hc, vc)
    a = helpers_mc.blank(w, h)
    po.mc_picture_rpr(a, [refs0, refs1], pocs, cmds, sc, log2, impl="port", lmcs=luts, prof=prof)
    # the SSE / AVX2 kernels read that intermediate as signed, so they agree with the C templates only where it is not negative
    # (natural content); the C templates are the arithmetic of record
    for simd in ((0, 1) if kind == "natural" else (0,)):
        b = helpers_mc.blank(w, h)
        po.mc_picture_rpr(b, [refs0, refs1], pocs, cmds, sc, log2, impl="ref", lmcs=luts, prof=prof, simd=simd)
        helpers.assert_planes_equal(a, b, "RPR port vs reference (simd=%d)" % simd)


@pytest.mark.parametrize("c", [c for c in RPR_GOLDEN if c["w"] <= 832], ids=RIDS)
def test_rpr_emulated_kernel_matches_golden(oracle, emu, c):
    """McRprK's source + the device-side routing of commands that name a scaled reference, emulated on the CPU"""
    refs0, refs1, pocs, cmds, luts, prof, sc = _rpr_case(oracle, c)
    out = helpers_mc.emu_mc_rpr(emu, oracle, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, sc, prof=prof)
    assert [helpers.md5(p) for p in out] == c["md5"]


def test_rpr_bdof_dmvr_on_scaled_reference_rejected(oracle, emu):
    """the decoder switches BDOF / DMVR off when a reference is scaled: such a command is malformed"""
    refs0, refs1, pocs, cmds, luts, prof, sc = helpers_mc.rpr_case(oracle, 256, 128, 5)
    i = int(np.nonzero((cmds["tool"] == 0) & (cmds["inter_dir"] == 3) & (cmds["log2_w"] >= 4) & (cmds["log2_h"] >= 4))[0][0])
    cmds["log2_w"][i] = cmds["log2_h"][i] = 4; cmds["ref_idx0"][i] = 0
    for tool, flags in ((1, 1), (3, 3)):
        bad = cmds.copy()
        bad["tool"][i] = tool; bad["flags"][i] = flags; bad["bcw_idx_plus1"][i] = 0
        with pytest.raises(AssertionError):
            helpers_mc.emu_mc_rpr(emu, oracle, 256, 128, refs0, refs1, pocs, bad, luts, sc, prof=prof)
    ok = cmds.copy()
    ok["tool"][i] = 1; ok["flags"][i] = 1; ok["bcw_idx_plus1"][i] = 0; ok["ref_idx0"][i] = 1; ok["ref_idx1"][i] = 1    # both unscaled: fine
    helpers_mc.emu_mc_rpr(emu, oracle, 256, 128, refs0, refs1, pocs, ok, luts, sc, prof=prof)


@pytest.mark.gpu
@pytest.mark.parametrize("c", RPR_GOLDEN, ids=RIDS)
def test_rpr_gpu_matches_golden(oracle, gpu_ctx, c):
    refs0, refs1, pocs, cmds, luts, prof, sc = _rpr_case(oracle, c)
    out = helpers_mc.gpu_mc_rpr(gpu_ctx, c["w"], c["h"], refs0, refs1, pocs, cmds, luts, sc, prof=prof)
    assert [helpers.md5(p) for p in out] == c["md5"]


@pytest.mark.gpu
def test_rpr_gpu_matches_oracle_and_refuses_unscaled_size_mismatch(oracle, gpu_ctx):
    from openvvc_b200 import api
    w, h = 320, 192
    refs0, refs1, pocs, cmds, luts, prof, sc = helpers_mc.rpr_case(oracle, w, h, 77, 7, True, "noise", 0, 0)
    want = helpers_mc.blank(w, h)
    oracle.mc_picture_rpr(want, [refs0, refs1], pocs, cmds, sc, 7, impl="port", lmcs=luts, prof=prof)
    helpers.assert_planes_equal(helpers_mc.gpu_mc_rpr(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts, sc, prof=prof), want, "CUDA RPR vs oracle")
    # a reference of another size whose scale factor was not announced: refused, not read out of bounds
    sc.scale[0][0][0] = sc.scale[0][0][1] = 1 << 14
    with pytest.raises(api.OVGPUError):
        helpers_mc.gpu_mc_rpr(gpu_ctx, w, h, refs0, refs1, pocs, cmds, luts, sc, prof=prof)
