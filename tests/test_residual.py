"""Residual path (dequantisation, LFNST, DCT-2/DST-7/DCT-8, transform skip, residual add / joint CbCr / LMCS chroma
scale): oracle port vs golden / compiled reference (CPU), kernel-source emulation vs oracle (CPU), and the sm_100a
kernels through the C ABI vs oracle / golden (GPU).

Arithmetic of record = the reference's scalar C templates.  Its SSE/AVX2 variants saturate asymmetrically where the
C code clips symmetrically (dequantisation: -32768 vs -32767; LFNST: +32767 vs a wrapped +32768), so SIMD is only
compared on `legal` coefficient levels -- the only ones a conforming stream can carry."""
import json
import os

import numpy as np
import pytest

import helpers
import helpers_tr

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tr_golden.json")))
IDS = lambda c: "%dx%d_ict%d_tsd%d_dq%d%s" % (c["w"], c["h"], c["ict_type"], c["sh_ts_disabled"], c["dep_quant"], "_bdpcm" if c["p_bdpcm"] else "")
KW = lambda c: dict(dep_quant=c["dep_quant"], sh_ts_disabled=c["sh_ts_disabled"], ict_type=c["ict_type"])


@pytest.fixture(scope="module")
def pt(oracle):
    from oracle import pyoracle_tr
    return pyoracle_tr


def _case(c):
    planes, cmds, arena = helpers_tr.tr_case(c["w"], c["h"], c["seed"], c["log2_ctu"], c["ict_type"], c["sh_ts_disabled"],
                                             c["dep_quant"], c["legal"], **(dict(p_intra=0.6, p_bdpcm=c["p_bdpcm"]) if c["p_bdpcm"] else {}))
    assert helpers.md5(cmds, arena, *planes) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return planes, cmds, arena


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(pt, c):
    planes, cmds, arena = _case(c)
    pt.residual_picture(planes, cmds, arena, c["log2_ctu"], impl="port", **KW(c))
    assert [helpers.md5(p) for p in planes] == c["md5"]


@pytest.mark.parametrize("w,h,log2,ict,tsd,dq", [(416, 240, 7, 0, 0, 1), (416, 240, 7, 3, 1, 0), (832, 480, 7, 1, 0, 1),
                                                  (264, 136, 6, 2, 1, 1), (72, 40, 5, 1, 0, 0)])
def test_port_vs_compiled_reference(oracle, pt, w, h, log2, ict, tsd, dq):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, cmds, arena = helpers_tr.tr_case(w, h, 5 * w + h, log2, ict, tsd, dq, legal=False, p_intra=0.35)
    a = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    b = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="ref", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict, simd=0)
    helpers.assert_planes_equal(a, b, "residual port vs reference (scalar C)")
    assert sum(int((x != y).sum()) for x, y in zip(a, planes)) > 1000


@pytest.mark.parametrize("w,h,log2,ict,tsd,dq", [(416, 240, 7, 1, 0, 1), (832, 480, 7, 2, 0, 1), (256, 128, 6, 0, 1, 0)])
def test_port_vs_reference_simd_on_legal_levels(oracle, pt, w, h, log2, ict, tsd, dq):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, cmds, arena = helpers_tr.tr_case(w, h, 3 * w + h, log2, ict, tsd, dq, legal=True, p_intra=0.35)
    a = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    b = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="ref", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict, simd=1)
    helpers.assert_planes_equal(a, b, "residual port vs reference (SSE/AVX2 table)")


@pytest.mark.parametrize("w,h,log2,ict,tsd,dq,legal", [(416, 240, 7, 0, 0, 1, False), (416, 240, 7, 3, 1, 0, False),
                                                        (256, 128, 6, 1, 0, 1, True), (136, 72, 5, 2, 1, 1, False),
                                                        (832, 480, 7, 1, 0, 0, False)])
def test_emulated_kernel_matches_oracle(pt, emu, w, h, log2, ict, tsd, dq, legal):
    planes, cmds, arena = helpers_tr.tr_case(w, h, 7 * w + h, log2, ict, tsd, dq, legal, p_intra=0.35)
    want = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    got = helpers_tr.emu_residual(emu, planes, cmds, arena, dq, tsd, ict)
    helpers.assert_planes_equal(got, want, "emulated residual kernels vs oracle")


@pytest.mark.parametrize("w,h,log2,ict,tsd,dq", [(416, 240, 7, 0, 0, 1), (416, 240, 7, 1, 1, 1), (832, 480, 7, 3, 1, 0), (264, 136, 6, 2, 0, 1)])
def test_bdpcm_port_vs_reference_and_emulation(oracle, pt, emu, w, h, log2, ict, tsd, dq):
    """BDPCM coding units (rcn_bdpcm_tb, rcn_transform_tree.c:630-688): row / column accumulation of the transform-skip
    levels, luma and chroma, raster and sub-block coefficient order; port pinned to the scalar reference, kernel to the port."""
    planes, cmds, arena = helpers_tr.tr_case(w, h, 11 * w + h, log2, ict, tsd, dq, legal=True, p_intra=0.6, p_bdpcm=0.9)
    assert (cmds["bdpcm"] & 1).sum() > 2 and (cmds["bdpcm"] & 4).any() and (cmds["bdpcm"] & 2).any() and (cmds["bdpcm"] & 8).any()
    kw = dict(dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    want = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", **kw)
    if oracle.ref_available():
        ref = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="ref", simd=0, **kw)
        helpers.assert_planes_equal(want, ref, "BDPCM port vs reference (scalar C)")
    plain = cmds.copy()
    plain["bdpcm"] = 0
    assert any((a != b).any() for a, b in zip(want, pt.residual_picture([p.copy() for p in planes], plain, arena, log2, impl="port", **kw)))
    helpers.assert_planes_equal(helpers_tr.emu_residual(emu, planes, cmds, arena, dq, tsd, ict), want, "emulated BDPCM vs oracle")


def test_item_lists_cover_every_block_once_in_decreasing_size(emu):
    """host work lists (tr_host.h): every coded transform block exactly once, <= 256 samples in the warp list, larger ones in
    the CTA list, both ordered by decreasing size (the kernels rely on nothing else)"""
    import ctypes as C
    _, cmds, _ = helpers_tr.tr_case(832, 480, 4242, 7, 1, 0, 1, p_intra=0.3)
    items = np.zeros((3 * len(cmds) + 1, 2), np.uint32)
    ns, nl = C.c_int(), C.c_int()
    n = emu.emu_tu_items(cmds.ctypes.data_as(C.c_void_p), len(cmds), items.ctypes.data_as(C.c_void_p), C.byref(ns), C.byref(nl))
    want = set()
    for i, c in enumerate(cmds):
        m = int(c["cbf_mask"])
        kinds = ([0] if m & 0x10 else []) + ([3] if m & 0x08 else ([1] if m & 0x02 else []) + ([2] if m & 0x01 else []))
        want |= {(i, k) for k in kinds}
    got = [(int(a), int(b)) for a, b in items[:n]]
    assert len(got) == len(set(got)) == len(want) and set(got) == want and n == ns.value + nl.value
    size = [int(cmds["log2_w"][i]) + int(cmds["log2_h"][i]) - (2 if k else 0) for i, k in got]
    assert all(s <= 8 for s in size[:ns.value]) and all(s > 8 for s in size[ns.value:]) and ns.value and nl.value
    assert size[:ns.value] == sorted(size[:ns.value], reverse=True) and size[ns.value:] == sorted(size[ns.value:], reverse=True)


def test_zero_residual_is_identity(pt, emu):
    planes, cmds, arena = helpers_tr.tr_case(256, 128, 9, 7)
    arena[:] = 0
    got = helpers_tr.emu_residual(emu, planes, cmds, arena)
    helpers.assert_planes_equal(got, planes, "all-zero coefficients")


def test_emu_rejects_out_of_range_offsets(emu):
    planes, cmds, arena = helpers_tr.tr_case(256, 128, 10, 7)
    cmds = cmds.copy()
    i = int(np.nonzero(cmds["cbf_mask"] & 0x10)[0][0])
    cmds["coeff_off"][i][2] = len(arena)
    with pytest.raises(AssertionError):
        helpers_tr.emu_residual(emu, planes, cmds, arena)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_matches_golden(gpu_ctx, c):
    planes, cmds, arena = _case(c)
    out = helpers_tr.gpu_residual(gpu_ctx, planes, cmds, arena, **KW(c))
    assert [helpers.md5(p) for p in out] == c["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,ict,tsd,dq", [(416, 240, 7, 1, 0, 1), (264, 136, 6, 3, 1, 0), (1920, 1080, 7, 1, 0, 1)])
def test_gpu_matches_oracle(pt, gpu_ctx, w, h, log2, ict, tsd, dq):
    planes, cmds, arena = helpers_tr.tr_case(w, h, 9 * w + h, log2, ict, tsd, dq, p_intra=0.35)
    want = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    helpers.assert_planes_equal(helpers_tr.gpu_residual(gpu_ctx, planes, cmds, arena, dq, tsd, ict), want, "CUDA residual vs oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,ict,tsd,dq", [(416, 240, 7, 1, 1, 1), (1920, 1080, 7, 3, 0, 1)])
def test_gpu_bdpcm_matches_oracle(pt, gpu_ctx, w, h, log2, ict, tsd, dq):
    planes, cmds, arena = helpers_tr.tr_case(w, h, 13 * w + h, log2, ict, tsd, dq, legal=True, p_intra=0.6, p_bdpcm=0.9)
    assert (cmds["bdpcm"] != 0).sum() > 5
    want = pt.residual_picture([p.copy() for p in planes], cmds, arena, log2, impl="port", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict)
    helpers.assert_planes_equal(helpers_tr.gpu_residual(gpu_ctx, planes, cmds, arena, dq, tsd, ict), want, "CUDA BDPCM vs oracle")


@pytest.mark.gpu
def test_gpu_full_size_4k_vs_reference(oracle, pt, gpu_ctx):
    """BASELINE.json config 2 geometry (3840x2160): CUDA vs the compiled reference (scalar C) when present."""
    w, h = 3840, 2160
    planes, cmds, arena = helpers_tr.tr_case(w, h, 78, 7, 1, 0, 1, legal=True, p_intra=0.1, min_log2=3)
    want = pt.residual_picture([p.copy() for p in planes], cmds, arena, 7, impl="ref" if oracle.ref_available() else "port",
                               dep_quant=1, sh_ts_disabled=0, ict_type=1, simd=0)
    helpers.assert_planes_equal(helpers_tr.gpu_residual(gpu_ctx, planes, cmds, arena, 1, 0, 1), want, "CUDA residual vs reference at 4K")


@pytest.mark.gpu
def test_gpu_rejects_malformed_commands(gpu_ctx):
    from openvvc_b200 import api
    planes, cmds, arena = helpers_tr.tr_case(256, 128, 10, 7)
    cmds = cmds.copy()
    cmds["log2_w"][0] = 9
    with pytest.raises(api.OVGPUError):
        helpers_tr.gpu_residual(gpu_ctx, planes, cmds, arena)
