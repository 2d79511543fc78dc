"""Luma and chroma intra prediction (reference fill, [1 2 1] filter, planar / DC / angular with PDPC, wide angles): oracle port vs
golden / compiled reference (CPU), availability replay vs the reference's progress bit-field, kernel-source emulation vs
oracle incl. the dependency-level schedule (CPU), sm_100a kernel through the C ABI vs oracle / golden (GPU)."""
import json
import os

import numpy as np
import pytest

import helpers
import helpers_intra

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "intra_golden.json")))
IDS = lambda c: "%dx%d_ctu%d%s" % (c["w"], c["h"], 1 << c["log2_ctu"], ("", "_yuv", "_yuv_ciip", "_yuv_ciip_mrl", "_yuv_cclm", "_yuv_cclm_collocated", "_yuv_mip")[c["chroma"]])


@pytest.fixture(scope="module")
def pi(oracle):
    from oracle import pyoracle_intra
    return pyoracle_intra


def _case(c):
    planes, cmds = helpers_intra.intra_case(c["w"], c["h"], c["seed"], c["log2_ctu"], chroma=bool(c["chroma"]),
                                            p_ciip=0.4 if c["chroma"] in (2, 3) else 0.0, p_mrl=0.4 if c["chroma"] == 3 else 0.0,
                                            p_cclm=0.4 if c["chroma"] in (4, 5) else 0.0, collocated=c["chroma"] == 5,
                                            p_mip=0.4 if c["chroma"] == 6 else 0.0)
    assert helpers.md5(cmds, planes[0]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return planes, cmds


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(pi, c):
    planes, cmds = _case(c)
    # our replay of the availability words == what the reference's bit-field held (golden from the reference run)
    assert helpers.md5(cmds["avail_cols"], cmds["avail_rows"]) == c["avail_md5"]
    pi.intra_picture(planes, cmds, c["log2_ctu"], impl="port")
    assert (helpers.md5(*planes) if c["chroma"] else helpers.md5(planes[0])) == c["md5"]


@pytest.mark.parametrize("chroma", [False, True], ids=["y", "yuv"])
@pytest.mark.parametrize("w,h,log2", [(128, 128, 7), (416, 240, 7), (832, 480, 7), (264, 136, 6), (72, 40, 5), (200, 120, 5)])
def test_port_vs_compiled_reference(oracle, pi, w, h, log2, chroma):
    """every mode class over random partitions incl. partial CTUs; scalar C and the SSE/AVX2 table"""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, cmds = helpers_intra.intra_case(w, h, 5 * w + h, log2, chroma=chroma)
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    for simd in (0, 1):
        rc = cmds.copy()
        rc["avail_cols"] = 0; rc["avail_rows"] = 0
        got = pi.intra_picture([p.copy() for p in planes], rc, log2, impl="ref", simd=simd)
        assert (rc["avail_cols"] == cmds["avail_cols"]).all() and (rc["avail_rows"] == cmds["avail_rows"]).all()
        for k in range(3):
            assert (got[k] == want[k]).all(), "intra port vs reference (simd=%d, plane %d): %d samples differ" % (simd, k, (got[k] != want[k]).sum())
    for k in range(3 if chroma else 1):
        assert 0.4 < (want[k] != planes[k]).mean() < 0.9         # ~30 % of the CUs are non-intra placeholders
        assert len(np.unique(want[k])) > 200                    # and the predictions carry real content
    if not chroma:
        assert (want[1] == planes[1]).all() and (want[2] == planes[2]).all()


def test_every_mode_and_shape_is_covered():
    _, cmds = helpers_intra.intra_case(832, 480, 4645, 7)
    assert set(cmds["mode"]) >= set(range(0, 67))
    shapes = set(zip(cmds["log2_w"], cmds["log2_h"]))
    assert len(shapes) >= 20 and any(int(a) - int(b) >= 2 for a, b in shapes) and any(int(b) - int(a) >= 2 for a, b in shapes)      # wide angles both ways


@pytest.mark.parametrize("chroma", [False, True], ids=["y", "yuv"])
@pytest.mark.parametrize("w,h,log2", [(128, 128, 7), (416, 240, 7), (264, 136, 6), (200, 120, 5), (832, 480, 7)])
def test_emulated_kernel_matches_oracle(pi, emu, w, h, log2, chroma):
    planes, cmds = helpers_intra.intra_case(w, h, 7 * w + h, log2, chroma=chroma)
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    for rev in (0, 1):          # CTAs of a level may run in any order
        emu.emu_set_reverse_block_order(rev)
        try:
            got, nl = helpers_intra.emu_intra(emu, planes, cmds, log2)
        finally:
            emu.emu_set_reverse_block_order(0)
        assert all((got[k] == want[k]).all() for k in range(3)) and 1 < nl <= 2 * len(cmds)          # luma levels + chroma levels


def test_chroma_only_and_luma_only_commands(pi, emu):
    """dual-tree style commands: luma-only, chroma-only and both in one list"""
    planes, cmds = helpers_intra.intra_case(416, 240, 77, 7, chroma=True, p_inter=0.0)
    cmds = cmds.copy()
    cmds["flags"][::3] = 1          # luma only
    cmds["flags"][1::3] = 2         # chroma only
    want = pi.intra_picture([p.copy() for p in planes], cmds, 7, impl="port")
    got, _ = helpers_intra.emu_intra(emu, planes, cmds, 7)
    assert all((got[k] == want[k]).all() for k in range(3))
    bad = cmds.copy(); bad["flags"][0] = 0                                         # neither
    with pytest.raises(AssertionError):
        helpers_intra.emu_intra(emu, planes, bad, 7, keep_placeholders=True)
    bad = cmds.copy(); bad["flags"][0] = 2; bad["mode_c"][0] = 70                  # beyond the MDLM modes
    with pytest.raises(AssertionError):
        helpers_intra.emu_intra(emu, planes, bad, 7)


@pytest.mark.parametrize("w,h,log2", [(416, 240, 7), (264, 136, 6)])
def test_ciip_port_vs_reference_and_emu(oracle, pi, emu, w, h, log2):
    """CIIP (a14): planar intra blended over the inter prediction in place, weights 1..3, narrow-chroma rule (4xN CUs)"""
    planes, cmds = helpers_intra.intra_case(w, h, 31 * w + h, log2, chroma=True, p_ciip=0.5)
    ciip = (cmds["flags"] & 4) != 0
    assert ciip.sum() > 10 and (ciip & (cmds["log2_w"] == 2)).any() and set(cmds["ciip_wt"][ciip]) == {1, 2, 3}
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    if oracle.ref_available():
        for simd in (0, 1):
            got = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="ref", simd=simd)
            assert all((got[k] == want[k]).all() for k in range(3)), "CIIP port vs reference (simd=%d)" % simd
    got, _ = helpers_intra.emu_intra(emu, planes, cmds, log2)
    assert all((got[k] == want[k]).all() for k in range(3))


@pytest.mark.parametrize("w,h,log2", [(416, 240, 7), (264, 136, 6), (200, 120, 5)])
def test_mrl_port_vs_reference_and_emu(oracle, pi, emu, w, h, log2):
    """multi reference line (a18): lines 1 and 2, every mode class, blocks on the picture's left edge (no left line)"""
    planes, cmds = helpers_intra.intra_case(w, h, 17 * w + h, log2, chroma=True, p_mrl=0.5)
    mrl = cmds["mrl_idx"] > 0
    assert mrl.sum() > 20 and set(cmds["mrl_idx"]) == {0, 1, 2} and ((mrl & (cmds["x0"] == 0)).any() or w == 264)
    assert ((cmds["y0"][mrl] & ((1 << log2) - 1)) > 0).all()
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    if oracle.ref_available():
        for simd in (0, 1):
            got = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="ref", simd=simd)
            assert all((got[k] == want[k]).all() for k in range(3)), "MRL port vs reference (simd=%d)" % simd
    got, _ = helpers_intra.emu_intra(emu, planes, cmds, log2)
    assert all((got[k] == want[k]).all() for k in range(3))


@pytest.mark.parametrize("collocated", [False, True], ids=["6tap", "collocated"])
@pytest.mark.parametrize("w,h,log2", [(416, 240, 7), (264, 136, 6), (200, 120, 5)])
def test_cclm_port_vs_reference_and_emu(oracle, pi, emu, w, h, log2, collocated):
    """CCLM / MDLM (a20): neighbour selection from the availability words, min / max pairing, the division LUT, both luma
    down-sampling variants, CTU-first-line rule; the luma block of the same command is predicted first"""
    planes, cmds = helpers_intra.intra_case(w, h, 13 * w + h, log2, chroma=True, p_cclm=0.5, collocated=collocated)
    assert {67, 68, 69} <= set(cmds["mode_c"])
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    if oracle.ref_available():
        for simd in (0, 1):
            got = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="ref", simd=simd)
            assert all((got[k] == want[k]).all() for k in range(3)), "CCLM port vs reference (simd=%d)" % simd
    for rev in (0, 1):
        emu.emu_set_reverse_block_order(rev)
        try:
            got, _ = helpers_intra.emu_intra(emu, planes, cmds, log2)
        finally:
            emu.emu_set_reverse_block_order(0)
        assert all((got[k] == want[k]).all() for k in range(3))


@pytest.mark.parametrize("w,h,log2", [(416, 240, 7), (832, 480, 7), (264, 136, 6), (200, 120, 5)])
def test_mip_port_vs_reference_and_emu(oracle, pi, emu, w, h, log2):
    """MIP (a21): the three matrix classes, plain and transposed, every up-sampling ratio incl. none (4x4, 8x8) and one-sided"""
    planes, cmds = helpers_intra.intra_case(w, h, 11 * w + h, log2, chroma=True, p_mip=0.5)
    mip = (cmds["flags"] & 16) != 0
    shapes = set(zip(cmds["log2_w"][mip], cmds["log2_h"][mip]))
    assert mip.sum() > 30 and (2, 2) in shapes and len(shapes) >= 12 and set(cmds["mode"][mip] >> 7) == {0, 1}
    assert (3, 3) in shapes or w == 416
    want = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    if oracle.ref_available():
        for simd in (0, 1):
            got = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="ref", simd=simd)
            assert all((got[k] == want[k]).all() for k in range(3)), "MIP port vs reference (simd=%d)" % simd
    for rev in (0, 1):
        emu.emu_set_reverse_block_order(rev)
        try:
            got, _ = helpers_intra.emu_intra(emu, planes, cmds, log2)
        finally:
            emu.emu_set_reverse_block_order(0)
        assert all((got[k] == want[k]).all() for k in range(3))
    bad = np.ascontiguousarray(cmds[mip][:1]).copy()
    bad["mode"] = 0x80 | 16                                                       # matrix index beyond the class
    with pytest.raises(AssertionError):
        helpers_intra.emu_intra(emu, planes, bad, log2)


def test_emu_rejects_blocks_outside_the_picture(emu):
    planes, cmds = helpers_intra.intra_case(128, 128, 3, 7, p_inter=0.0)
    cmds = cmds.copy()
    cmds["x0"][5] = 126
    with pytest.raises(AssertionError):
        helpers_intra.emu_intra(emu, planes, cmds, 7)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_matches_golden(gpu_ctx, c):
    planes, cmds = _case(c)
    out, nl = helpers_intra.gpu_intra(gpu_ctx, planes, cmds, c["log2_ctu"])
    if c["chroma"]:
        assert helpers.md5(*out) == c["md5"]
    else:
        assert helpers.md5(out[0]) == c["md5"]
        assert (out[1] == planes[1]).all() and (out[2] == planes[2]).all()


@pytest.mark.gpu
def test_gpu_full_size_4k_vs_reference(oracle, pi, gpu_ctx):
    """an all-intra 3840x2160 picture (the I picture of a RA GOP): CUDA level schedule vs the compiled reference"""
    w, h = 3840, 2160
    planes, cmds = helpers_intra.intra_case(w, h, 85, 7, min_log2=3, chroma=True)
    want = pi.intra_picture([p.copy() for p in planes], cmds.copy(), 7, impl="ref" if oracle.ref_available() else "port")
    out, nl = helpers_intra.gpu_intra(gpu_ctx, planes, cmds, 7)
    for k in range(3):
        assert (out[k] == want[k]).all()
