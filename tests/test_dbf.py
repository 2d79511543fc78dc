"""Deblocking filter: map generator vs the reference's DBFInfo bookkeeping, oracle port vs golden / compiled reference
(CPU), kernel-source emulation vs oracle (CPU), and the sm_100a kernels through the C ABI vs oracle / golden (GPU)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers
import helpers_df

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "df_golden.json")))
IDS = lambda c: "%dx%d_ctu%d_%s" % (c["w"], c["h"], 1 << c["log2_ctu"], c["kind"])


@pytest.fixture(scope="module")
def pd(oracle):
    from oracle import pyoracle_df
    return pyoracle_df


def _case(c):
    planes, blks, maps = helpers_df.dbf_case(c["w"], c["h"], c["seed"], c["log2_ctu"], c["kind"], c["beta_offset"], c["tc_offset"])
    assert helpers.md5(blks, maps["ver"], maps["hor"], maps["qp"], *planes) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    return planes, blks, maps


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(pd, c):
    planes, blks, maps = _case(c)
    pd.dbf_picture_port(planes, maps)
    assert [helpers.md5(p) for p in planes] == c["md5"]


@pytest.mark.parametrize("w,h,log2", [(256, 128, 7), (416, 240, 7), (264, 136, 6), (200, 120, 5), (832, 480, 7)])
def test_picture_level_maps_equal_reference_bookkeeping(oracle, pd, w, h, log2):
    """The maps our generator builds == the reference's per-CTU DBFInfo (filled by its own dbf_utils.h helpers and
    carried by dbf_store_info / dbf_load_info) converted the way the integration shim converts it."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, blks, maps = helpers_df.dbf_case(w, h, 11 * w + h, log2)
    rec = pd.dbf_picture_ref(planes, blks, log2, record=True)
    for k in ("ver", "hor", "qp"):
        assert (rec[k] == maps[k]).all(), k


@pytest.mark.parametrize("w,h,log2,kind,bo,to", [(416, 240, 7, "blocky", 0, 0), (416, 240, 7, "natural", 0, 0), (832, 480, 7, "blocky", 2, 2),
                                                  (264, 136, 6, "blocky", -2, 0), (72, 40, 5, "blocky", 0, -2), (832, 480, 7, "noise", 0, 0)])
def test_port_vs_compiled_reference(oracle, pd, w, h, log2, kind, bo, to):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, blks, maps = helpers_df.dbf_case(w, h, 5 * w + h, log2, kind, bo, to)
    a = pd.dbf_picture_port([p.copy() for p in planes], maps)
    for simd in (0, 1):
        b = [p.copy() for p in planes]
        pd.dbf_picture_ref(b, blks, log2, bo, to, simd=simd)
        helpers.assert_planes_equal(a, b, "DBF port vs reference (simd=%d)" % simd)
    if kind != "noise":
        assert sum(int((x != y).sum()) for x, y in zip(a, planes)) > w * h // 20


def test_blocky_input_reaches_every_filter(oracle, pd):
    """coverage: long (several P/Q length pairs), strong-short, weak, chroma strong and chroma weak filters all run"""
    st = (C.c_long * 32)()
    oracle.port().ovo_dbf_stats(st, 1)
    planes, blks, maps = helpers_df.dbf_case(832, 480, 31, 7, "blocky")
    pd.dbf_picture_port(planes, maps)
    oracle.port().ovo_dbf_stats(st, 1)
    s = list(st)
    assert all(v > 0 for v in s[:5]), s
    assert sum(1 for v in s[5:] if v > 0) >= 5, s


@pytest.mark.parametrize("w,h,log2,kind,bo,to", [(416, 240, 7, "blocky", 0, 0), (416, 240, 7, "natural", 2, -2), (264, 136, 6, "blocky", 0, 0),
                                                  (136, 72, 5, "blocky", -2, 2), (832, 480, 7, "blocky", 0, 0)])
def test_emulated_kernel_matches_oracle(pd, emu, w, h, log2, kind, bo, to):
    planes, blks, maps = helpers_df.dbf_case(w, h, 7 * w + h, log2, kind, bo, to)
    want = pd.dbf_picture_port([p.copy() for p in planes], maps)
    helpers.assert_planes_equal(helpers_df.emu_dbf(emu, planes, maps), want, "emulated DBF kernels vs oracle")


def test_no_edges_is_identity(emu):
    planes, blks, maps = helpers_df.dbf_case(256, 128, 9, 7)
    maps["ver"][:] = 0; maps["hor"][:] = 0
    helpers.assert_planes_equal(helpers_df.emu_dbf(emu, planes, maps), planes, "no edges")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_matches_golden(gpu_ctx, c):
    planes, blks, maps = _case(c)
    out = helpers_df.gpu_dbf(gpu_ctx, planes, maps, c["log2_ctu"])
    assert [helpers.md5(p) for p in out] == c["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,kind", [(416, 240, 7, "blocky"), (264, 136, 6, "natural"), (1920, 1080, 7, "blocky")])
def test_gpu_matches_oracle(pd, gpu_ctx, w, h, log2, kind):
    planes, blks, maps = helpers_df.dbf_case(w, h, 9 * w + h, log2, kind)
    want = pd.dbf_picture_port([p.copy() for p in planes], maps)
    helpers.assert_planes_equal(helpers_df.gpu_dbf(gpu_ctx, planes, maps, log2), want, "CUDA DBF vs oracle")


@pytest.mark.gpu
def test_gpu_full_size_4k_vs_reference(oracle, pd, gpu_ctx):
    """BASELINE.json config 3 geometry (3840x2160): CUDA (parallel, in place) vs the compiled reference (CTU by CTU)."""
    w, h = 3840, 2160
    planes, blks, maps = helpers_df.dbf_case(w, h, 79, 7, "blocky", min_log2=3)
    want = [p.copy() for p in planes]
    if oracle.ref_available():
        pd.dbf_picture_ref(want, blks, 7, simd=1)
    else:
        pd.dbf_picture_port(want, maps)
    helpers.assert_planes_equal(helpers_df.gpu_dbf(gpu_ctx, planes, maps, 7), want, "CUDA DBF vs reference at 4K")


@pytest.mark.gpu
def test_gpu_dbf_is_deterministic(gpu_ctx):
    """the in-place parallel passes have no data races: repeated runs give identical pictures"""
    planes, blks, maps = helpers_df.dbf_case(1920, 1080, 33, 7, "blocky")
    first = helpers_df.gpu_dbf(gpu_ctx, planes, maps, 7)
    for _ in range(3):
        helpers.assert_planes_equal(helpers_df.gpu_dbf(gpu_ctx, planes, maps, 7), first, "repeat")
