"""GPU parity tests (run with -m gpu on a B200): the sm_100a kernels, called through the C ABI
(libovvc_b200.so), against the oracle on the same seeded inputs, against the golden fixtures generated from
the unmodified reference, against the prebuilt reference itself (oracle/_ref travels as a .so), and at
BASELINE.json's full picture sizes."""
import json
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "inloop_golden.json")))


def gpu_alf(ctx, planes, tab, ctu, log2):
    from openvvc_b200 import api
    return ctx.alf_frame_host(planes, helpers.gpu_tables(tab), api.alf_ctu_array(ctu), (ctu["cc_cb"], ctu["cc_cr"]),
                              (ctu["cc_cb_enabled"], ctu["cc_cr_enabled"]), log2)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_ctu%d_%s" % (c["w"], c["h"], 1 << c["log2_ctu"], c["kind"]))
def test_gpu_matches_golden_md5(oracle, gpu_ctx, case):
    planes, tab, ctu, _ = helpers.alf_case(oracle, case["w"], case["h"], case["seed"], case["log2_ctu"], case["kind"])
    assert helpers.md5(*planes) == case["input_md5"]
    alf = gpu_alf(gpu_ctx, planes, tab, ctu, case["log2_ctu"])
    assert [helpers.md5(p) for p in alf] == case["alf_md5"]
    splanes, sao = helpers.sao_case(oracle, case["w"], case["h"], case["seed"], case["log2_ctu"], case["kind"])
    so = gpu_ctx.sao_frame_host(splanes, sao, case["log2_ctu"])
    assert [helpers.md5(p) for p in so] == case["sao_md5"]


GPU_CASES = [(416, 240, 7, "natural", False), (416, 240, 7, "noise", False), (264, 136, 7, "natural", False),
             (64, 64, 7, "natural", False), (128, 64, 7, "noise", False), (256, 128, 6, "natural", False),
             (200, 120, 5, "natural", False), (416, 240, 7, "natural", True), (1920, 1080, 7, "natural", False),
             (136, 264, 7, "noise", False)]


@pytest.mark.parametrize("w,h,log2,kind,wide", GPU_CASES)
def test_gpu_alf_matches_oracle(oracle, gpu_ctx, w, h, log2, kind, wide):
    planes, tab, ctu, _ = helpers.alf_case(oracle, w, h, 17 * w + h, log2, kind, wide=wide)
    ref = oracle.alf_picture(planes, tab, ctu, log2, impl="port")
    helpers.assert_planes_equal(gpu_alf(gpu_ctx, planes, tab, ctu, log2), ref, "CUDA ALF vs oracle")


@pytest.mark.parametrize("w,h,log2,kind,wide", GPU_CASES)
def test_gpu_sao_matches_oracle(oracle, gpu_ctx, w, h, log2, kind, wide):
    planes, sao = helpers.sao_case(oracle, w, h, 19 * w + h, log2, kind)
    ref = oracle.sao_picture(planes, sao, log2, impl="port")
    helpers.assert_planes_equal(gpu_ctx.sao_frame_host(planes, sao, log2), ref, "CUDA SAO vs oracle")


@pytest.mark.parametrize("w,h", [(3840, 2160), (7680, 4320)])
def test_gpu_full_size_vs_reference(oracle, gpu_ctx, w, h):
    """BASELINE.json sizes (4K, 8K, CTU 128): CUDA vs the compiled reference's SIMD path (fast enough at
    full size) when oracle/_ref is present, else vs the port; plus the size-independent property that a
    picture with every CTU flag off comes back untouched."""
    planes, tab, ctu, _ = helpers.alf_case(oracle, w, h, 23, 7, "natural")
    impl = dict(impl="ref", simd=1) if oracle.ref_available() else dict(impl="port")
    ref = oracle.alf_picture(planes, tab, ctu, 7, **impl)
    helpers.assert_planes_equal(gpu_alf(gpu_ctx, planes, tab, ctu, 7), ref, "CUDA ALF vs reference at %dx%d" % (w, h))
    splanes, sao = helpers.sao_case(oracle, w, h, 29, 7, "natural")
    ref = oracle.sao_picture(splanes, sao, 7, **impl)
    helpers.assert_planes_equal(gpu_ctx.sao_frame_host(splanes, sao, 7), ref, "CUDA SAO vs reference at %dx%d" % (w, h))
    off = dict(ctu); off["flag"] = np.zeros_like(ctu["flag"]); off["cc_cb_enabled"] = 0; off["cc_cr_enabled"] = 0
    helpers.assert_planes_equal(gpu_alf(gpu_ctx, planes, tab, off, 7), planes, "ALF off must be identity")


def test_gpu_device_resident_path_and_profile(oracle, gpu_ctx):
    """Device-resident API (what bench.py's `value` times): upload once, run twice, download."""
    from openvvc_b200 import api
    w, h = 1920, 1080
    planes, tab, ctu, _ = helpers.alf_case(oracle, w, h, 31, 7, "natural")
    ref = oracle.alf_picture(planes, tab, ctu, 7, impl="port")
    src, dst = gpu_ctx.picture(w, h), gpu_ctx.picture(w, h)
    src.upload(planes)
    ap = api.AlfParams(gpu_ctx, w, h, 7)
    ap.upload(helpers.gpu_tables(tab), api.alf_ctu_array(ctu), (ctu["cc_cb"], ctu["cc_cr"]))
    n0 = gpu_ctx.launch_count()
    gpu_ctx.profile(True); gpu_ctx.profile_reset()
    ap.run(dst, src); ap.run(dst, src)
    ms, n = gpu_ctx.profile_get("alf_luma")
    gpu_ctx.profile(False)
    assert n == 2 and ms > 0
    assert gpu_ctx.launch_count() - n0 == 4
    helpers.assert_planes_equal(dst.download(), ref, "device-resident ALF")
    ap.destroy(); src.destroy(); dst.destroy()
