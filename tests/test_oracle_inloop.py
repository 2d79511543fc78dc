"""CPU tests (no GPU): pin the oracle port (oracle/port_alf.c, port_sao.c) against
 (1) the golden fixtures generated from the unmodified reference (tests/golden/gen_golden.py),
 (2) the compiled reference itself when oracle/_ref is present (always in the build container)."""
import json
import os

import numpy as np
import pytest

import helpers

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "inloop_golden.json")))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_ctu%d_%s" % (c["w"], c["h"], 1 << c["log2_ctu"], c["kind"]))
def test_port_matches_golden_md5(oracle, case):
    po = oracle
    planes, tab, ctu, _ = helpers.alf_case(po, case["w"], case["h"], case["seed"], case["log2_ctu"], case["kind"])
    # the fixture is only meaningful if the seeded generator reproduced the same inputs
    assert helpers.md5(*planes) == case["input_md5"], "synthetic generator drifted; regenerate golden"
    alf = po.alf_picture(planes, tab, ctu, case["log2_ctu"], impl="port")
    assert [helpers.md5(p) for p in alf] == case["alf_md5"]
    splanes, sao = helpers.sao_case(po, case["w"], case["h"], case["seed"], case["log2_ctu"], case["kind"])
    assert helpers.md5(sao) == case["sao_param_md5"]
    so = po.sao_picture(splanes, sao, case["log2_ctu"], impl="port")
    assert [helpers.md5(p) for p in so] == case["sao_md5"]


def _tiny():
    z = np.load(os.path.join(GOLDEN, "inloop_tiny.npz"))
    return z


def test_port_matches_golden_tiny_arrays(oracle):
    """Generator-independent: inputs and reference outputs stored in full."""
    po = oracle
    z = _tiny()
    planes = [np.ascontiguousarray(z[k]) for k in ("y", "cb", "cr")]
    tab = po.AlfTables()
    for k in ("luma_coeff", "luma_clip", "chroma_coeff", "chroma_clip", "cc_coeff"):
        tab.np(k)[:] = z[k]
    rows = z["ctu_rows"]
    ctu = dict(flag=rows[:, 0], idx=rows[:, 1], cb_alt=rows[:, 2], cr_alt=rows[:, 3], cc_cb=z["cc_cb"], cc_cr=z["cc_cr"],
               cc_cb_enabled=1, cc_cr_enabled=1)
    alf = po.alf_picture(planes, tab, ctu, int(z["log2_ctu"]), impl="port")
    helpers.assert_planes_equal(alf, [z["alf_y"], z["alf_cb"], z["alf_cr"]], "ALF tiny golden")
    so = po.sao_picture(planes, z["sao_params"].astype(po.SaoCtu), int(z["log2_ctu"]), impl="port")
    helpers.assert_planes_equal(so, [z["sao_y"], z["sao_cb"], z["sao_cr"]], "SAO tiny golden")


def test_alf_tables_match_reference(oracle):
    """rcn_alf_reconstruct_coeff_APS (rcn_alf.c:264): fixed sets, APS sets, transposes, clip LUT."""
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from openvvc_b200 import synth
    for seed in range(5):
        ap = synth.alf_params(256, 128, 300 + seed)
        a = po.alf_build_tables(ap["luma_aps"], ap["chroma_aps"], impl="port")
        b = po.alf_build_tables(ap["luma_aps"], ap["chroma_aps"], impl="ref")
        n = 16 + len(ap["luma_aps"])
        assert (a.np("luma_coeff")[:n] == b.np("luma_coeff")[:n]).all()
        assert (a.np("luma_clip")[:n] == b.np("luma_clip")[:n]).all()
        na = ap["chroma_aps"]["num_chroma_alts"]
        assert (a.np("chroma_coeff")[:na] == b.np("chroma_coeff")[:na]).all()
        assert (a.np("chroma_clip")[:na] == b.np("chroma_clip")[:na]).all()


REF_CASES = [(416, 240, 7, "natural"), (416, 240, 7, "noise"), (1280, 720, 7, "natural"), (832, 480, 6, "natural"),
             (264, 136, 7, "natural"), (200, 120, 5, "natural"), (64, 64, 7, "natural"), (128, 64, 7, "noise"),
             (136, 264, 7, "natural")]


@pytest.mark.parametrize("w,h,log2,kind", REF_CASES)
def test_port_vs_compiled_reference(oracle, w, h, log2, kind):
    """Differential test against the reference's scalar C path (arithmetic of record) -- and, for CTU 128
    CTC-like geometries, its SSE/AVX2 table as well."""
    po = oracle
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    seed = w * 7 + h
    planes, tab, ctu, _ = helpers.alf_case(po, w, h, seed, log2, kind)
    port = po.alf_picture(planes, tab, ctu, log2, impl="port")
    helpers.assert_planes_equal(port, po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=0), "ALF port vs ref C")
    splanes, sao = helpers.sao_case(po, w, h, seed, log2, kind)
    sport = po.sao_picture(splanes, sao, log2, impl="port")
    helpers.assert_planes_equal(sport, po.sao_picture(splanes, sao, log2, impl="ref", simd=0), "SAO port vs ref C")
    if log2 == 7 and w % 16 == 0 and h > 128:
        helpers.assert_planes_equal(port, po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=1), "ALF port vs ref SIMD")
        helpers.assert_planes_equal(sport, po.sao_picture(splanes, sao, log2, impl="ref", simd=1), "SAO port vs ref SIMD")


def test_alf_disabled_is_identity(oracle):
    po = oracle
    planes, tab, ctu, _ = helpers.alf_case(po, 256, 128, 5)
    ctu = dict(ctu); ctu["flag"] = np.zeros_like(ctu["flag"]); ctu["cc_cb_enabled"] = 0; ctu["cc_cr_enabled"] = 0
    out = po.alf_picture(planes, tab, ctu, 7, impl="port")
    helpers.assert_planes_equal(out, planes, "ALF off")


def test_lmcs_reshape_backward_ref_slot_matches_port(oracle):
    """the reference's own lmcs_reshape_backward slot, CTU by CTU (what bench.py's reference arm times), against the port's plane pass"""
    from openvvc_b200 import synth
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    luts = oracle.lmcs_build_luts(synth.lmcs_data(41), impl="ref")
    a = synth.planes(416, 240, 9)[0]
    b = a.copy()
    oracle.lmcs_reshape_backward(a, luts, 7, impl="ref")
    oracle.lmcs_reshape_backward(b, luts, 7, impl="port")
    assert (a == b).all() and not (a == synth.planes(416, 240, 9)[0]).all()
