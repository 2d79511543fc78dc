"""Intra sub-partitions (recon_isp_subtree_{v,h}, rcn_transform_tree.c:1087-1205; vvc_intra_pred_isp, rcn_intra.c:567-643; the
1xN / 2xN / Nx1 / Nx2 transform blocks, :919-1066): oracle port vs golden / compiled reference, availability replay."""
import collections
import json
import os

import numpy as np
import pytest

import helpers
import helpers_intra

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "isp_golden.json")))
IDS = lambda c: "%dx%d_ctu%d" % (c["w"], c["h"], 1 << c["log2_ctu"])


@pytest.fixture(scope="module")
def pi(oracle):
    from oracle import pyoracle_intra
    return pyoracle_intra


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(pi, c):
    planes, cmds, words, tus, arena = helpers_intra.isp_case(c["w"], c["h"], c["seed"], c["log2_ctu"])
    assert helpers.md5(cmds, tus, arena, planes[0]) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    # the replayed availability words == what the reference's progress field held
    assert helpers.md5(cmds["avail_cols"], cmds["avail_rows"], words["common"], words["part"]) == c["avail_md5"]
    pi.intra_picture_isp(planes, cmds, pi.isp_info_array(words, tus), arena, c["log2_ctu"], impl="port")
    assert helpers.md5(*planes) == c["md5"]


@pytest.mark.parametrize("w,h,log2", [(416, 240, 7), (832, 480, 7), (264, 136, 6), (200, 120, 5)])
def test_port_vs_compiled_reference(oracle, pi, w, h, log2):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, cmds, words, tus, arena = helpers_intra.isp_case(w, h, 3 * w + h, log2)
    t = tus[tus["cbf_mask"] != 0]
    shapes = collections.Counter(zip(t["log2_w"].tolist(), t["log2_h"].tolist()))
    if w >= 416:      # every thin shape, LFNST, DST-VII both ways
        assert any(a == 0 for a, _ in shapes) and any(b == 0 for _, b in shapes) and any(a == 1 for a, _ in shapes) and any(b == 1 for _, b in shapes)
        assert ((t["flags"] & 2) != 0).any() and set(t["cu_mts_idx"].tolist()) >= {1, 2, 3}
    isp = pi.isp_info_array(words, tus)
    want = pi.intra_picture_isp([p.copy() for p in planes], cmds, isp, arena, log2, impl="port")
    for simd in (0, 1):
        rc, ri = cmds.copy(), isp.copy()
        rc["avail_cols"] = 0; rc["avail_rows"] = 0; ri["common"] = 0; ri["part"] = 0
        got = pi.intra_picture_isp([p.copy() for p in planes], rc, ri, arena, log2, impl="ref", simd=simd)
        assert (ri["common"] == isp["common"]).all() and (ri["part"] == isp["part"]).all() and (rc["avail_cols"] == cmds["avail_cols"]).all()
        for k in range(3):
            assert (got[k] == want[k]).all(), "ISP port vs reference (simd=%d, plane %d): %d samples differ" % (simd, k, (got[k] != want[k]).sum())
    # the partitions' residual is part of the result
    off = isp.copy(); off["tu"]["cbf_mask"] = 0
    assert (pi.intra_picture_isp([p.copy() for p in planes], cmds, off, arena, log2, impl="port")[0] != want[0]).any()
