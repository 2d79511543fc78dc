"""CPU tests (no GPU): the product's kernel SOURCE (openvvc_b200/csrc/*_kernels.h) executed through the
serial CTA emulation of csrc/dev.h must match the oracle bit for bit.  This checks tiling, halo staging,
virtual-boundary logic and the packed-int16 arithmetic before any GPU time is spent; the GPU tests
(test_gpu_inloop.py) then check the real sm_100a build."""
import pytest

import helpers

CASES = [(416, 240, 7, "natural", False), (416, 240, 7, "noise", False), (264, 136, 7, "natural", False),
         (64, 64, 7, "natural", False), (128, 64, 7, "noise", False), (256, 128, 6, "natural", False),
         (200, 120, 5, "natural", False), (416, 240, 7, "natural", True), (1280, 720, 7, "natural", False)]


@pytest.mark.parametrize("w,h,log2,kind,wide", CASES)
def test_emulated_alf_matches_oracle(oracle, emu, w, h, log2, kind, wide):
    planes, tab, ctu, _ = helpers.alf_case(oracle, w, h, 11 * w + h, log2, kind, wide=wide)
    ref = oracle.alf_picture(planes, tab, ctu, log2, impl="port")
    out, used_wide = helpers.emu_alf(emu, planes, tab, ctu, log2)
    assert bool(used_wide) == wide
    helpers.assert_planes_equal(out, ref, "emulated ALF kernel vs oracle")


@pytest.mark.parametrize("w,h,log2,kind,wide", CASES[:7])
def test_emulated_sao_matches_oracle(oracle, emu, w, h, log2, kind, wide):
    planes, sao = helpers.sao_case(oracle, w, h, 13 * w + h, log2, kind)
    ref = oracle.sao_picture(planes, sao, log2, impl="port")
    helpers.assert_planes_equal(helpers.emu_sao(emu, planes, sao, log2), ref, "emulated SAO kernel vs oracle")
