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


def test_results_do_not_depend_on_cta_order(oracle, emu):
    """CTAs run in any order on the GPU; the serial emulation is repeated with the block order reversed so that a kernel
    writing outside its own samples (hidden by the forward order) is caught on the CPU."""
    import helpers_mc, helpers_tr, helpers_df
    from oracle import pyoracle_tr as pt, pyoracle_df as pd
    try:
        emu.emu_set_reverse_block_order(1)
        w, h = 416, 240
        refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(oracle, w, h, 41, 7, True, 2)
        a = helpers_mc.blank(w, h)
        oracle.mc_picture(a, [refs0, refs1], pocs, cmds, 7, impl="port", lmcs=luts, prof=prof)
        helpers.assert_planes_equal(helpers_mc.emu_mc(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts, prof=prof), a, "MC, reversed CTA order")
        refs0, refs1, pocs, cmds, luts = helpers_mc.mc_case(oracle, w, h, 42, 7, False, 2)
        a = helpers_mc.blank(w, h)
        oracle.mc_picture(a, [refs0, refs1], pocs, cmds, 7, impl="port", lmcs=luts)
        helpers.assert_planes_equal(helpers_mc.emu_mc(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts), a, "MC 4x4, reversed CTA order")
        planes, tus, arena = helpers_tr.tr_case(w, h, 43, 7, 1, 0, 1, p_intra=0.3)
        want = pt.residual_picture([p.copy() for p in planes], tus, arena, 7, impl="port", dep_quant=1, sh_ts_disabled=0, ict_type=1)
        helpers.assert_planes_equal(helpers_tr.emu_residual(emu, planes, tus, arena, 1, 0, 1), want, "residual, reversed CTA order")
        planes, blks, maps = helpers_df.dbf_case(w, h, 44, 7, "blocky")
        want = pd.dbf_picture_port([p.copy() for p in planes], maps)
        helpers.assert_planes_equal(helpers_df.emu_dbf(emu, planes, maps), want, "DBF, reversed CTA order")
        planes, sao = helpers.sao_case(oracle, w, h, 45, 7)
        helpers.assert_planes_equal(helpers.emu_sao(emu, planes, sao, 7), oracle.sao_picture(planes, sao, 7, impl="port"), "SAO, reversed CTA order")
        # the kernels added last: motion compensation from scaled references, the up-scaler and the film grain synthesis
        import ctypes as C
        from openvvc_b200 import synth
        refs0, refs1, pocs, cmds, luts, prof, sc = helpers_mc.rpr_case(oracle, w, h, 46, 7, True, "noise")
        a = helpers_mc.blank(w, h)
        oracle.mc_picture_rpr(a, [refs0, refs1], pocs, cmds, sc, 7, impl="port", lmcs=luts, prof=prof)
        helpers.assert_planes_equal(helpers_mc.emu_mc_rpr(emu, oracle, w, h, refs0, refs1, pocs, cmds, luts, sc, prof=prof), a, "RPR MC, reversed CTA order")
        pl = synth.planes(272, 160, 47, "noise")
        want = oracle.pp_upscale(pl, w, h, (4, 4, 2, 6), 0, 0, "port")
        out = helpers_mc.blank(w, h)
        dst, ds = oracle._planes_arg(out); src, ss = oracle._planes_arg(pl)
        assert emu.emu_pp_upscale(dst, ds, w, h, src, ss, 272, 160, (C.c_int * 6)(4, 4, 2, 6, 0, 0)) == 0
        helpers.assert_planes_equal(out, want, "up-scaler, reversed CTA order")
        pl = synth.planes(w, h, 48, "noise")
        fg = oracle.make_film_grain(48)
        want = oracle.fg_apply(pl, fg, 5, "port")
        out = helpers_mc.blank(w, h)
        dst, ds = oracle._planes_arg(out); src, ss = oracle._planes_arg(pl)
        assert emu.emu_film_grain(dst, ds, src, ss, w, h, C.byref(fg), 5) == 0
        helpers.assert_planes_equal(out, want, "film grain, reversed CTA order")
    finally:
        emu.emu_set_reverse_block_order(0)
