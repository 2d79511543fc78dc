"""Intra block copy (rcn_ibc_l / rcn_ibc_c, rcn_ibc.c:8-140): the oracle port (picture coordinates) pinned to the reference's
slots (CTU ring buffer with wrap) and to golden vectors made from them; the dataflow kernel's block-copy path (source units
awaited like reference samples) emulated on the CPU and run on the GPU through ovgpu_recon_picture."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers
import helpers_intra

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ibc_golden.json")))
IDS = lambda c: "%dx%d_ctu%d" % (c["w"], c["h"], 1 << c["log2_ctu"])
IBC = 128


@pytest.fixture(scope="module")
def pi(oracle):
    from oracle import pyoracle_intra
    return pyoracle_intra


def _case(c):
    planes, cmds = helpers_intra.ibc_case(c["w"], c["h"], c["seed"], c["log2_ctu"], c["chroma"])
    assert helpers.md5(cmds, *planes) == c["input_md5"], "synthetic generator drifted; regenerate golden"
    assert ((cmds["flags"] & IBC) != 0).sum() == c["nb_ibc"] > 3
    return planes, cmds


def _emu_recon(emu, planes, cmds, log2):
    """prediction-only reconstruction (no transform units) through the emulated dataflow schedule"""
    out = [p.copy() for p in planes]
    h, w = out[0].shape
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    arena = np.zeros(64, np.int16)
    nl = emu.emu_recon_picture(ptrs, strides, w, h, log2, 10, gc.ctypes.data_as(C.c_void_p), len(gc), None, 0,
                               arena.ctypes.data_as(C.c_void_p), C.c_size_t(len(arena)), 1, 0, 0, None, 0, None, None, None, 0)
    return out, nl


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_port_matches_golden(pi, c):
    planes, cmds = _case(c)
    pi.intra_picture(planes, cmds, c["log2_ctu"], impl="port")
    assert helpers.md5(*planes) == c["md5"]


@pytest.mark.parametrize("w,h,log2,chroma", [(416, 240, 7, True), (832, 480, 7, True), (264, 136, 6, True), (136, 72, 5, False)])
@pytest.mark.parametrize("simd", [0, 1])
def test_port_vs_reference_ring_buffer(oracle, pi, w, h, log2, chroma, simd):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    planes, cmds = helpers_intra.ibc_case(w, h, 17 * w + h, log2, chroma)
    bv = cmds["pad"][(cmds["flags"] & IBC) != 0][:, :4].copy().view(np.int16)
    assert len(bv) > 10 and ((bv & 3) != 0).any() and (bv[:, 0] < -(1 << log2) // 2).any()      # unaligned vectors, reach into the left CTU
    a = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="port")
    b = pi.intra_picture([p.copy() for p in planes], cmds.copy(), log2, impl="ref", simd=simd)
    helpers.assert_planes_equal(a, b, "intra block copy: port vs the reference's ring-buffer slots")
    plain = cmds.copy()
    plain["pad"][:, :4] = 0
    plain["flags"][(plain["flags"] & IBC) != 0] = 0
    assert any((x != y).any() for x, y in zip(a, pi.intra_picture([p.copy() for p in planes], plain, log2, impl="port")))


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_emulated_dataflow_matches_golden(emu, c):
    planes, cmds = _case(c)
    out, nl = _emu_recon(emu, planes, cmds, c["log2_ctu"])
    assert nl >= 0, nl
    assert helpers.md5(*out) == c["md5"]


def test_emu_rejects_copies_of_later_or_overlapping_blocks(emu):
    planes, cmds = helpers_intra.ibc_case(416, 240, 991, 7, True)
    real = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    i = int(np.nonzero((real["flags"] & IBC) != 0)[0][0])

    def with_bv(bx, by):
        c = real.copy()
        c["pad"][i, :4] = np.array([bx, by], np.int16).view(np.uint8)
        return c
    w = 1 << int(real["log2_w"][i])
    j = int(np.nonzero(real["y0"] > real["y0"][i] + 64)[0][0])                 # a command decoded later
    for bad in (with_bv(w // 2, 0),                                              # overlaps the block itself
                with_bv(-int(real["x0"][i]) - 4, 0),                             # left of the picture
                with_bv(int(real["x0"][j]) - int(real["x0"][i]), int(real["y0"][j]) - int(real["y0"][i]))):   # not reconstructed yet
        _, nl = _emu_recon(emu, planes, bad, 7)
        assert nl < 0
    _, nl = _emu_recon(emu, planes, real, 7)
    assert nl >= 0


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_recon(ctx, planes, cmds, log2):
    from openvvc_b200 import api
    h, w = planes[0].shape
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    pic = ctx.picture(w, h)
    pic.upload(planes)
    ib = api.IntraBatch(ctx, max(1, len(gc)))
    try:
        ib.upload(gc, w, h, log2)
        api.recon_picture(ctx, pic, intra=ib)
        return pic.download()
    finally:
        ib.destroy(); pic.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_matches_golden(gpu_ctx, c):
    planes, cmds = _case(c)
    assert helpers.md5(*_gpu_recon(gpu_ctx, planes, cmds, c["log2_ctu"])) == c["md5"]


@pytest.mark.gpu
def test_gpu_full_hd_matches_oracle(pi, gpu_ctx):
    planes, cmds = helpers_intra.ibc_case(1920, 1080, 993, 7, True)
    want = pi.intra_picture([p.copy() for p in planes], cmds.copy(), 7, impl="port")
    helpers.assert_planes_equal(_gpu_recon(gpu_ctx, planes, cmds, 7), want, "CUDA intra block copy vs oracle")


@pytest.mark.gpu
def test_gpu_rejects_bad_block_vectors(gpu_ctx):
    from openvvc_b200 import api
    planes, cmds = helpers_intra.ibc_case(416, 240, 991, 7, True)
    real = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    i = int(np.nonzero((real["flags"] & IBC) != 0)[0][0])
    j = int(np.nonzero(real["y0"] > real["y0"][i] + 64)[0][0])
    bad = real.copy()
    bad["pad"][i, :4] = np.array([int(real["x0"][j]) - int(real["x0"][i]), int(real["y0"][j]) - int(real["y0"][i])], np.int16).view(np.uint8)
    with pytest.raises(api.OVGPUError):        # would wait for a block that comes later: refused on the host, never launched
        _gpu_recon(gpu_ctx, planes, bad, 7)
    ib = api.IntraBatch(gpu_ctx, len(real))
    ib.upload(real, 416, 240, 7)
    pic = gpu_ctx.picture(416, 240)
    with pytest.raises(api.OVGPUError):        # prediction alone is not defined for block copies
        ib.run(pic)
    ib.destroy(); pic.destroy()
