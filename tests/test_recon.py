"""Reconstruction order (ovgpu_recon_picture): intra blocks must read neighbours that hold prediction + residual, as in the
reference's CU loop (prediction, then rcn_transform_tree, coding unit by coding unit).  Oracle = oracle/port_recon.c (the two
pinned pieces, intra prediction and residual, applied CU by CU in decode order); checked against the emulated kernels with the
product's level schedule (CPU) and the sm_100a kernels through the C ABI (GPU)."""
import ctypes as C

import numpy as np
import pytest

import helpers
from openvvc_b200 import synth, synth_recon

CASES = [(128, 128, 7, 0, 0.3), (416, 240, 7, 1, 0.3), (264, 136, 6, 0, 0.0), (200, 120, 5, 3, 0.3), (832, 480, 7, 1, 0.3)]
IDS = ["%dx%d_ctu%d_ict%d%s" % (w, h, 1 << l, i, "_cclm" if p else "") for w, h, l, i, p in CASES]


@pytest.fixture(scope="module")
def pr(oracle):
    from oracle import pyoracle_recon
    return pyoracle_recon


def _case(w, h, log2, ict, p_cclm):
    cmds, tus, arena = synth_recon.recon_case(w, h, 9 * w + h, log2, ict_type=ict, p_cclm=p_cclm)
    return synth.planes(w, h, 3 * w + h), cmds, tus, arena


def _emu(emu, planes, cmds, tus, arena, log2, ict):
    out = [p.copy() for p in planes]
    h, w = out[0].shape
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    gt = np.ascontiguousarray(tus[tus["cbf_mask"] != 0])
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    nl = emu.emu_recon_picture(ptrs, strides, w, h, log2, 10, gc.ctypes.data_as(C.c_void_p), len(gc), gt.ctypes.data_as(C.c_void_p), len(gt),
                               arena.ctypes.data_as(C.c_void_p), C.c_size_t(len(arena)), 1, 0, ict)
    assert nl >= 0, nl
    return out, nl


@pytest.mark.parametrize("w,h,log2,ict,p_cclm", CASES, ids=IDS)
def test_emulated_schedule_matches_cu_order(pr, emu, w, h, log2, ict, p_cclm):
    planes, cmds, tus, arena = _case(w, h, log2, ict, p_cclm)
    want = pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, log2, ict_type=ict)
    # the order matters: all predictions first, then all residuals, is a different picture
    from oracle import pyoracle_intra as pi, pyoracle_tr as pt
    naive = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    pt.residual_picture(naive, tus[tus["cbf_mask"] != 0], arena, log2, impl="port", ict_type=ict)
    assert any((naive[k] != want[k]).any() for k in range(3))
    for rev in (0, 1):
        emu.emu_set_reverse_block_order(rev)
        try:
            got, nl = _emu(emu, planes, cmds, tus, arena, log2, ict)
        finally:
            emu.emu_set_reverse_block_order(0)
        for k in range(3):
            assert (got[k] == want[k]).all(), "plane %d: %d samples differ (reverse=%d)" % (k, (got[k] != want[k]).sum(), rev)
        assert nl > 1


def test_inter_only_picture_is_plain_residual(pr, emu):
    """no intra commands: the schedule degenerates to the residual pass"""
    planes, cmds, tus, arena = _case(416, 240, 7, 1, 0.0)
    cmds["flags"] = 0
    want = pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, 7, ict_type=1)
    got, nl = _emu(emu, planes, cmds, tus, arena, 7, 1)
    assert nl == 0 and all((got[k] == want[k]).all() for k in range(3))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,ict,p_cclm", CASES + [(1920, 1080, 7, 1, 0.3)], ids=IDS + ["1080p"])
def test_gpu_recon_matches_cu_order(pr, gpu_ctx, w, h, log2, ict, p_cclm):
    from openvvc_b200 import api
    planes, cmds, tus, arena = _case(w, h, log2, ict, p_cclm)
    want = pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, log2, ict_type=ict)
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    gt = np.ascontiguousarray(tus[tus["cbf_mask"] != 0])
    pic = gpu_ctx.picture(w, h)
    pic.upload(planes)
    ib = api.IntraBatch(gpu_ctx, max(1, len(gc)))
    ib.upload(gc, w, h, log2)
    tb = api.TuBatch(gpu_ctx, max(1, len(gt)), len(arena))
    tb.upload(gt, arena, 1, 0, ict)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib)
    got = pic.download()
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())
    # a second run on a fresh prediction re-uses the binding; the plain residual entry point still works on a bound batch
    pic.upload(planes)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib)
    got2 = pic.download()
    assert all((got2[k] == want[k]).all() for k in range(3))
    ib.destroy(); tb.destroy(); pic.destroy()
