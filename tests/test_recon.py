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


class Lm:
    """LMCS state of a case: with chroma residual scaling on (ict_type & 1) the scales are NOT given with the transform
    units: they are derived from the reconstructed luma at the recorded anchors, by the oracle in CU order and by the product
    between its luma and chroma passes"""
    def __init__(self, po, cmds, tus, log2, ict, seed):
        self.on = bool(ict & 1)
        self.anchors, self.luts, self.data, self.tus = None, None, None, tus
        if self.on:
            d = synth.lmcs_data(seed)
            self.data = po.make_lmcs_data(d)
            self.luts = po.lmcs_build_luts(d, impl="port")
            self.anchors, self.tus = synth_recon.lmcs_anchors(cmds, tus, log2)
            self.tus["chroma_scale"] = 77                    # must not be used


def _case(po, w, h, log2, ict, p_cclm):
    cmds, tus, arena = synth_recon.recon_case(w, h, 9 * w + h, log2, ict_type=ict, p_cclm=p_cclm)
    lm = Lm(po, cmds, tus, log2, ict, 5 * w + h)
    return synth.planes(w, h, 3 * w + h), cmds, lm.tus, arena, lm


def _want(pr, planes, cmds, tus, arena, log2, ict, lm):
    return pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, log2, ict_type=ict, anchors=lm.anchors, luts=lm.luts, lmcs_data=lm.data)


def _emu(emu, planes, cmds, tus, arena, log2, ict, lm, isp=None):
    out = [p.copy() for p in planes]
    h, w = out[0].shape
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    gt = np.ascontiguousarray(tus[tus["cbf_mask"] != 0])
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in out])
    strides = (C.c_int * 3)(*[p.strides[0] // 2 for p in out])
    nl = emu.emu_recon_picture(ptrs, strides, w, h, log2, 10, gc.ctypes.data_as(C.c_void_p), len(gc), gt.ctypes.data_as(C.c_void_p), len(gt),
                               arena.ctypes.data_as(C.c_void_p), C.c_size_t(len(arena)), 1, 0, ict,
                               lm.anchors.ctypes.data_as(C.c_void_p) if lm.on else None, len(lm.anchors) if lm.on else 0,
                               C.byref(lm.luts) if lm.on else None, C.byref(lm.data) if lm.on else None,
                               isp.ctypes.data_as(C.c_void_p) if isp is not None else None, len(isp) if isp is not None else 0)
    assert nl >= 0, nl
    return out, nl


@pytest.mark.parametrize("w,h,log2,ict,p_cclm", CASES, ids=IDS)
def test_emulated_schedule_matches_cu_order(oracle, pr, emu, w, h, log2, ict, p_cclm):
    planes, cmds, tus, arena, lm = _case(oracle, w, h, log2, ict, p_cclm)
    want = _want(pr, planes, cmds, tus, arena, log2, ict, lm)
    if lm.on:                                                   # the derived scales vary over the picture and matter
        flat = tus.copy(); flat["flags"] &= 0xf7; flat["chroma_scale"] = 2048
        other = pr.recon_picture([p.copy() for p in planes], cmds, flat, arena, log2, ict_type=ict)
        assert (other[1] != want[1]).any() and (other[0] == want[0]).all()
    # the order matters: all predictions first, then all residuals, is a different picture
    from oracle import pyoracle_intra as pi, pyoracle_tr as pt
    naive = pi.intra_picture([p.copy() for p in planes], cmds, log2, impl="port")
    pt.residual_picture(naive, tus[tus["cbf_mask"] != 0], arena, log2, impl="port", ict_type=ict)
    assert any((naive[k] != want[k]).any() for k in range(3))
    for rev in (0, 1):
        emu.emu_set_reverse_block_order(rev)
        try:
            got, nl = _emu(emu, planes, cmds, tus, arena, log2, ict, lm)
        finally:
            emu.emu_set_reverse_block_order(0)
        for k in range(3):
            assert (got[k] == want[k]).all(), "plane %d: %d samples differ (reverse=%d)" % (k, (got[k] != want[k]).sum(), rev)
        assert nl > 1


@pytest.mark.parametrize("ict", [0, 1])
def test_inter_only_picture_is_plain_residual(oracle, pr, emu, ict):
    """no intra commands: the schedule degenerates to the residual pass (luma, chroma scales, chroma with LMCS)"""
    planes, cmds, tus, arena, lm = _case(oracle, 416, 240, 7, ict, 0.0)
    cmds["flags"] = 0
    want = _want(pr, planes, cmds, tus, arena, 7, ict, lm)
    got, nl = _emu(emu, planes, cmds, tus, arena, 7, ict, lm)
    assert nl == 0 and all((got[k] == want[k]).all() for k in range(3))


def test_chroma_scale_port_vs_compiled_reference(oracle, pr):
    """rcn_lmcs_compute_chroma_scale: port == the reference's slot, incl. partial and non-contiguous availability masks"""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    seen = set()
    for t in range(200):
        d = synth.lmcs_data(100 + t)
        data, luts = oracle.make_lmcs_data(d), oracle.lmcs_build_luts(d, impl="port")
        luma = synth.planes(256, 192, t)[0]
        x0, y0 = 64 * int(rng.integers(1, 3)), 64
        am = [0, 0xffff, 0x00ff, 0x0001, 0x7fff, 0x0f0f][int(rng.integers(0, 6))]
        lm = [0, 0xffff, 0x003f, 0x0fff][int(rng.integers(0, 4))]
        a = pr.lmcs_chroma_scale(luma, x0, y0, am, lm, luts, data, impl="port")
        assert a == pr.lmcs_chroma_scale(luma, x0, y0, am, lm, luts, data, impl="ref")
        seen.add(a)
    assert len(seen) > 20


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,ict,p_cclm", CASES + [(1920, 1080, 7, 1, 0.3)], ids=IDS + ["1080p"])
def test_gpu_recon_matches_cu_order(oracle, pr, gpu_ctx, w, h, log2, ict, p_cclm):
    from openvvc_b200 import api
    planes, cmds, tus, arena, lm = _case(oracle, w, h, log2, ict, p_cclm)
    want = _want(pr, planes, cmds, tus, arena, log2, ict, lm)
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    gt = np.ascontiguousarray(tus[tus["cbf_mask"] != 0])
    pic = gpu_ctx.picture(w, h)
    pic.upload(planes)
    ib = api.IntraBatch(gpu_ctx, max(1, len(gc)))
    ib.upload(gc, w, h, log2)
    tb = api.TuBatch(gpu_ctx, max(1, len(gt)), len(arena))
    lmcs = None
    if lm.on:
        tb.set_anchors(lm.anchors)
        gl = api.LmcsLuts()
        C.memmove(C.byref(gl), C.byref(lm.luts), C.sizeof(gl))
        lmcs = api.Lmcs(gpu_ctx, gl)
        gd = api.LmcsData()
        C.memmove(C.byref(gd), C.byref(lm.data), C.sizeof(gd))
        lmcs.set_data(gd)
    tb.upload(gt, arena, 1, 0, ict)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib, lmcs=lmcs)
    got = pic.download()
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())
    # a second run on a fresh prediction re-uses the binding; the plain residual entry point still works on a bound batch
    pic.upload(planes)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib, lmcs=lmcs)
    got2 = pic.download()
    assert all((got2[k] == want[k]).all() for k in range(3))
    ib.destroy(); tb.destroy(); pic.destroy()


# ------------------------------------------------------------------------------------------------ ISP inside the schedule
def _isp_case(w, h, log2, seed):
    """single-tree intra picture with ISP coding units: commands, ISP words, all transform units (the partitions' blocks flagged
    OVGPU_TU_ISP in decode order), arena; oracle = ovo_intra_picture2 (ISP CUs reconstructed, the others predicted)"""
    import helpers_intra
    planes, cmds, words, tus, arena = helpers_intra.isp_case(w, h, seed, log2)
    flat = np.ascontiguousarray(tus.reshape(-1)[tus.reshape(-1)["cbf_mask"] != 0])
    return planes, cmds, words, tus, flat, arena


@pytest.mark.parametrize("w,h,log2", [(128, 128, 7), (416, 240, 7), (832, 480, 7), (264, 136, 6), (200, 120, 5)])
def test_emulated_isp_items_match_oracle(oracle, emu, w, h, log2):
    from oracle import pyoracle_intra as pi
    planes, cmds, words, tus, flat, arena = _isp_case(w, h, log2, 5 * w + h + 3)
    want = pi.intra_picture_isp([p.copy() for p in planes], cmds, pi.isp_info_array(words, tus), arena, log2, impl="port")
    got, nl = _emu(emu, planes, cmds, flat, arena, log2, 0, Lm(oracle, cmds, flat, log2, 0, 1), isp=np.ascontiguousarray(words))
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2", [(128, 128, 7), (416, 240, 7), (832, 480, 7), (264, 136, 6), (200, 120, 5), (1920, 1080, 7)])
def test_gpu_isp_matches_oracle(oracle, gpu_ctx, w, h, log2):
    from oracle import pyoracle_intra as pi
    from openvvc_b200 import api
    planes, cmds, words, tus, flat, arena = _isp_case(w, h, log2, 5 * w + h + 3)
    want = pi.intra_picture_isp([p.copy() for p in planes], cmds, pi.isp_info_array(words, tus), arena, log2, impl="port")
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    pic = gpu_ctx.picture(w, h)
    pic.upload(planes)
    ib = api.IntraBatch(gpu_ctx, max(1, len(gc)))
    ib.upload(gc, w, h, log2, isp=words)
    tb = api.TuBatch(gpu_ctx, max(1, len(flat)), len(arena))
    tb.upload(flat, arena, 1, 0, 0)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib)
    got = pic.download()
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())
    with pytest.raises(api.OVGPUError):          # prediction alone is not defined for ISP commands
        ib.run(pic)
    ib.destroy(); tb.destroy(); pic.destroy()


def _mixed_isp_case(po, w, h, log2, ict):
    cmds, tus, arena, words, isp_tus = synth_recon.recon_case(w, h, 13 * w + h, log2, ict_type=ict, p_cclm=0.3, p_mip=0.1, p_isp=0.4, p_bdpcm=0.8, p_ibc=0.15)
    lm = Lm(po, cmds, tus, log2, ict, 5 * w + h)
    flat = np.ascontiguousarray(np.concatenate([lm.tus[lm.tus["cbf_mask"] != 0], isp_tus.reshape(-1)[isp_tus.reshape(-1)["cbf_mask"] != 0]]))
    return synth.planes(w, h, 3 * w + h), cmds, lm.tus, flat, arena, words, isp_tus, lm


@pytest.mark.parametrize("w,h,log2,ict", [(416, 240, 7, 1), (832, 480, 7, 0), (264, 136, 6, 1)])
def test_emulated_mixed_picture_with_isp(oracle, pr, emu, w, h, log2, ict):
    """everything at once: inter CUs with residual, CIIP, plain / MRL / MIP / CCLM / ISP / BDPCM intra CUs, intra block copy, LMCS chroma scales from luma"""
    from oracle import pyoracle_intra as pi
    planes, cmds, tus, flat, arena, words, isp_tus, lm = _mixed_isp_case(oracle, w, h, log2, ict)
    assert ((cmds["flags"] & 32) != 0).sum() > 5 and (tus["bdpcm"] != 0).sum() > 2 and ((cmds["flags"] & 128) != 0).sum() > 5
    want = pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, log2, ict_type=ict, anchors=lm.anchors, luts=lm.luts,
                            lmcs_data=lm.data, isp=pi.isp_info_array(words, isp_tus))
    got, _ = _emu(emu, planes, cmds, flat, arena, log2, ict, lm, isp=np.ascontiguousarray(words))
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,ict", [(416, 240, 7, 1), (832, 480, 7, 0), (264, 136, 6, 1), (1920, 1080, 7, 1)])
def test_gpu_mixed_picture_with_isp(oracle, pr, gpu_ctx, w, h, log2, ict):
    from oracle import pyoracle_intra as pi
    from openvvc_b200 import api
    planes, cmds, tus, flat, arena, words, isp_tus, lm = _mixed_isp_case(oracle, w, h, log2, ict)
    want = pr.recon_picture([p.copy() for p in planes], cmds, tus, arena, log2, ict_type=ict, anchors=lm.anchors, luts=lm.luts,
                            lmcs_data=lm.data, isp=pi.isp_info_array(words, isp_tus))
    gc = np.ascontiguousarray(cmds[cmds["flags"] != 0])
    pic = gpu_ctx.picture(w, h)
    pic.upload(planes)
    ib = api.IntraBatch(gpu_ctx, max(1, len(gc)))
    ib.upload(gc, w, h, log2, isp=words)
    tb = api.TuBatch(gpu_ctx, max(1, len(flat)), len(arena))
    lmcs = None
    if lm.on:
        tb.set_anchors(lm.anchors)
        gl = api.LmcsLuts(); C.memmove(C.byref(gl), C.byref(lm.luts), C.sizeof(gl))
        lmcs = api.Lmcs(gpu_ctx, gl)
        gd = api.LmcsData(); C.memmove(C.byref(gd), C.byref(lm.data), C.sizeof(gd))
        lmcs.set_data(gd)
    tb.upload(flat, arena, 1, 0, ict)
    api.recon_picture(gpu_ctx, pic, tu=tb, intra=ib, lmcs=lmcs)
    got = pic.download()
    for k in range(3):
        assert (got[k] == want[k]).all(), "plane %d: %d samples differ" % (k, (got[k] != want[k]).sum())
    ib.destroy(); tb.destroy(); pic.destroy()


def test_position_maps_are_reused_without_full_clears(emu):
    """the host keeps its level / position maps zero between pictures by clearing only the blocks it wrote: a second picture
    on the same maps must give the lists a fresh build gives"""
    a = synth_recon.recon_case(416, 240, 31, 7, p_cclm=0.3, p_ibc=0.1)[0]
    b = synth_recon.recon_case(416, 240, 32, 7, p_cclm=0.3, p_inter=0.7, p_ibc=0.1)[0]
    a = np.ascontiguousarray(a[a["flags"] != 0]); b = np.ascontiguousarray(b[b["flags"] != 0])
    assert emu.emu_levels_reuse_check(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), 416, 240) == 0
    assert emu.emu_levels_reuse_check(b.ctypes.data_as(C.c_void_p), len(b), a.ctypes.data_as(C.c_void_p), len(a), 416, 240) == 0
