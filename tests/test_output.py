"""Output stage: conformance-window crop and the MD5 of the output byte stream computed on the GPU, against
what examples/dectest.c:376-406 writes and CI/ovnreg hashes (restated here with numpy + hashlib)."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dectest_bytes(planes, window):
    """write_decoded_frame_to_file (dectest.c:376-406): output window offsets in chroma units, luma twice that; Y, Cb, Cr rows"""
    l, r, a, b = window
    out = b""
    for c, p in enumerate(planes):
        s = 1 if c else 2
        hh, ww = p.shape
        out += np.ascontiguousarray(p[a * s:hh - b * s, l * s:ww - r * s]).astype("<u2").tobytes()
    return out


def test_dectest_restatement_is_plain_planes_without_window():
    from openvvc_b200 import synth
    pl = synth.planes(64, 32, 3)
    assert dectest_bytes(pl, (0, 0, 0, 0)) == b"".join(p.astype("<u2").tobytes() for p in pl)
    assert len(dectest_bytes(pl, (1, 2, 3, 0))) == 2 * ((64 - 6) * (32 - 6) + 2 * (32 - 3) * (16 - 3))


@pytest.fixture(scope="module")
def gpu_ctx():
    from openvvc_b200 import api
    c = api.Ctx(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,window", [(416, 240, (0, 0, 0, 0)), (416, 240, (0, 4, 0, 0)), (416, 240, (1, 2, 3, 5)), (1920, 1080 + 8, (0, 0, 0, 4)),
                                        (64, 32, (3, 0, 1, 0))])
def test_gpu_window_download_and_md5(gpu_ctx, w, h, window):
    from openvvc_b200 import api, synth
    pl = synth.planes(w, h, 77)
    pic = gpu_ctx.picture(w, h)
    pic.upload(pl)
    want = dectest_bytes(pl, window)
    got = api.download_window(gpu_ctx, pic, window, w, h)
    assert b"".join(p.astype("<u2").tobytes() for p in got) == want
    m = api.Md5(gpu_ctx)
    m.picture(pic, window)
    assert m.hexdigest() == hashlib.md5(want).hexdigest()
    # the hash runs on: a second picture appended = the MD5 of the two-frame file
    pl2 = synth.planes(w, h, 78)
    pic2 = gpu_ctx.picture(w, h)
    pic2.upload(pl2)
    m.picture(pic2, window)
    assert m.hexdigest() == hashlib.md5(want + dectest_bytes(pl2, window)).hexdigest()
    m.reset()
    m.picture(pic2, None)
    assert m.hexdigest() == hashlib.md5(dectest_bytes(pl2, (0, 0, 0, 0))).hexdigest()
    m.destroy(); pic.destroy(); pic2.destroy()


# ---- up-scaling of output pictures (pp_sample_rate_conv: the post-processing step of sequences with reference picture resampling)
UPSCALE_CASES = [(208, 120, 416, 240, (0, 0, 0, 0), 1, 0, "natural"), (320, 192, 416, 240, (0, 0, 0, 0), 0, 1, "noise"),
                 (416, 240, 416, 240, (0, 0, 0, 0), 1, 1, "noise"), (272, 160, 416, 240, (4, 4, 2, 6), 0, 0, "noise"),
                 (624, 360, 832, 480, (0, 8, 0, 0), 1, 0, "natural")]


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle


@pytest.mark.parametrize("sw,sh,dw,dh,win,hc,vc,kind", UPSCALE_CASES)
def test_upscale_port_vs_compiled_reference_and_emulated_kernel(po, sw, sh, dw, dh, win, hc, vc, kind):
    """oracle port == the reference's own pp_sample_rate_conv; UpscaleK's source emulated on the CPU == port"""
    import ctypes as C
    import subprocess
    from openvvc_b200 import synth
    pl = synth.planes(sw, sh, 11, kind)
    want = po.pp_upscale(pl, dw, dh, win, hc, vc, "port")
    if po.ref_available():
        got = po.pp_upscale(pl, dw, dh, win, hc, vc, "ref")
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    emu = C.CDLL(os.path.join(d, "libovvc_b200_emu.so"))
    out = [np.zeros((dh, dw), np.uint16), np.zeros((dh // 2, dw // 2), np.uint16), np.zeros((dh // 2, dw // 2), np.uint16)]
    dst, ds = po._planes_arg(out)
    src, ss = po._planes_arg(pl)
    assert emu.emu_pp_upscale(dst, ds, dw, dh, src, ss, sw, sh, (C.c_int * 6)(*win, hc, vc)) == 0
    for a, b in zip(want, out):
        assert np.array_equal(a, b)
    # a target smaller than the source is the reference's down-sampling branch: refused
    assert emu.emu_pp_upscale(src, ss, sw // 2, sh // 2, src, ss, sw, sh, (C.c_int * 6)(0, 0, 0, 0, 1, 0)) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("sw,sh,dw,dh,win,hc,vc,kind", UPSCALE_CASES + [(1920, 1088, 3840, 2160, (0, 0, 0, 2), 1, 0, "natural")])
def test_gpu_upscale_matches_oracle(gpu_ctx, po, sw, sh, dw, dh, win, hc, vc, kind):
    from openvvc_b200 import api, synth
    pl = synth.planes(sw, sh, 11, kind)
    want = po.pp_upscale(pl, dw, dh, win, hc, vc, "port")
    src = gpu_ctx.picture(sw, sh); src.upload(pl)
    dst = gpu_ctx.picture(dw, dh)
    api.upscale(gpu_ctx, dst, src, win, hc, vc)
    got = dst.download()
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    if (dw, dh) != (sw, sh):        # the other way round is the reference's down-sampling branch: refused
        with pytest.raises(api.OVGPUError):
            api.upscale(gpu_ctx, src, dst)
    src.destroy(); dst.destroy()
