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
