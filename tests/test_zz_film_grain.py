"""Output stage: film grain synthesis (fg_grain_apply_pic, pp_film_grain.c) -- the oracle restatement against
the compiled reference's own function and pattern data base, the kernels' source emulated on the CPU against the oracle, and the
sm_100a kernels through the C ABI.  (Last in the collection order on purpose: the newest GPU test of the tree.)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (w, h, seed, poc, kind, components with a model): 1080p has a chroma height of 540 = 33 x 16 + 12 (a partial block row)
CASES = [(416, 240, 1, 0, "natural", (1, 1, 1)), (416, 240, 2, 17, "noise", (1, 0, 1)), (832, 480, 5, 255, "blocky", (1, 1, 0)),
         (1920, 1080, 3, 300, "natural", (1, 1, 1)), (432, 248, 6, -7, "noise", (1, 1, 1))]


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return C.CDLL(os.path.join(d, "libovvc_b200_emu.so"))


def test_database_port_vs_compiled_reference(po):
    """the 13 x 13 grain patterns (fg_data_base_generation) == the reference's static data base, read back window by window"""
    a = po.fg_database("port")
    assert int(np.abs(a.astype(int)).max()) == 127 and len(np.unique(a)) > 200
    if po.ref_available():
        assert np.array_equal(a, po.fg_database("ref"))


@pytest.mark.parametrize("w,h,seed,poc,kind,present", CASES)
def test_port_vs_compiled_reference_and_emulated_kernels(po, emu, w, h, seed, poc, kind, present):
    from openvvc_b200 import synth
    pl = synth.planes(w, h, seed, kind)
    fg = po.make_film_grain(seed, present)
    want = po.fg_apply(pl, fg, poc, "port")
    assert any(not np.array_equal(a, b) for a, b in zip(want, pl)), "the case synthesises no grain at all"
    for c in range(3):
        if not present[c]:
            assert np.array_equal(want[c], pl[c])
    if po.ref_available():
        got = po.fg_apply(pl, fg, poc, "ref")
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
    out = [np.zeros_like(p) for p in pl]
    dst, ds = po._planes_arg(out)
    src, ss = po._planes_arg(pl)
    assert emu.emu_film_grain(dst, ds, src, ss, w, h, C.byref(fg), poc) == 0
    for a, b in zip(want, out):
        assert np.array_equal(a, b)


def test_sizes_the_reference_reads_out_of_bounds_are_refused(po, emu):
    """a block row / column of fewer than 8 samples (size % 16 in 1..7) makes the reference's unsigned min() wrap: refused"""
    from openvvc_b200 import synth
    w, h = 264, 136                      # chroma height 68 = 4 x 16 + 4
    pl = synth.planes(w, h, 1)
    fg = po.make_film_grain(1)
    out = [np.zeros_like(p) for p in pl]
    dst, ds = po._planes_arg(out)
    src, ss = po._planes_arg(pl)
    assert emu.emu_film_grain(dst, ds, src, ss, w, h, C.byref(fg), 0) == -1
    assert po.port().ovo_fg_size_ok(w, h) == 0 and po.port().ovo_fg_size_ok(1920, 1080) == 1 and po.port().ovo_fg_size_ok(3840, 2160) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,seed,poc,kind,present", CASES + [(3840, 2160, 9, 1000, "natural", (1, 1, 1))])
def test_gpu_matches_oracle(po, w, h, seed, poc, kind, present):
    from openvvc_b200 import api, synth
    ctx = api.Ctx(0)
    try:
        pl = synth.planes(w, h, seed, kind)
        fg = po.make_film_grain(seed, present)
        want = po.fg_apply(pl, fg, poc, "port")
        src = ctx.picture(w, h); src.upload(pl)
        dst = ctx.picture(w, h)
        synth_ = api.FilmGrainSynth(ctx)
        synth_.picture(dst, src, fg, poc)
        got = dst.download()
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        small = ctx.picture(264, 136)
        with pytest.raises(api.OVGPUError):
            synth_.picture(small, small, fg, poc)
        synth_.destroy(); src.destroy(); dst.destroy(); small.destroy()
    finally:
        ctx.close()
