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


# ---- the output stage inside the drop-in: pp_process_frame (post_proc.c:76) taken over by shim/rcn_cuda.c ------------------------
import test_dropin as td  # noqa: E402

pp_streams = pytest.mark.skipif(not os.path.exists(os.path.join(td.REF, "dectest_gen")), reason="oracle/_ref (compiled reference + generative decoder) was not built")


def _pp_stream(stream_dir, name):
    path = os.path.join(str(stream_dir), name + ".266")
    if not os.path.exists(path):
        if name == "fg":          # the CTC toolset + a film grain characteristics SEI in front of the first picture
            td.vs.make_stream(path, 416, 240, 9, gop=8, seed=41, preset="ctc", verify=True, film_grain=True)
        else:                     # pictures of three sizes, decoded with -u 1: every output frame is brought to 416 x 240
            td.vs.make_stream(path, 416, 240, 9, gop=8, seed=42, preset="rpr", verify=True, rpr_sizes=[(208, 120), (320, 192)])
    return path


def _check_output_stage(stream_dir, binary, name, extra, what, threads):
    stream = _pp_stream(stream_dir, name)
    base = os.path.join(str(stream_dir), name)
    want, _ = td.decode(os.path.join(td.REF, "dectest_c"), stream, base + "_c.yuv", extra=extra)
    got, err = td.decode(binary, stream, base + "_%s_t%d.yuv" % (os.path.basename(binary), threads), threads=threads, backend="cuda", extra=extra, verbose=True)
    assert err.count("output stage of POC") == 9 and what in err and "FAILED" not in err, err[-1500:]      # taken by the back end, not the stock path
    assert got == want, "%s: %s decodes to %s, the stock decoder to %s" % (name, os.path.basename(binary), got, want)
    return want


@pp_streams
@pytest.mark.parametrize("name,extra,what", [("fg", (), "film grain"), ("rpr_up", ("-u", "1"), "up-scaling")])
def test_output_stage_in_the_dropin_emulated(stream_dir, name, extra, what):
    """film grain SEI / dectest -u 1 through build/dectest_emu: the frames come from the (emulated) device mirrors"""
    binary = os.path.join(td.BUILD, "dectest_emu")
    if not os.path.exists(binary):
        pytest.skip("build/dectest_emu not built (python __graft_entry__.py)")
    for threads in (1, 4):
        _check_output_stage(stream_dir, binary, name, extra, what, threads)
    if name == "fg":              # the SEI really changes the output: the same stream without its SEI NAL decodes differently
        stream = _pp_stream(stream_dir, name)
        data = open(stream, "rb").read()
        i = data.find(b"\x00\x00\x01\x00\xb9")            # NAL header of the prefix SEI (type 23, temporal id 0)
        j = data.find(b"\x00\x00\x01", i + 3)
        assert i > 0 and j > i
        bare = os.path.join(str(stream_dir), "fg_bare.266")
        open(bare, "wb").write(data[:i] + data[j:])
        a, _ = td.decode(os.path.join(td.REF, "dectest_c"), bare, bare + ".yuv")
        b, _ = td.decode(os.path.join(td.REF, "dectest_c"), stream, bare + ".fg.yuv")
        assert a != b


@pytest.fixture(scope="module")
def stream_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("pp_streams")


@pp_streams
@pytest.mark.gpu
@pytest.mark.parametrize("name,extra,what", [("fg", (), "film grain"), ("rpr_up", ("-u", "1"), "up-scaling")])
def test_output_stage_in_the_dropin_gpu(stream_dir, name, extra, what):
    binary = os.path.join(td.BUILD, "dectest")
    assert os.path.exists(binary), "build/dectest missing: run python __graft_entry__.py where /root/reference exists"
    for threads in (1, 4):
        _check_output_stage(stream_dir, binary, name, extra, what, threads)


# ---- tile edges as filter borders: SAO --------------------------------------------------------------------
# (w, h, log2_ctu, tile column boundaries, tile row boundaries in CTUs, seed); the third grid has tiles of a single CTU row
TILE_CASES = [(416, 240, 6, [0, 3, 7], [0, 2, 4], 1), (832, 480, 7, [0, 2, 5, 7], [0, 1, 4], 2), (416, 240, 5, [0, 5, 13], [0, 1, 2, 8], 3),
              (1920, 1080, 7, [0, 4, 9, 15], [0, 3, 9], 4)]


@pytest.mark.parametrize("w,h,log2,cols,rows,seed", TILE_CASES)
def test_sao_tile_edges_port_vs_compiled_reference_and_emulated_kernel(po, emu, w, h, log2, cols, rows, seed):
    """the reference driven one rect entry per tile (rcn_sao_first_pix_rows / rcn_sao_filter_line) == the oracle with a tile grid ==
    SaoK's source with the grid; and the grid matters (the result differs from the one-tile picture)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    planes, sao = helpers.sao_case(po, w, h, seed, log2)
    tg = po.make_tile_grid(cols, rows)
    want = po.sao_picture(planes, sao, log2, impl="port", tiles=tg)
    assert any(not np.array_equal(a, b) for a, b in zip(want, po.sao_picture(planes, sao, log2, impl="port")))
    if po.ref_available():
        for simd in (0, 1):
            got = po.sao_picture(planes, sao, log2, impl="ref", simd=simd, tiles=tg)
            if simd == 0 or log2 == 7:            # (the SIMD chroma SAO differs from the C code on non-CTC geometries)
                for a, b in zip(want, got):
                    assert np.array_equal(a, b), "simd=%d" % simd
    out = [np.zeros_like(p) for p in planes]
    dst, ds = po._planes_arg(out)
    src, ss = po._planes_arg(planes)
    emu.emu_sao_picture_tiles(dst, ds, src, ss, w, h, log2, 10, C.c_void_p(sao.ctypes.data), C.byref(tg))
    for a, b in zip(want, out):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,cols,rows,seed", TILE_CASES)
def test_gpu_sao_tile_edges(po, w, h, log2, cols, rows, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from openvvc_b200 import api
    ctx = api.Ctx(0)
    try:
        planes, sao = helpers.sao_case(po, w, h, seed, log2)
        want = po.sao_picture(planes, sao, log2, impl="port", tiles=po.make_tile_grid(cols, rows))
        src = ctx.picture(w, h); src.upload(planes)
        dst = ctx.picture(w, h)
        prm = api.SaoParams(ctx, w, h, log2)
        prm.upload(sao.astype(api.SAO_CTU))
        prm.set_tiles(cols, rows)
        prm.run(dst, src)
        for a, b in zip(want, dst.download()):
            assert np.array_equal(a, b)
        with pytest.raises(api.OVGPUError):           # a grid that does not end at the picture's last CTU
            prm.set_tiles(cols[:-1] + [cols[-1] + 1], rows)
        prm.set_tiles(None)                           # back to one tile
        prm.run(dst, src)
        for a, b in zip(po.sao_picture(planes, sao, log2, impl="port"), dst.download()):
            assert np.array_equal(a, b)
        prm.destroy(); src.destroy(); dst.destroy()
    finally:
        ctx.close()


# ---- tile edges as filter borders: ALF (classification, luma 7x7, chroma 5x5, CC-ALF) -----------------------------------------------
ALF_TILE_CASES = [(416, 240, 6, [0, 3, 7], [0, 2, 4], 1, "natural"), (832, 480, 7, [0, 2, 5, 7], [0, 1, 4], 2, "noise"),
                  (416, 240, 5, [0, 5, 13], [0, 1, 2, 8], 3, "noise"), (1920, 1080, 7, [0, 4, 9, 15], [0, 3, 9], 4, "natural")]


@pytest.mark.parametrize("w,h,log2,cols,rows,seed,kind", ALF_TILE_CASES)
def test_alf_tile_edges_port_vs_compiled_reference_and_emulated_kernels(po, emu, w, h, log2, cols, rows, seed, kind):
    """the reference driven one rect entry per tile (rcn_alf_filter_line pads an entry's edges by replication) == the oracle with a
    tile grid == the ALF kernels' source with the grid"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    planes, tab, ctu, _ = helpers.alf_case(po, w, h, seed, log2, kind)
    tg = po.make_tile_grid(cols, rows)
    want = po.alf_picture(planes, tab, ctu, log2, impl="port", tiles=tg)
    assert all(not np.array_equal(a, b) for a, b in zip(want, po.alf_picture(planes, tab, ctu, log2, impl="port")))
    if po.ref_available():
        got = po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=0, tiles=tg)
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
    out = [np.zeros_like(p) for p in planes]
    src, ss = helpers.planes_arg(planes); dst, ds = helpers.planes_arg(out)
    rows_ = helpers.alf_ctu_rows(ctu)
    u8p = C.POINTER(C.c_uint8)
    cc0 = np.ascontiguousarray(ctu["cc_cb"], dtype=np.uint8); cc1 = np.ascontiguousarray(ctu["cc_cr"], dtype=np.uint8)
    cc = (u8p * 2)(cc0.ctypes.data_as(u8p), cc1.ctypes.data_as(u8p))
    emu.emu_alf_picture_tiles(dst, ds, src, ss, w, h, log2, 10, C.byref(tab), C.c_void_p(rows_.ctypes.data), cc,
                              int(ctu["cc_cb_enabled"]), int(ctu["cc_cr_enabled"]), C.byref(tg))
    for a, b in zip(want, out):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,log2,cols,rows,seed,kind", ALF_TILE_CASES)
def test_gpu_alf_tile_edges(po, w, h, log2, cols, rows, seed, kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from openvvc_b200 import api
    ctx = api.Ctx(0)
    try:
        planes, tab, ctu, _ = helpers.alf_case(po, w, h, seed, log2, kind)
        want = po.alf_picture(planes, tab, ctu, log2, impl="port", tiles=po.make_tile_grid(cols, rows))
        src = ctx.picture(w, h); src.upload(planes)
        dst = ctx.picture(w, h)
        prm = api.AlfParams(ctx, w, h, log2)
        prm.upload(helpers.gpu_tables(tab), api.alf_ctu_array(ctu), (ctu["cc_cb"], ctu["cc_cr"]), (ctu["cc_cb_enabled"], ctu["cc_cr_enabled"]))
        prm.set_tiles(cols, rows)
        prm.run(dst, src)
        for a, b in zip(want, dst.download()):
            assert np.array_equal(a, b)
        prm.set_tiles(None)
        prm.run(dst, src)
        for a, b in zip(po.alf_picture(planes, tab, ctu, log2, impl="port"), dst.download()):
            assert np.array_equal(a, b)
        prm.destroy(); src.destroy(); dst.destroy()
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------ pictures of several tiles, drop-in
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(td.TILE_STREAMS) + sorted(td.TILE_STREAMS_GPU))
def test_tiles_dropin_gpu(stream_dir, name):
    """the product on pictures of several tiles: build/libovvc.so over libovvc_b200.so == stock decoder (tests/test_dropin.py has
    the emulated form; kept in this file so that it runs after the single-tile drop-in tests)"""
    binary = os.path.join(td.BUILD, "dectest")
    assert os.path.exists(binary), "build/dectest missing: run python __graft_entry__.py where /root/reference exists"
    td.check_tiles(stream_dir, name, binary, td.TILE_STREAMS if name in td.TILE_STREAMS else td.TILE_STREAMS_GPU, configs=((1, 1), (4, 4)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(td.SMALL_STREAMS))
def test_small_pictures_gpu(stream_dir, name):
    """pictures of one CTU row / a few CTUs through the product (emulated form in tests/test_dropin.py)"""
    binary = os.path.join(td.BUILD, "dectest")
    assert os.path.exists(binary), "build/dectest missing: run python __graft_entry__.py where /root/reference exists"
    td.check_dropin(stream_dir, name, binary, 1, td.SMALL_STREAMS)
    td.check_dropin(stream_dir, name, binary, 4, td.SMALL_STREAMS)
