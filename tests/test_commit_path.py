"""The path bench.py's end-to-end number is measured on, checked against the oracle: decoder state written into the PINNED
staging buffers (ovgpu_*_staging), every *_commit, the stage calls and the download -- from several host threads, each with its
own context / stream, sharing the LMCS object and the reference pictures, two picture sets alternating (bench.py: e2e_worker).
Plus the whole chain at 7680x4320 against the compiled reference's own slots, and an all-intra picture."""
import os
import sys
import threading

import numpy as np
import pytest

import helpers
from helpers_chain import oracle_picture

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def _destroy(d):
    for k in ("rec", "mid", "dst", "alf", "sao", "mc", "tu", "dbf", "intra"):
        if d[k]:
            d[k].destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("size,n_ctx", [("1080p", 4), ("4k", 4)])
def test_gpu_commit_path_concurrent_contexts_match_oracle(oracle, gpu_ctx, size, n_ctx):
    from openvvc_b200 import api
    w, h = bench.SIZES[size]
    n_sets, rounds = 2, 4
    wl = bench.make_workload(w, h, n_ctx, seed0=7300)
    pics = wl["pics"]
    refs = [[gpu_ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
    for lst, src in zip(refs, wl["refs"]):
        for pic, planes in zip(lst, src):
            pic.upload(planes)
    lmcs = api.Lmcs(gpu_ctx, wl["lmcs"])                       # shared by every context, like the slice's LUTs
    lmcs.set_data(wl["lmcs_data"])
    ctxs = [[api.Ctx(0) for _ in range(n_ctx)] for _ in range(n_sets)]       # private stream each
    devs = [[bench.make_dev(api, gpu_ctx, p, w, h, upload=False) for p in pics] for _ in range(n_sets)]
    gpu_ctx.sync()
    out = [[[np.zeros((hh, ww), np.uint16) for (hh, ww) in ((h, w), (h // 2, w // 2), (h // 2, w // 2))] for _ in pics] for _ in range(n_sets)]
    errors = []

    def worker(t):
        try:
            for r in range(rounds):
                half = r % n_sets
                c, d, p = ctxs[half][t], devs[half][t], pics[t]
                bench.stage_picture(d, p)                       # the staging call waits for the buffer's last H2D copy: rounds 2, 3 rewrite
                bench.commit_picture(api, c, d, p, w, h)        # buffers whose copies of rounds 0, 1 may still be in flight
                bench.run_stages(api, c, d, refs, wl["pocs"], lmcs)
                api._check(c.h, api.lib().ovgpu_picture_download(c.h, d["dst"].p, *api._host_planes(out[half][t])), "download")
            for half in range(n_sets):
                ctxs[half][t].sync()
        except Exception as e:      # noqa: BLE001
            errors.append((t, repr(e)))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_ctx)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
    for t, p in enumerate(pics):
        want = oracle_picture(oracle, wl, p)
        for half in range(n_sets):
            helpers.assert_planes_equal(out[half][t], want, "commit path, %s, context %d, set %d" % (size, t, half))
    for half in range(n_sets):
        for d in devs[half]:
            _destroy(d)
        for c in ctxs[half]:
            c.close()
    for lst in refs:
        for pic in lst:
            pic.destroy()


def _whole_chain(oracle, ctx, w, h, seed0, impl, all_intra=False):
    from openvvc_b200 import api
    wl = bench.make_workload(w, h, 1, seed0=seed0, all_intra=all_intra)
    p = wl["pics"][0]
    refs = [[ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
    for lst, src in zip(refs, wl["refs"]):
        for pic, planes in zip(lst, src):
            pic.upload(planes)
    lmcs = api.Lmcs(ctx, wl["lmcs"])
    lmcs.set_data(wl["lmcs_data"])
    d = bench.make_dev(api, ctx, p, w, h)
    d["rec"].upload([np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16)])
    bench.run_stages(api, ctx, d, refs, wl["pocs"], lmcs)
    got = d["dst"].download()
    want = oracle_picture(oracle, wl, p, impl=impl)
    helpers.assert_planes_equal(got, want, "whole picture %dx%d (all_intra=%s) vs oracle chain (%s)" % (w, h, all_intra, impl))
    _destroy(d)
    for lst in refs:
        for pic in lst:
            pic.destroy()
    return p


@pytest.mark.gpu
def test_gpu_whole_chain_8k_vs_reference_slots(oracle, gpu_ctx):
    """7680x4320: MC, deblocking, SAO and ALF of the checker run through the compiled reference's own slots when oracle/_ref exists"""
    _whole_chain(oracle, gpu_ctx, 7680, 4320, 7400, "ref" if oracle.ref_available() else "port")


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["1080p", "4k"])
def test_gpu_all_intra_whole_chain(oracle, gpu_ctx, size):
    """an I picture: no PU list, every coding unit intra (the dataflow kernel carries the whole picture)"""
    w, h = bench.SIZES[size]
    p = _whole_chain(oracle, gpu_ctx, w, h, 7500, "port", all_intra=True)
    assert len(p["cmds"]) == 0 and (p["icmds_all"]["flags"] != 0).all()
