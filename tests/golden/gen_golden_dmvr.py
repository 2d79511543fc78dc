"""Generates tests/golden/mc_dmvr_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C): MD5 of the planes
and of the refined MVs produced by rcn_dmvr_mv_refine (mixed with BDOF / affine / plain commands).
Run in the build container:  python tests/golden/gen_golden_dmvr.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
import helpers  # noqa: E402
import helpers_mc  # noqa: E402

CASES = [(416, 240, 801, 7, False, "natural"), (416, 240, 802, 7, True, "blocky"), (832, 480, 803, 7, True, "natural"),
         (264, 136, 804, 6, False, "natural"), (1920, 1080, 806, 7, True, "natural")]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, lmcs, kind) in CASES:
        refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_dmvr_case(po, w, h, seed, log2, lmcs, kind)
        nd = int((cmds["tool"] == 3).sum())
        mvs = np.zeros(nd, po.DmvrMv)
        dst = helpers_mc.blank(w, h)
        po.mc_picture(dst, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts, prof=prof, simd=0, dmvr_out=mvs)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, lmcs=lmcs, kind=kind, nb_cmds=len(cmds), nb_dmvr=nd,
                        input_md5=helpers.md5(cmds, prof, *refs0[0], *refs1[-1]), md5=[helpers.md5(p) for p in dst], mv_md5=helpers.md5(mvs)))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "mc_dmvr_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
