"""Generates tests/golden/mc_rpr_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C): MD5 of the prediction
planes of seeded PU lists whose references have their own sizes (reference picture resampling: rcn_mcp_rpr_*, rcn_mc_rpr_b_*,
rcn_mc_rpr_prof_b_l, rcn_inter.c:1990-2747).
Run in the build container:  python tests/golden/gen_golden_rpr.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
import helpers  # noqa: E402
import helpers_mc  # noqa: E402

# (w, h, seed, log2_ctu, lmcs, kind, chroma_hor_col, chroma_ver_col)
CASES = [(416, 240, 901, 7, False, "natural", 1, 0), (416, 240, 902, 7, True, "noise", 0, 1), (264, 136, 903, 6, True, "natural", 1, 1),
         (832, 480, 904, 7, True, "noise", 1, 0), (1920, 1080, 905, 7, True, "natural", 1, 0)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, lmcs, kind, hc, vc) in CASES:
        refs0, refs1, pocs, cmds, luts, prof, sc = helpers_mc.rpr_case(po, w, h, seed, log2, lmcs, kind, hc, vc)
        dst = helpers_mc.blank(w, h)
        po.mc_picture_rpr(dst, [refs0, refs1], pocs, cmds, sc, log2, impl="ref", lmcs=luts, prof=prof, simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, lmcs=lmcs, kind=kind, hor_col=hc, ver_col=vc, nb_cmds=len(cmds),
                        input_md5=helpers.md5(cmds, prof, *refs0[0], *refs1[-1]), md5=[helpers.md5(p) for p in dst]))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "mc_rpr_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
