"""Generates tests/golden/mc_gpm_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C): MD5 of the prediction
planes of seeded PU lists in which 60 % of the eligible CUs use geometric partitioning (rcn_gpm_b).
Run in the build container:  python tests/golden/gen_golden_gpm.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
import helpers  # noqa: E402
import helpers_mc  # noqa: E402

CASES = [(416, 240, 801, 7, False, 3), (832, 480, 802, 7, True, 3), (264, 136, 804, 6, True, 3), (1920, 1080, 806, 7, True, 3)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, lmcs, min_log2) in CASES:
        refs0, refs1, pocs, cmds, luts, prof = helpers_mc.mc_tools_case(po, w, h, seed, log2, lmcs, min_log2, p_gpm=0.6)
        dst = helpers_mc.blank(w, h)
        po.mc_picture(dst, [refs0, refs1], pocs, cmds, log2, impl="ref", lmcs=luts, prof=prof, simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, lmcs=lmcs, min_log2=min_log2, nb_cmds=len(cmds),
                        input_md5=helpers.md5(cmds, prof, *refs0[0], *refs1[-1]), md5=[helpers.md5(p) for p in dst]))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "mc_gpm_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
