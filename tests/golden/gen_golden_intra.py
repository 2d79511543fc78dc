"""Generates tests/golden/intra_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C): MD5 of the luma plane
after rcn_funcs.intra_pred (and intra_pred_c for the chroma cases) ran over seeded intra blocks in decode order, and of the availability words the reference's
progress bit-field held at each call.  Run in the build container:  python tests/golden/gen_golden_intra.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po, pyoracle_intra as pi  # noqa: E402
import helpers  # noqa: E402
import helpers_intra  # noqa: E402

CASES = [(128, 128, 901, 7, 0), (416, 240, 902, 7, 0), (832, 480, 903, 7, 0), (264, 136, 904, 6, 0), (200, 120, 905, 5, 0), (1920, 1080, 906, 7, 0),
         # luma + Cb / Cr (rcn_funcs.intra_pred_c): md5 over the three planes
         (128, 128, 911, 7, 1), (416, 240, 912, 7, 1), (264, 136, 914, 6, 1), (200, 120, 915, 5, 1), (1920, 1080, 916, 7, 1),
         # + CIIP blocks (chroma = 2)
         (416, 240, 922, 7, 2), (832, 480, 923, 7, 2),
         # + multi reference line blocks (chroma = 3)
         (416, 240, 932, 7, 3), (264, 136, 934, 6, 3),
         # + CCLM / MDLM chroma (4: 6-tap luma down-sampling, 5: vertically collocated chroma)
         (416, 240, 942, 7, 4), (832, 480, 943, 7, 4), (200, 120, 945, 5, 4), (416, 240, 952, 7, 5), (264, 136, 954, 6, 5),
         # + MIP luma blocks (chroma = 6)
         (416, 240, 962, 7, 6), (832, 480, 963, 7, 6), (200, 120, 965, 5, 6)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, chroma) in CASES:
        planes, cmds = helpers_intra.intra_case(w, h, seed, log2, chroma=bool(chroma), p_ciip=0.4 if chroma in (2, 3) else 0.0,
                                                p_mrl=0.4 if chroma == 3 else 0.0, p_cclm=0.4 if chroma in (4, 5) else 0.0,
                                                collocated=chroma == 5, p_mip=0.4 if chroma == 6 else 0.0)
        inp = helpers.md5(cmds, planes[0])
        rc = cmds.copy()
        rc["avail_cols"] = 0; rc["avail_rows"] = 0
        pi.intra_picture(planes, rc, log2, impl="ref", simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, chroma=chroma, nb_cmds=len(cmds), input_md5=inp,
                        md5=helpers.md5(*planes) if chroma else helpers.md5(planes[0]),
                        avail_md5=helpers.md5(rc["avail_cols"], rc["avail_rows"])))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "intra_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
