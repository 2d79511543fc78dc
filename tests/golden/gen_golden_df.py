"""Generates tests/golden/df_golden.json from the UNMODIFIED reference (oracle/_ref): MD5 of the planes after
df.rcn_dbf_ctu ran over every CTU of seeded pictures (state built with the reference's own dbf_utils.h helpers).
Run in the build container:  python tests/golden/gen_golden_df.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po, pyoracle_df as pd  # noqa: E402
import helpers  # noqa: E402
import helpers_df  # noqa: E402

# (w, h, seed, log2_ctu, kind, beta_offset, tc_offset)
CASES = [(416, 240, 601, 7, "blocky", 0, 0), (416, 240, 602, 7, "natural", 2, -2), (832, 480, 603, 7, "blocky", 0, 0),
         (264, 136, 604, 6, "blocky", -4, 4), (200, 120, 605, 5, "natural", 0, 0), (1920, 1080, 606, 7, "blocky", 0, 0)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, kind, bo, to) in CASES:
        planes, blks, maps = helpers_df.dbf_case(w, h, seed, log2, kind, bo, to)
        inp = helpers.md5(blks, maps["ver"], maps["hor"], maps["qp"], *planes)
        pd.dbf_picture_ref(planes, blks, log2, bo, to, simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, kind=kind, beta_offset=bo, tc_offset=to, nb_blks=len(blks),
                        input_md5=inp, md5=[helpers.md5(p) for p in planes]))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "df_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
