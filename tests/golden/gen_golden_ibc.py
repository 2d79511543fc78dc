"""Generates tests/golden/ibc_golden.json from the UNMODIFIED reference (oracle/_ref): MD5 of the planes after the reference's
own rcn_ibc_l / rcn_ibc_c slots (working in a CTU ring buffer, see oracle/ref_harness_intra.c) and intra_pred* slots ran over
seeded single-tree pictures in which a third of the coding units are intra block copies.
Run in the build container:  python tests/golden/gen_golden_ibc.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po, pyoracle_intra as pi  # noqa: E402
import helpers  # noqa: E402
import helpers_intra  # noqa: E402

# (w, h, seed, log2_ctu, chroma)
CASES = [(128, 128, 981, 7, True), (416, 240, 982, 7, True), (832, 480, 983, 7, True), (264, 136, 984, 6, True), (200, 120, 985, 5, False)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, chroma) in CASES:
        planes, cmds = helpers_intra.ibc_case(w, h, seed, log2, chroma)
        inp = helpers.md5(cmds, *planes)
        pi.intra_picture(planes, cmds.copy(), log2, impl="ref", simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, chroma=chroma, nb_cmds=len(cmds), nb_ibc=int(((cmds["flags"] & 128) != 0).sum()),
                        input_md5=inp, md5=helpers.md5(*planes)))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "ibc_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
