"""Generates tests/golden/isp_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C): MD5 of the planes after the
reference's own tmp.recon_isp_subtree_{v,h} slots (prediction + residual of every partition) and intra_pred* slots ran over
seeded single-tree intra pictures, and of the availability words the reference held for the ISP partitions.
Run in the build container:  python tests/golden/gen_golden_isp.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po, pyoracle_intra as pi  # noqa: E402
import helpers  # noqa: E402
import helpers_intra  # noqa: E402

CASES = [(128, 128, 971, 7), (416, 240, 972, 7), (832, 480, 973, 7), (264, 136, 974, 6), (200, 120, 975, 5)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2) in CASES:
        planes, cmds, words, tus, arena = helpers_intra.isp_case(w, h, seed, log2)
        inp = helpers.md5(cmds, tus, arena, planes[0])
        rc = cmds.copy()
        rc["avail_cols"] = 0; rc["avail_rows"] = 0
        isp = pi.isp_info_array(words * 0 if False else words, tus)
        isp["common"] = 0; isp["part"] = 0
        pi.intra_picture_isp(planes, rc, isp, arena, log2, impl="ref", simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, nb_cmds=len(cmds), nb_isp=len(words), input_md5=inp, md5=helpers.md5(*planes),
                        avail_md5=helpers.md5(rc["avail_cols"], rc["avail_rows"], isp["common"], isp["part"])))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "isp_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
