"""Generates tests/golden/tr_golden.json from the UNMODIFIED reference (oracle/_ref, scalar C templates =
arithmetic of record): MD5 of the planes after tmp.rcn_tu_st / tmp.rcn_tu_c added the residual of seeded TU lists.
Run in the build container:  python tests/golden/gen_golden_tr.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po, pyoracle_tr as pt  # noqa: E402
import helpers  # noqa: E402
import helpers_tr  # noqa: E402

# (w, h, seed, log2_ctu, ict_type, sh_ts_disabled, dep_quant, legal, p_bdpcm); with p_bdpcm the picture is mostly intra so that
# BDPCM coding units (rcn_bdpcm_tb) are frequent
CASES = [(416, 240, 501, 7, 0, 0, 1, False, 0), (416, 240, 502, 7, 1, 0, 1, True, 0), (832, 480, 503, 7, 3, 1, 0, False, 0),
         (256, 128, 504, 6, 2, 0, 1, True, 0), (136, 72, 505, 5, 1, 1, 1, False, 0), (1920, 1080, 506, 7, 1, 0, 1, True, 0),
         (416, 240, 507, 7, 1, 0, 1, True, 0.8), (832, 480, 508, 7, 2, 1, 1, False, 0.8), (264, 136, 509, 6, 0, 1, 0, True, 0.8)]


def main():
    assert po.ref_available()
    out = []
    for (w, h, seed, log2, ict, tsd, dq, legal, pb) in CASES:
        planes, cmds, arena = helpers_tr.tr_case(w, h, seed, log2, ict, tsd, dq, legal, **(dict(p_intra=0.6, p_bdpcm=pb) if pb else {}))
        inp = helpers.md5(cmds, arena, *planes)
        pt.residual_picture(planes, cmds, arena, log2, impl="ref", dep_quant=dq, sh_ts_disabled=tsd, ict_type=ict, simd=0)
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, ict_type=ict, sh_ts_disabled=tsd, dep_quant=dq, legal=legal,
                        nb_cmds=len(cmds), p_bdpcm=pb, nb_bdpcm=int((cmds["bdpcm"] != 0).sum()), input_md5=inp, md5=[helpers.md5(p) for p in planes]))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "tr_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
