"""Generates tests/golden/inloop_golden.json and inloop_tiny.npz from the UNMODIFIED reference
(oracle/_ref, built from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/gen_golden.py

Golden source = the reference's scalar C templates (rcn_init_*_functions_10), the arithmetic of record;
where the case is a CTC-like geometry the SSE/AVX2 table is asserted to give the same bytes.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
import helpers  # noqa: E402

CASES = [  # (w, h, seed, log2_ctu, kind, check_simd)
    (416, 240, 101, 7, "natural", True),
    (416, 240, 102, 7, "noise", True),
    (1920, 1080, 103, 7, "natural", True),
    (832, 480, 104, 6, "natural", False),
    (264, 136, 105, 7, "natural", False),
    (200, 120, 106, 5, "natural", False),
    (64, 64, 107, 7, "noise", False),
]


def main():
    assert po.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = []
    for (w, h, seed, log2, kind, check_simd) in CASES:
        planes, tab, ctu, _ = helpers.alf_case(po, w, h, seed, log2, kind)
        alf = po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=0)
        splanes, sao = helpers.sao_case(po, w, h, seed, log2, kind)
        so = po.sao_picture(splanes, sao, log2, impl="ref", simd=0)
        if check_simd:
            alf_s = po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=1)
            so_s = po.sao_picture(splanes, sao, log2, impl="ref", simd=1)
            helpers.assert_planes_equal(alf, alf_s, "ALF C vs SIMD")
            helpers.assert_planes_equal(so, so_s, "SAO C vs SIMD")
        out.append(dict(w=w, h=h, seed=seed, log2_ctu=log2, kind=kind,
                        input_md5=helpers.md5(*planes),
                        alf_param_md5=helpers.md5(tab.np("luma_coeff"), tab.np("luma_clip"), tab.np("chroma_coeff"),
                                                  tab.np("chroma_clip"), tab.np("cc_coeff"), helpers.alf_ctu_rows(ctu),
                                                  ctu["cc_cb"], ctu["cc_cr"]),
                        sao_param_md5=helpers.md5(sao),
                        alf_md5=[helpers.md5(p) for p in alf], sao_md5=[helpers.md5(p) for p in so]))
        print(out[-1])
    json.dump(out, open(os.path.join(HERE, "inloop_golden.json"), "w"), indent=1)

    # one tiny case stored in full (inputs + outputs), independent of the numpy generator
    w, h, seed, log2 = 72, 40, 201, 5
    planes, tab, ctu, _ = helpers.alf_case(po, w, h, seed, log2, "natural")
    alf = po.alf_picture(planes, tab, ctu, log2, impl="ref", simd=0)
    splanes, sao = helpers.sao_case(po, w, h, seed, log2, "natural")
    so = po.sao_picture(splanes, sao, log2, impl="ref", simd=0)
    np.savez_compressed(os.path.join(HERE, "inloop_tiny.npz"), w=w, h=h, log2_ctu=log2,
                        y=planes[0], cb=planes[1], cr=planes[2],
                        luma_coeff=tab.np("luma_coeff"), luma_clip=tab.np("luma_clip"),
                        chroma_coeff=tab.np("chroma_coeff"), chroma_clip=tab.np("chroma_clip"),
                        cc_coeff=tab.np("cc_coeff"), ctu_rows=helpers.alf_ctu_rows(ctu),
                        cc_cb=ctu["cc_cb"], cc_cr=ctu["cc_cr"],
                        alf_y=alf[0], alf_cb=alf[1], alf_cr=alf[2],
                        sao_params=sao, sao_y=so[0], sao_cb=so[1], sao_cr=so[2])


if __name__ == "__main__":
    main()
