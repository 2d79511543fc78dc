"""The oracle's rendering of one picture of bench.py's workload: the chain of CPU ports (each pinned to the compiled reference by
its own test file) in the reference's order.  TEST INFRASTRUCTURE (checker only): used by tests/ and by bench.py's untimed
verification of its end-to-end leg."""
import ctypes as C

import numpy as np


def oracle_picture(po, wl, p, stop_after=None, impl="port"):
    """the oracle's rendering of picture p of workload wl (CU-order reconstruction, then the in-loop filters).
    impl="ref": motion compensation, deblocking, SAO and ALF run through the COMPILED REFERENCE's own slots (oracle/_ref, SIMD table);
    the CU-order composition of intra prediction and residual stays the port (pinned slot by slot in test_recon / test_intra)."""
    from oracle import pyoracle_recon as pr, pyoracle_intra as pi, pyoracle_df as pdf
    h, w = wl["refs"][0][0][0].shape
    luts = po.LmcsLuts()
    C.memmove(C.byref(luts), C.byref(wl["lmcs"]), C.sizeof(luts))
    t = po.AlfTables()
    C.memmove(C.byref(t), C.byref(p["tab"]), C.sizeof(t))
    rec = [np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16)]
    if len(p["cmds"]):
        po.mc_picture(rec, wl["refs"], wl["pocs"], p["cmds"].astype(po.McCmd), 7, impl=impl, lmcs=luts, prof=p["prof"])
    pr.recon_picture(rec, p["icmds_all"], p["tus_all"], p["coeffs"], 7, ict_type=1, anchors=p["anchors"], luts=luts,
                     lmcs_data=po.make_lmcs_data(wl["lmcs_dict"]), isp=pi.isp_info_array(p["isp_words"], p["isp_tus"]))
    if stop_after == "recon":
        return rec
    po.port().ovo_lmcs_reshape_plane(rec[0].ctypes.data_as(po.u16p), w, w, h, luts.np("bwd_lut").ctypes.data_as(po.u16p), 10)
    if impl == "ref":
        pdf.dbf_picture_ref(rec, p["dbf_blks"], 7, simd=1)
    else:
        pdf.dbf_picture_port(rec, p["dbf"])
    s = po.sao_picture(rec, p["sao"].astype(po.SaoCtu), 7, impl=impl)
    return po.alf_picture(s, t, p["ctu"], 7, impl=impl)
