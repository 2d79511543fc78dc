#!/usr/bin/env python
"""bench.py -- frames/sec of the rcn_* reconstruction hot path on synthetic 4K/8K 10-bit pictures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 4k|8k]

One "step" = one pass of the hot path over a batch of `pictures_per_step` synthetic inter pictures per GPU
(frame-parallel sharding: every rank owns whole pictures); value = pictures/s over all
ranks with inputs resident in HBM (the pictures of a step spread over --dev-streams streams); e2e = the same
through the host-buffer C-ABI calls: every step the decoder state of every picture (PU / TU / intra command
lists, coefficients, deblocking maps, SAO / ALF parameters: 15 MB per 4K picture) is committed from pinned
host memory (host-side validation + work lists + H2D) and the output picture is copied back (D2H, 25 MB),
all inside the timed region, one host thread per picture like the reference's frame threads, two picture
sets in flight.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# A frame-parallel decoder keeps dozens of streams busy (8 pictures in flight, each forking its independent kernels onto a side
# stream): with the default of 8 hardware work queues they alias and serialise each other (measured: 1935 against 1974 pictures/s).
# Must be set before the CUDA context exists; an application embedding the library does the same.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZES = {"1080p": (1920, 1080), "4k": (3840, 2160), "8k": (7680, 4320)}
STAGES_IMPLEMENTED = ["recon[mc(uni/bi/bcw, bdof, dmvr, gpm, affine 4x4 + prof)+lmcs_fwd, then per plane: residual(dequant/lfnst/idct/jcbcr) of inter CUs, "
                      "dataflow intra(planar/dc/angular/mrl/mip/isp/cclm/ciip)+residual, lmcs chroma scales from reconstructed luma]",
                      "lmcs_bwd", "dbf(v+h,luma+chroma)", "sao", "alf+ccalf"]   # widened as §8 rows land


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(w, h, n_pics, seed0=5000, all_intra=False, tables="product"):
    """Shared reference pictures + n_pics synthetic inter pictures: PU list, SAO and ALF decoder state
    (SURVEY.md §8(d) generators).  all_intra: every coding unit is intra and there is no PU list (an I picture).
    tables: who builds the LMCS LUTs and the ALF coefficient tables from the APS data -- "product" = the C ABI's host-side
    builders (our arm), "oracle" = the compiled reference / its port (the reference arm: that process never loads the product
    library; the two builders are checked against each other by the tests)."""
    from openvvc_b200 import api, synth, synth_tu, synth_dbf, synth_recon
    refs = [[synth.planes(w, h, seed0 + 900 + i) for i in range(2)], [synth.planes(w, h, seed0 + 950 + i) for i in range(2)]]
    pocs = [[8, 16], [16, 24]]
    if tables == "oracle":
        from oracle import pyoracle as po_t
        impl_t = "ref" if po_t.ref_available() else "port"
        lmcs = po_t.lmcs_build_luts(synth.lmcs_data(seed0), impl=impl_t)
    else:
        lmcs = api.lmcs_build_luts(synth.lmcs_data(seed0))
    pics = []
    for i in range(n_pics):
        ap = synth.alf_params(w, h, seed0 + 100 + i, 7)
        if tables == "oracle":
            tab = po_t.alf_build_tables(ap["luma_aps"], ap["chroma_aps"], impl=impl_t)
        else:
            tab = api.alf_tables_from_aps(ap["luma_aps"], ap["chroma_aps"])
        tab.np("cc_coeff")[:] = ap["cc_coeff"]
        sao = synth.sao_params(w, h, seed0 + 200 + i, 7, api.SAO_CTU)
        cmds, prof = synth.mc_commands_tools(w, h, seed0 + 300 + i, 7, p_bdof=0.25, p_affine=0.15, p_dmvr=0.15, p_gpm=0.1)
        cmds = cmds.astype(api.MC_CMD)
        if all_intra:
            cmds, prof = cmds[:0], prof[:0]
        # one CU partition for the residual and the intra side: 10 % intra CUs (MRL / MIP / CCLM among them), 4 % CIIP, every TU
        # takes its LMCS chroma scale from the anchor in force (derived on the GPU from reconstructed luma)
        icmds, tus, coeffs, isp_words, isp_tus = synth_recon.recon_case(w, h, seed0 + 400 + i, 7, min_log2=3, p_inter=0.0 if all_intra else 0.86, p_ciip=0.3, p_mrl=0.15,
                                                                        p_cclm=0.3, ict_type=1, p_cbf=0.4, p_mip=0.1, p_isp=0.1, p_bdpcm=0.3)
        anchors, tus = synth_recon.lmcs_anchors(icmds, tus, 7)
        icmds_all, tus_all = icmds, tus                     # one entry per coding unit, decode order (what the CU-order oracle walks)
        icmds = np.ascontiguousarray(icmds[icmds["flags"] != 0])
        tus = np.ascontiguousarray(np.concatenate([tus, isp_tus.reshape(-1)[isp_tus.reshape(-1)["cbf_mask"] != 0]]))   # + the ISP partitions' blocks
        blks = synth_dbf.dbf_blocks(w, h, seed0 + 500 + i, 7, min_log2=3)
        dbf = synth_dbf.dbf_maps(w, h, blks, 7)
        pics.append(dict(tab=tab, ctu=ap["ctu"], alf_ctu=api.alf_ctu_array(ap["ctu"]), sao=sao, cmds=cmds, prof=prof, tus=tus, coeffs=coeffs,
                         dbf_blks=blks, dbf=dbf, icmds=icmds, anchors=anchors, isp_words=isp_words, isp_tus=isp_tus,
                         icmds_all=icmds_all, tus_all=tus_all))
    return dict(refs=refs, pocs=pocs, lmcs=lmcs, lmcs_data=api.LmcsData.from_dict(synth.lmcs_data(seed0)), lmcs_dict=synth.lmcs_data(seed0), pics=pics)


def make_dev(api, c, p, w, h, upload=True):
    """device objects of one picture (C-ABI handles); upload=True: with its decoder state uploaded (device-resident leg),
    False: empty, to be filled through stage_picture / commit_picture (end-to-end leg)"""
    d = dict(rec=c.picture(w, h), mid=c.picture(w, h), dst=c.picture(w, h),
             alf=api.AlfParams(c, w, h, 7), sao=api.SaoParams(c, w, h, 7),
             mc=api.McBatch(c, len(p["cmds"]), api.mc_nb_items(p["cmds"])) if len(p["cmds"]) else None,
             tu=api.TuBatch(c, len(p["tus"]), len(p["coeffs"])), dbf=api.DbfParams(c, w, h, 7),
             intra=api.IntraBatch(c, max(1, len(p["icmds"]))))
    if not upload:
        return d
    d["alf"].upload(p["tab"], p["alf_ctu"], (p["ctu"]["cc_cb"], p["ctu"]["cc_cr"]))
    d["sao"].upload(p["sao"])
    if d["mc"]:
        d["mc"].upload(p["cmds"], prof=p["prof"])
    d["tu"].set_anchors(p["anchors"])
    d["tu"].upload(p["tus"], p["coeffs"], 1, 0, 1)
    d["intra"].upload(p["icmds"], w, h, 7, isp=p["isp_words"])
    d["dbf"].upload(p["dbf"])
    return d


def run_stages(api, c, d, ref_pics, pocs, lmcs):
    """one pass of the hot path over one picture: MC + residual + intra (reconstruction order), LMCS backward reshape, DBF, SAO, ALF"""
    api.recon_picture(c, d["rec"], tu=d["tu"], intra=d["intra"], mc=d["mc"], ref_lists=ref_pics, pocs=pocs, lmcs=lmcs)
    lmcs.reshape(d["rec"], forward=False, ctx=c)
    d["dbf"].run(d["rec"], ctx=c)
    d["sao"].run(d["mid"], d["rec"], ctx=c)
    d["alf"].run(d["dst"], d["mid"], ctx=c)


def stage_picture(d, p):
    """what the recorder does while the CPU parses picture p: writes its PU / TU commands, coefficients and deblocking maps
    straight into the PINNED staging buffers of the picture's objects (ovgpu_*_staging)"""
    if d["mc"]:
        mc_c, mc_p = d["mc"].staging()
        mc_c[:len(p["cmds"])] = p["cmds"]; mc_p[:len(p["prof"])] = p["prof"]
    hc, hk = d["tu"].staging()
    hc[:len(p["tus"])] = p["tus"]; hk[:len(p["coeffs"])] = p["coeffs"]
    st = d["dbf"].staging()
    for k in range(9):
        st["ver"][k][:] = p["dbf"]["ver"][k]; st["hor"][k][:] = p["dbf"]["hor"][k]
    for k in range(3):
        st["qp"][k][:] = p["dbf"]["qp"][k]


def commit_picture(api, c, d, p, w, h, tt=None):
    """end of the parse of picture p on context c: every *_commit (validation + work lists + asynchronous H2D out of the pinned
    staging buffers) and the small per-CTU tables"""
    if d["mc"]:
        d["mc"].commit(len(p["cmds"]), len(p["prof"]), ctx=c)
    if tt is not None: tt.append(time.perf_counter())
    d["tu"].set_anchors(p["anchors"], ctx=c)
    d["tu"].commit(len(p["tus"]), len(p["coeffs"]), 1, 0, 1, ctx=c)
    if tt is not None: tt.append(time.perf_counter())
    d["intra"].upload(p["icmds"], w, h, 7, ctx=c, isp=p["isp_words"])
    if tt is not None: tt.append(time.perf_counter())
    d["dbf"].commit(0, 0, ctx=c)
    if tt is not None: tt.append(time.perf_counter())
    d["sao"].upload(p["sao"], ctx=c)
    d["alf"].upload(p["tab"], p["alf_ctu"], (p["ctu"]["cc_cb"], p["ctu"]["cc_cr"]), ctx=c)
    if tt is not None: tt.append(time.perf_counter())


STAGE_LAUNCHES = {"TrK": 4, "Lmcs": 2, "ovd_recon_kernel": 2, "ReconPrepK": 2}      # launches per picture of kernels that run once per plane
STAGE_KERNELS = {"mc": ("McK", "McBigK", "BdofK", "DmvrK", "ProfK", "McSmallCK", "McPrep", "McGuardK"), "residual": ("TrK",), "intra": ("ovd_recon_kernel", "ReconPrepK"),
                 "lmcs": ("Lmcs",), "dbf_v": ("DbfK<0>",), "dbf_h": ("DbfK<1>",), "sao": ("SaoK",), "alf_luma": ("AlfLumaK",),
                 "alf_chroma": ("AlfChromaK",)}


def ncu_traffic(stage):
    """DRAM bytes (read + write) per picture of a stage's kernels, from the newest committed `ncu --set full` summary under
    profiles/ (tools/ncu_summary.py over a capture of tools/ncu_picture.py: exactly ONE picture between cudaProfilerStart / Stop,
    L2 flushed before it; older summaries without a `pictures` field are normalised per kernel by its own launch count).
    None when there is no summary."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[0-9]_ncu_full_all_stages_v*.json")),
                   key=lambda f: (int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)), int(re.search(r"_v(\d+)", f).group(1))))
    if not files:
        return None, None
    doc = json.load(open(files[-1]))
    launches = doc["launches"]
    total, found = 0.0, False
    for k in STAGE_KERNELS[stage]:
        sel = [l for l in launches if k in l["kernel"]]
        if not sel:
            continue
        found = True
        # pictures this kernel's launches stand for: the capture's picture count, or (older captures whose window straddled
        # pictures) the kernel's OWN launch count over its launches per picture
        n_pics = doc.get("pictures") or max(1.0, len(sel) / STAGE_LAUNCHES.get(k, 1))
        total += sum(l.get("dram_read_MB", 0.0) + l.get("dram_write_MB", 0.0) for l in sel) / n_pics
    if not found:
        return None, os.path.basename(files[-1])
    return 1e6 * total, os.path.basename(files[-1])


def residual_bytes(tus):
    """Algorithmic bytes of the residual stage: coded coefficient sub-blocks (32 B each) + read-modify-write of the
    samples of every coded transform block."""
    area = (1 << tus["log2_w"].astype(np.int64)) * (1 << tus["log2_h"].astype(np.int64))
    cbf = tus["cbf_mask"]
    luma = (cbf & 0x10) != 0
    joint = (cbf & 0x08) != 0
    n_c = np.where(joint, 2, ((cbf & 2) != 0).astype(np.int64) + ((cbf & 1) != 0))        # chroma planes touched
    n_ctb = np.where(joint, 1, n_c)                                                       # chroma blocks coded
    rmw = 4 * (area * luma + (area // 4) * n_c)
    sb = np.zeros(len(tus), np.int64)
    for k, used in ((2, luma), (0, n_ctb >= 1), (1, (n_ctb == 2))):
        bits = np.array([bin(int(v)).count("1") for v in tus["sig_sb_map"][:, k]], np.int64)
        sb += np.where(used, np.maximum(bits, 1), 0)
    return float(rmw.sum() + 32 * sb.sum())


def mc_dirs_per_sample(cmds):
    """average number of prediction directions per luma sample (the d of the (d + 1) * P MC floor)"""
    luma = (cmds["flags"] & 1) != 0
    area = (1 << cmds["log2_w"].astype(np.int64)) * (1 << cmds["log2_h"].astype(np.int64))
    ident = (cmds["tool"] == 0) & (cmds["mv0_x"] == cmds["mv1_x"]) & (cmds["mv0_y"] == cmds["mv1_y"]) & (cmds["ref_idx0"] == 1) & (cmds["ref_idx1"] == 0)
    nref = np.where((cmds["inter_dir"] == 3) & ~ident, 2, 1)
    return float((area * nref)[luma].sum() / area[luma].sum())


# ------------------------------------------------------------------------------------------------ reference arm
def reference_step(po, wl, threads):
    """The reference's own CPU implementation (oracle/_ref: rcn_mcp_b, rcn_sao_filter_line, rcn_alf_filter_line
    through the SSE4.1/AVX2 function table) of the same stages, one picture per host thread (the reference's
    frame parallelism, ovdec.h:47).  Falls back to the C port only if oracle/_ref was not built."""
    from oracle import pyoracle_tr as pt, pyoracle_df as pdf, pyoracle_intra as pi
    impl = "ref" if po.ref_available() else "port"
    h, w = wl["refs"][0][0][0].shape
    luts = po.LmcsLuts()
    C.memmove(C.byref(luts), C.byref(wl["lmcs"]), C.sizeof(luts))

    def one(p):
        t = po.AlfTables()
        C.memmove(C.byref(t), C.byref(p["tab"]), C.sizeof(t))
        rec = [np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16)]
        po.mc_picture(rec, wl["refs"], wl["pocs"], p["cmds"].astype(po.McCmd), 7, impl=impl, lmcs=luts, prof=p["prof"])
        pt.residual_picture(rec, p["tus"][(p["tus"]["flags"] & 16) == 0], p["coeffs"], 7, impl=impl, dep_quant=1, sh_ts_disabled=0, ict_type=1, simd=1)
        # same blocks through the reference's own slots in CU order (intra_pred*, mip, recon_isp_subtree_*); the ISP partitions'
        # transform blocks are part of that call, so they are left out of the residual pass above
        pi.intra_picture_isp(rec, p["icmds"].copy(), pi.isp_info_array(p["isp_words"], p["isp_tus"]), p["coeffs"], 7, impl=impl, simd=1)
        po.lmcs_reshape_backward(rec[0], luts, 7, impl=impl)           # the reference's own lmcs_reshape_backward slot, CTU by CTU
        if impl == "ref":
            pdf.dbf_picture_ref(rec, p["dbf_blks"], 7, simd=1)
        else:
            pdf.dbf_picture_port(rec, p["dbf"])
        s = po.sao_picture(rec, p["sao"].astype(po.SaoCtu), 7, impl=impl)
        po.alf_picture(s, t, p["ctu"], 7, impl=impl)

    ths = [threading.Thread(target=one, args=(wl["pics"][i % len(wl["pics"])],)) for i in range(threads)]     # read-only inputs: pictures may repeat
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, len(ths)


def run_reference(args, w, h):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    wl = make_workload(w, h, min(cores, 8), tables="oracle")
    threads = cores                      # every host thread the box has: one picture per thread (the reference's frame threads)
    for _ in range(max(1, min(args.warmup, 1))):
        reference_step(po, wl, threads)
    tot, n = 0.0, 0
    for _ in range(args.steps):
        dt, k = reference_step(po, wl, threads)
        tot += dt; n += k
    fps = n / tot
    kind = "reference" if po.ref_available() else "port"
    line = {"impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/int32", "data": "synthetic",
            "config": workload_config(w, h, threads),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": "%d pictures per step, one per host thread, %d steps" % (threads, args.steps)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # this arm is the reference alone: the product library must not even be mapped into the process
    line["product_library_loaded"] = "libovvc_b200" in open("/proc/self/maps").read()
    print(json.dumps(line))


def workload_config(w, h, pics_per_step, dev_streams=1, dag=0):
    par = ("random-access GOP32 as a dependency graph (a step = one GOP of 32 pictures, 1-1-2-4-8-16 wide): picture k of the decoding order on "
           "rank k mod N, every picture predicted from its two reference pictures of the hierarchy, a finished reference sent to exactly the "
           "ranks whose pictures name it (ovgpu_picture_publish / consume: grouped ncclSend / ncclRecv of the 3 planes, ordered by events); "
           "the e2e number is measured on independent pictures per step with one picture per step handed to all ranks"
           if dag else
           "frame-parallel: every rank reconstructs its own pictures; per step one finished picture is handed to all ranks as a reference "
           "(device copy + ncclBroadcast of 3 planes)")
    return {"workload": "%dx%d 10-bit 4:2:0 CTU128 synthetic RA inter picture, rcn stages in the timed path: %s "
                        "(intra block copy: off, as for natural video in the common test conditions)" % (w, h, " -> ".join(STAGES_IMPLEMENTED)),
            "pictures_per_step": pics_per_step, "streams_per_gpu": dev_streams, "parallelism": par,
            "l2": "inputs larger than L2: every picture in flight has its own recon + mid planes and writes its own DPB slot (72 slots), "
                  "%d pictures per step" % pics_per_step}


def run_dag(api, ovdist, torch, dist, args, w, h, rank, world, local, stream, wl, lmcs, d_streams, d_ctx, barrier):
    """`value` leg: warmup + steps GOPs of a random-access stream, reconstructed as the dependency graph it is.
    Returns (milliseconds of the timed GOPs on the device, max over nothing -- the caller reduces over ranks --, info)."""
    import helpers
    n_ds = len(d_streams)
    nb_slots = 168                                  # DPB copies: the issue order keeps 3-4 GOPs in flight, a slot is reused 5 GOPs later
    W, K, R = max(1, args.warmup), args.steps, max(1, args.repeats)
    sched = ovdist.GopSchedule(W + R * K, world)
    pics = wl["pics"]
    mk_ctx = d_ctx[0]
    slots = [mk_ctx.picture(w, h) for _ in range(nb_slots)]
    slots[0].upload(wl["refs"][0][0])               # POC 0: given, resident on every rank
    dsets = [make_dev(api, d_ctx[j], pics[j % len(pics)], w, h) for j in range(n_ds)]
    for j in range(n_ds):
        d_ctx[j].sync()
    # The sends / receives of one communicator run in issue order on one stream (that is what keeps them matched between the
    # ranks), so a receive that waits for a picture deep in one GOP's dependency chain holds back everything queued behind it: in
    # decoding order that is one GOP per chain latency (7.2 ms on 8 GPUs).  The issue order (wave order, dist.py) interleaves the
    # chains of consecutive GOPs instead.  (More communicators -- lanes -- were tried and starve each other.)
    L = max(1, args.dag_lanes) if world > 1 else 1
    xstreams = [torch.cuda.Stream() for _ in range(L)]
    xctxs = [api.Ctx(local, st_.cuda_stream) for st_ in xstreams]
    xds = [None] * L
    for l in range(L if world > 1 else 0):
        uid = torch.zeros(api.Dist.ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(api.Dist.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        torch.cuda.synchronize()
        xds[l] = api.Dist(xctxs[l], world, rank, bytes(uid.cpu().numpy().tobytes()))

    def lane(poc):
        return ((poc - 1) // 32) % L
    ready, readers, nxt = {}, {}, [0]
    ev0 = torch.cuda.Event(); ev0.record(stream); ready[0] = ev0

    def wait_slot_free(s, poc):
        for ev in readers.pop(poc % nb_slots, []):
            s.wait_event(ev)                        # kernels / sends that still read the slot's previous picture

    def reconstruct(poc, parents):
        j = nxt[0]; nxt[0] = (j + 1) % n_ds
        s, c, d = d_streams[j], d_ctx[j], dsets[j]
        for p in set(parents):
            s.wait_event(ready[p])
        wait_slot_free(s, poc)
        ra, rb = slots[parents[0] % nb_slots], slots[parents[1] % nb_slots]
        run_stages(api, c, dict(d, dst=slots[poc % nb_slots]), [[ra, rb], [rb, ra]], [[parents[0], parents[1]], [parents[1], parents[0]]], lmcs)
        ev = torch.cuda.Event(); ev.record(s); ready[poc] = ev
        for p in set(parents):
            readers.setdefault(p % nb_slots, []).append(ev)

    def publish(poc, consumers):
        l = lane(poc)
        xstreams[l].wait_event(ready[poc])
        xds[l].publish(slots[poc % nb_slots], consumers, ctx=xctxs[l])
        ev = torch.cuda.Event(); ev.record(xstreams[l])
        readers.setdefault(poc % nb_slots, []).append(ev)

    def consume(poc, owner):
        l = lane(poc)
        wait_slot_free(xstreams[l], poc)
        xds[l].consume(slots[poc % nb_slots], owner, ctx=xctxs[l])
        ev = torch.cuda.Event(); ev.record(xstreams[l]); ready[poc] = ev

    def join():
        for s in list(d_streams[1:]) + xstreams:
            ev = torch.cuda.Event(); ev.record(s); stream.wait_event(ev)

    def traffic():
        return sum(x.traffic()[0] for x in xds if x)

    n0 = sum(c.launch_count() for c in d_ctx)
    order = sched.wave_order(args.dag_wave) if args.dag_wave else None
    ovdist.run_gop_stream(sched, rank, reconstruct, publish, consume, 0, 32 * W, order)
    join(); barrier()
    oracle_checked = 0
    if rank == 0 and W <= 3 and not os.environ.get("BENCH_SKIP_ORACLE_CHECK"):
        # the stream's first picture (POC 32 predicted from POC 0, owner rank 0, commands of picture 0) against the CPU oracle chain on
        # the same references: the graph plumbing (slots, reference lists, events) delivers what the oracle computes.  Untimed.
        import helpers_chain
        from oracle import pyoracle as po_chk
        p0 = wl["refs"][0][0]
        want = helpers_chain.oracle_picture(po_chk, dict(wl, refs=[[p0, p0], [p0, p0]], pocs=[[0, 0], [0, 0]]), pics[0])
        helpers.assert_planes_equal(slots[32 % nb_slots].download(), want, "bench.py: first picture of the GOP stream vs the oracle chain")
        oracle_checked = 1
    barrier()
    sent0 = traffic()
    ms_all, issue_ms = [], []
    for r_ in range(R):                             # R timed runs of exactly K steps (GOPs) each, back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in list(d_streams[1:]) + xstreams:
            s.wait_event(e0)
        t_issue = time.perf_counter()
        ovdist.run_gop_stream(sched, rank, reconstruct, publish, consume, 32 * (W + r_ * K), 32 * (W + (r_ + 1) * K), order)
        join()
        e1.record(stream)
        issue_ms.append((time.perf_counter() - t_issue) * 1e3)      # host time to ISSUE the run (no wait in it): against the device time below
        barrier()
        ms_all.append(e0.elapsed_time(e1))
    ms = ms_all
    sent = (traffic() - sent0) / R
    # every copy of a picture of the last GOP (owner and consumers) must be the same picture, and the stream's first picture
    # (POC 32 from POC 0, owner rank 0) must equal the CPU oracle chain on the same references
    last = [poc for poc, _ in sched.pictures[-32:] if sched.owner[poc] == rank or sched.needs(rank, poc)]
    md5 = {poc: helpers.md5(*slots[poc % nb_slots].download()) for poc in last}
    info = {"gops": K, "warmup_gops": W, "pictures": 32 * K, "repeats": R, "nvlink_bytes_per_step": None, "copies_identical": None, "oracle_checked": oracle_checked,
            "lanes": L, "launches": sum(c.launch_count() for c in d_ctx) - n0,
            # how long the one host thread of the rank takes to issue a timed run, per step: close to ms_per_step = the host, not the GPU, sets the pace
            "host_issue_ms_per_step": sorted(issue_ms)[len(issue_ms) // 2] / K}
    if world > 1:
        allmd5 = [None] * world
        dist.all_gather_object(allmd5, md5)
        t = torch.tensor([float(sent)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        info["nvlink_bytes_per_step"] = float(t[0]) / K
        if rank == 0:
            merged = {}
            for m in allmd5:
                for poc, dg in m.items():
                    merged.setdefault(poc, set()).add(dg)
            bad = [poc for poc, s_ in merged.items() if len(s_) != 1]
            if bad:
                raise SystemExit("bench.py: copies of pictures %s differ between ranks" % bad)
            info["copies_identical"] = len(merged)
    else:
        info["nvlink_bytes_per_step"] = 0.0
    return ms, info


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="4k", choices=sorted(SIZES))
    ap.add_argument("--pics", type=int, default=8, help="pictures per step per GPU")
    ap.add_argument("--dev-streams", type=int, default=8, help="streams the device-resident pictures of a step are spread over")
    ap.add_argument("--e2e-share", type=int, default=1, help="e2e contexts (host threads) that share one stream")
    ap.add_argument("--e2e-sets", type=int, default=2, help="picture sets in flight in the e2e leg (1 or 2)")
    ap.add_argument("--e2e-streams", type=int, default=8, help="independent contexts (streams) the e2e leg spreads pictures over")
    ap.add_argument("--dag", type=int, default=1, help="1: `value` = a random-access GOP32 stream reconstructed as a dependency graph, picture k on rank "
                    "k mod N, references handed to the ranks that need them (ovgpu_picture_publish / consume); 0: independent pictures per step")
    ap.add_argument("--repeats", type=int, default=3, help="timed runs of exactly --steps steps each; the median run is reported, all are listed")
    ap.add_argument("--dag-lanes", type=int, default=1, help="communicators / exchange streams the GOPs of the --dag stream alternate over "
                    "(measured on 8 GPUs: 4 lanes = 56 pictures/s, concurrent communicators starve each other; kept for the record)")
    ap.add_argument("--dag-wave", type=int, default=1, help="0: pictures are issued in decoding order; n: in wave order, a new GOP every n hierarchy levels (dist.py)")
    ap.add_argument("--e2e-dl-streams", type=int, default=0, help="1: the output copy of a picture runs on a stream of its own (measured: no gain)")
    args = ap.parse_args()
    w, h = SIZES[args.size]

    if args.impl == "reference":
        run_reference(args, w, h)
        return

    # Libraries (NCCL prints its version banner) may write to stdout; the contract is ONE JSON line there, so route
    # fd 1 to stderr for the duration of the run and keep the real stdout for the result line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from openvvc_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated (non-default) stream shared by torch events and the C-ABI context, so that
    # torch.cuda.Event timing brackets exactly the kernels the library launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = api.Ctx(local, stream.cuda_stream)

    P = args.pics
    wl = make_workload(w, h, P, seed0=5000 + 1000 * rank)
    pics = wl["pics"]
    pbytes = w * h * 2 * 3 // 2

    # device-resident state: reference pictures, per picture recon (prediction), mid (post-SAO), dst (post-ALF)
    # The DPB copies live in flat torch tensors wrapped as pictures (ovgpu_picture_wrap) so that the owner of a finished
    # picture can hand it to the other ranks with an NCCL broadcast (openvvc_b200/dist.py).
    from openvvc_b200 import dist as ovdist
    dev_name = "cuda:%d" % local
    ref_shared = [[ovdist.SharedPicture(w, h, dev_name) for _ in lst] for lst in wl["refs"]]
    ref_pics = [[ctx.wrap_picture(w, h, *sp.plane_ptrs()) for sp in lst] for lst in ref_shared]
    for lst, src in zip(ref_pics, wl["refs"]):
        for pic, planes in zip(lst, src):
            pic.upload(planes)
    lmcs = api.Lmcs(ctx, wl["lmcs"])
    lmcs.set_data(wl["lmcs_data"])
    dev = [make_dev(api, ctx, p, w, h) for p in pics]
    ctx.sync()

    # Frame parallelism on one GPU: the pictures of a step are independent (same temporal layer), so they are spread over
    # a few streams -- the latency-bound dataflow kernel of one picture then overlaps the throughput-bound kernels of the
    # others, like the reference's frame threads overlap on the CPU (ovdec.h:47).  Stream 0 is the timed stream: the other
    # streams start after an event recorded on it and it waits for their last kernels before the hand-off.
    n_ds = max(1, min(args.dev_streams, P))
    d_streams = [stream] + [torch.cuda.Stream() for _ in range(n_ds - 1)]
    d_ctx = [ctx] + [api.Ctx(local, st.cuda_stream) for st in d_streams[1:]]

    def run_picture(d, c=None):
        run_stages(api, c or ctx, d, ref_pics, wl["pocs"], lmcs)

    step_no = [0]

    def hand_off_reference(src=None):
        """the newest reconstructed picture of this step's owner rank becomes a reference on every rank: device copy
        into the owner's DPB slot, then ncclBroadcast of the 3 planes (P bytes) to the other ranks"""
        k = step_no[0]
        step_no[0] += 1
        slot, owner = k % 4, k % world
        if rank == owner:
            api._check(ctx.h, api.lib().ovgpu_picture_copy(ctx.h, ref_pics[slot // 2][slot % 2].p, (src or dev[0])["dst"].p), "picture_copy")
        ovdist.publish_reference(ref_shared[slot // 2][slot % 2], owner, world)

    def step_device(parallel=True, hand_off=True):
        if not parallel or n_ds == 1:
            for d in dev:
                run_picture(d)
        else:
            ev0 = torch.cuda.Event()
            ev0.record(stream)
            for st in d_streams[1:]:
                st.wait_event(ev0)
            for i, d in enumerate(dev):
                run_picture(d, d_ctx[i % n_ds])
            for st in d_streams[1:]:
                ev = torch.cuda.Event()
                ev.record(st)
                stream.wait_event(ev)
        if hand_off:
            hand_off_reference()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)

    # ---- value: device-resident throughput, CUDA events on the launching stream
    n0 = sum(c.launch_count() for c in d_ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dag_info = None
    if args.dag:
        ms_runs, dag_info = run_dag(api, ovdist, torch, dist, args, w, h, rank, world, local, stream, wl, lmcs, d_streams, d_ctx, barrier)
        pics_per_step = 32
    else:
        ms_runs = []
        for _ in range(max(1, args.repeats)):
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                step_device()
            e1.record(stream)
            barrier()
            ms_runs.append(e0.elapsed_time(e1))
        pics_per_step = P
    launches = (dag_info["launches"] if dag_info else sum(c.launch_count() for c in d_ctx) - n0) // max(1, args.repeats)

    # ---- per-kernel device time (CUDA events around each launch, same stream), separate pass
    stage_names = ("mc", "residual", "intra", "lmcs", "dbf_v", "dbf_h", "sao", "alf_luma", "alf_chroma")
    ctx.profile(True); ctx.profile_reset()
    nprof = max(2, min(args.steps, 5))
    for _ in range(nprof):
        step_device(parallel=False)          # stage by stage on one stream: per-kernel times without overlap
    prof = {s: ctx.profile_get(s) for s in stage_names}
    ctx.profile(False)

    # ---- e2e: decoder state in PINNED host memory -> H2D -> stages -> D2H of the output picture, every step.
    # Like the reference's frame threads (ovdec.h:47), pictures are spread over a few independent contexts (streams)
    # so the copies of one picture overlap the kernels of another.  The big inputs (coefficients, deblocking maps)
    # sit in the objects' pinned staging buffers -- where the north-star recorder writes them -- and are copied to
    # the device inside the timed region by commit(); the small per-CTU tables go through upload().
    # Two picture sets in flight (even / odd steps), each with its own contexts, streams, device objects and pinned output: the
    # H2D copies of step k+1 and the D2H copies of step k-1 then run under the kernels of step k instead of queueing behind
    # them in stream order -- 2 x n_ctx pictures in flight, like a decoder with that many frame threads.
    # host threads per rank: never more than this rank's share of the box's cores (an 8-GPU box with 32 cores runs 8 ranks: the
    # staging fences spin while they wait, so oversubscribed threads would take cores from the threads that have work)
    n_ctx = max(1, min(args.e2e_streams, P, max(1, (os.cpu_count() or 1) // world)))
    n_sets = max(1, min(2, args.e2e_sets))
    share = max(1, args.e2e_share)                 # contexts (host threads) per stream
    real_streams = [torch.cuda.Stream() for _ in range((n_sets * n_ctx + share - 1) // share)]
    e_streams = [real_streams[j // share] for j in range(n_sets * n_ctx)]
    e_ctx = [api.Ctx(local, st.cuda_stream) for st in e_streams]
    # the output copy of a picture runs on a stream of its own (like a decoder's output stage): the PCIe link out is as busy as the
    # SMs (25 MB per 4K picture as 16-bit samples), and queued on the picture's stream it would hold back the uploads of the next
    # picture that context handles
    dl_streams = [torch.cuda.Stream() for _ in e_streams] if args.e2e_dl_streams else e_streams
    dl_ctx = [api.Ctx(local, st.cuda_stream) for st in dl_streams] if args.e2e_dl_streams else e_ctx
    dl_done = [[None] * P for _ in range(n_sets)]
    e_dev = [dev] + [[make_dev(api, ctx, p, w, h) for p in pics] for _ in range(n_sets - 1)]
    ctx.sync()
    pinned_out = [[[torch.empty((hh, ww), dtype=torch.uint16).pin_memory() for (hh, ww) in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
                   for _ in pics] for _ in range(n_sets)]
    np_out = [[[t.numpy() for t in pin] for pin in half] for half in pinned_out]
    for half in e_dev:
        for p, d in zip(pics, half):
            stage_picture(d, p)

    def e2e_worker(t, ev_start, half):
        """one host thread per context, like the reference's frame threads (ovdec.h:47): stages, commits and launches its
        pictures; ctypes releases the GIL inside the C-ABI calls, so the host-side work of different pictures overlaps.
        Nothing here waits for the GPU: staging + H2D of a step run while the kernels of the previous step are still
        busy, and the D2H of a finished picture runs under the kernels of the next step (the staging fences of the C ABI keep a pinned buffer from being rewritten before its last copy left);
        only the kernels wait, on the device, for the previous step's reference hand-off (ev_start)."""
        torch.cuda.set_device(local)
        c, st = e_ctx[half * n_ctx + t], e_streams[half * n_ctx + t]
        evs = []
        tl = timeline is not None

        def mark(name, i):
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            timeline.append((e2e_no[0], i, name, e, time.perf_counter()))
        dc, dst_ = dl_ctx[half * n_ctx + t], dl_streams[half * n_ctx + t]
        for i in range(t, len(dev), n_ctx):
            p, d = pics[i], e_dev[half][i]
            tt = [time.perf_counter()]
            if tl: mark("h2d0", i)
            commit_picture(api, c, d, p, w, h, tt)
            if tl: mark("h2d1", i)
            st.wait_event(ev_start)
            if tl: mark("k0", i)
            api.recon_picture(c, d["rec"], tu=d["tu"], intra=d["intra"], mc=d["mc"], ref_lists=ref_pics, pocs=wl["pocs"], lmcs=lmcs); tt.append(time.perf_counter())
            lmcs.reshape(d["rec"], forward=False, ctx=c)
            d["dbf"].run(d["rec"], ctx=c)
            d["sao"].run(d["mid"], d["rec"], ctx=c)
            if args.e2e_dl_streams and dl_done[half][i] is not None:
                st.wait_event(dl_done[half][i])     # (on the device) ALF rewrites the planes the previous output copy of this set read
            d["alf"].run(d["dst"], d["mid"], ctx=c); tt.append(time.perf_counter())
            if host_prof is not None:
                host_prof.append([1e3 * (b - a) for a, b in zip(tt, tt[1:])])
            ev = torch.cuda.Event()
            ev.record(st)                           # the picture is final on the device: what the reference hand-off needs
            evs.append(ev)
            if tl: mark("k1", i)
            dst_.wait_event(ev)
            api._check(dc.h, api.lib().ovgpu_picture_download(dc.h, d["dst"].p, *api._host_planes(np_out[half][i])), "download")
            evd = torch.cuda.Event()
            evd.record(dst_)
            dl_done[half][i] = evd
            if tl:
                e = torch.cuda.Event(enable_timing=True); e.record(dst_)
                timeline.append((e2e_no[0], i, "d2h1", e, time.perf_counter()))
        return evs

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=n_ctx)

    e2e_no = [0]
    host_prof = [] if os.environ.get("BENCH_HOST_PROFILE") else None   # debugging aid: host time of the C-ABI calls, every thread
    timeline = [] if os.environ.get("BENCH_TIMELINE") else None      # debugging aid: device + host time of every phase of every picture

    def step_e2e(hand_off=True):
        ev_start = torch.cuda.Event()
        ev_start.record(stream)                     # after the previous step's hand-off
        half = e2e_no[0] % n_sets
        e2e_no[0] += 1
        for evs in [f.result() for f in [pool.submit(e2e_worker, t, ev_start, half) for t in range(n_ctx)]]:
            for ev in evs:
                stream.wait_event(ev)               # the hand-off waits, on the device, for the step's pictures to be final (their
        if hand_off:                                # D2H copies run on under the next step's kernels; the closing barrier covers them)
            hand_off_reference(e_dev[half][0])
        return half

    step_e2e(); step_e2e()
    e2e_steps = max(2, args.steps)
    e2e_runs = []
    for _ in range(max(1, args.repeats)):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_runs))
    clocks = sampler.stop()

    # ---- verification of the end-to-end leg (untimed): one more end-to-end step and one device-resident step with NO reference
    # hand-off in between, so both see the same reference pictures.  Every picture that came back through the pinned host buffers
    # must be identical to the device-resident leg's, and (rank 0, one GPU) picture 0 must equal the CPU oracle chain on those
    # references -- a fast end-to-end number over wrong pictures is worth nothing.
    import helpers
    half = step_e2e(hand_off=False)
    barrier()
    e2e_md5 = [helpers.md5(*np_out[half][i]) for i in range(P)]
    e2e_pic0 = [a.copy() for a in np_out[half][0]]
    step_device(hand_off=False)
    barrier()
    dev_md5 = [helpers.md5(*dev[i]["dst"].download()) for i in range(P)]
    if e2e_md5 != dev_md5:
        raise SystemExit("bench.py: end-to-end pictures differ from the device-resident ones: %s vs %s" % (e2e_md5, dev_md5))
    e2e_check = {"pictures_identical_to_device_leg": P, "md5_picture0": e2e_md5[0], "oracle_pictures_checked": 0}
    if rank == 0 and world == 1 and not os.environ.get("BENCH_SKIP_ORACLE_CHECK"):
        import helpers_chain
        from oracle import pyoracle as po_chk
        refs_now = [[pic.download() for pic in lst] for lst in ref_pics]
        want = helpers_chain.oracle_picture(po_chk, dict(wl, refs=refs_now), pics[0])
        helpers.assert_planes_equal(e2e_pic0, want, "bench.py: end-to-end picture 0 vs the oracle chain")
        e2e_check["oracle_pictures_checked"] = 1
    if os.environ.get("OVGPU_HOST_TIMING"):
        api.lib().ovgpu_host_timing_dump()
    if host_prof:
        a = np.array(host_prof)
        for k, name in enumerate(("mc_commit", "tu_commit", "intra_upload", "dbf_commit", "sao+alf", "recon_call", "inloop_calls")):
            sys.stderr.write("host ms %-13s median %.2f p90 %.2f max %.2f\n" % (name, np.median(a[:, k]), np.percentile(a[:, k], 90), a[:, k].max()))
        sys.stderr.write("host ms total per picture: median %.2f; step wall %.2f ms\n" % (np.median(a.sum(1)), 1e3 * e2e_s / e2e_steps))
    if timeline:
        base = min(timeline, key=lambda r: r[4])
        for (stp, i, name, e, th) in sorted(timeline, key=lambda r: (r[0], r[1], r[4])):
            sys.stderr.write("timeline step %d pic %d %-5s gpu %8.3f ms host %8.3f ms\n" % (stp, i, name, base[3].elapsed_time(e), 1e3 * (th - base[4])))

    # ---- max over ranks
    # every run: max over ranks; the reported number is the MEDIAN run (all runs are listed beside it)
    if world > 1:
        t = torch.tensor(ms_runs + e2e_runs, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_runs, e2e_runs = [float(v) for v in t[:len(ms_runs)]], [float(v) for v in t[len(ms_runs):]]
    ms, e2e_s = float(np.median(ms_runs)), float(np.median(e2e_runs))
    if world > 1:
        t = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        launches = int(t[0])

    if rank == 0:
        fps = (pics_per_step * args.steps if args.dag else world * P * args.steps) / (ms * 1e-3)
        e2e_fps = world * P * e2e_steps / e2e_s
        peak, peak_src = peaks()
        dirs = float(np.mean([mc_dirs_per_sample(p["cmds"]) for p in pics]))
        L, Cc = w * h * 2, w * h     # luma plane bytes, both chroma planes bytes
        # algorithmic bytes per launch: MC (dirs+1)*P, SAO 2P, ALF luma 2L, ALF chroma 2C
        alg = {"mc": (dirs + 1) * pbytes, "residual": float(np.mean([residual_bytes(p["tus"]) for p in pics])),
               "intra": float(np.mean([3.0 * ((1 << p["icmds"]["log2_w"].astype(np.int64)) * (1 << p["icmds"]["log2_h"].astype(np.int64))).sum() for p in pics])),
               "lmcs": 2 * L, "dbf_v": 2 * pbytes, "dbf_h": 2 * pbytes, "sao": 2 * pbytes, "alf_luma": 2 * L, "alf_chroma": 2 * Cc}
        stages = {}
        for s in stage_names:
            t_ms, n = prof[s]
            if n:
                # per PICTURE: a stage may take several launches per picture (the residual and intra stages run one
                # group of launches per dependency level, the MC stage one kernel per tool)
                avg = t_ms / (nprof * P)
                stages[s] = {"avg_launch_ms": avg, "scopes_per_picture": n / (nprof * P), "algorithmic_bytes": alg[s],
                             "achieved_gbs": alg[s] / (avg * 1e-3) / 1e9, "frac": alg[s] / (avg * 1e-3) / 1e9 / peak,
                             "traffic": ncu_traffic(s)[0]}
        top = max(stages, key=lambda s: stages[s]["avg_launch_ms"])
        h2d = sum(p["cmds"].nbytes + p["prof"].nbytes + p["tus"].nbytes + 3 * 8 * len(p["tus"]) + p["coeffs"].nbytes + p["icmds"].nbytes + 8 * len(p["icmds"]) + p["isp_words"].nbytes + p["anchors"].nbytes + p["dbf"]["ver"].nbytes + p["dbf"]["hor"].nbytes + p["dbf"]["qp"].nbytes + p["sao"].nbytes + 24 * 100 * 36 * 4 + 8 * len(p["alf_ctu"]) for p in pics)
        line = {"metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.dag else "weak", "vs_baseline": None, "dtype": "u16/int32", "data": "synthetic",
                "config": workload_config(w, h, pics_per_step, n_ds, args.dag),
                "gpu_launches": launches,
                "repeats": {"runs": len(ms_runs), "reported": "median", "value_runs": [(pics_per_step * args.steps if args.dag else world * P * args.steps) / (m * 1e-3) for m in ms_runs],
                            "e2e_runs": [world * P * e2e_steps / s_ for s_ in e2e_runs]},
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": P * pbytes, "steps": e2e_steps, "streams": len(real_streams), "host_threads": n_ctx, "pictures_in_flight": n_sets * P,
                        "check": e2e_check},
                "roofline": {"kernel": top, "bound": "hbm", "achieved": stages[top]["achieved_gbs"], "peak": peak,
                             "peak_source": peak_src, "unit": "GB/s", "frac": stages[top]["frac"], "traffic": ncu_traffic(top)[0],
                             "traffic_source": ncu_traffic(top)[1],
                             "algorithmic_bytes_per_launch": stages[top]["algorithmic_bytes"],
                             "avg_launch_ms": stages[top]["avg_launch_ms"]},
                "stages": stages, "mc_dirs_per_sample": dirs,
                "clocks": clocks}
        if dag_info:
            line["dag"] = {k: v for k, v in dag_info.items() if k != "launches"}
        # CPU baseline on the box's host cores (N=1 only): the compiled reference on a bounded sample
        if world == 1:
            from oracle import pyoracle as po
            cores = os.cpu_count() or 1
            reference_step(po, wl, cores)
            tot, n = 0.0, 0
            while tot < 10.0 and n < 128:
                dt, k = reference_step(po, wl, cores)
                tot += dt; n += k
            line["cpu_baseline"] = {"value": n / tot, "unit": "frames/s", "cores": cores,
                                    "kind": "reference" if po.ref_available() else "port",
                                    "sample": "%d pictures (%d per round, one per host thread, every host core), same stages" % (n, cores)}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
