#!/usr/bin/env python
"""bench.py -- frames/sec of the rcn_* reconstruction hot path on synthetic 4K/8K 10-bit pictures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 4k|8k]

One "step" = one pass of the hot path over a batch of `pictures_per_step` synthetic pictures per GPU
(frame-parallel sharding: every rank owns whole pictures); value = pictures/s over all
ranks with inputs resident in HBM; e2e = the same through the host-buffer C-ABI calls (H2D of planes and
parameters + D2H of the result inside the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZES = {"1080p": (1920, 1080), "4k": (3840, 2160), "8k": (7680, 4320)}
STAGES_IMPLEMENTED = ["sao", "alf"]       # widened as §8 rows land; named in config.workload


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


def make_workload(w, h, n_pics, seed0=5000):
    """n_pics synthetic pictures with their SAO/ALF decoder state (SURVEY.md §8(d) generators)."""
    import helpers
    from oracle import pyoracle as po   # only for the table struct layout helper + cpu arms
    from openvvc_b200 import api, synth
    pics = []
    for i in range(n_pics):
        planes = synth.planes(w, h, seed0 + i, "natural")
        ap = synth.alf_params(w, h, seed0 + 100 + i, 7)
        tab = api.alf_tables_from_aps(ap["luma_aps"], ap["chroma_aps"])
        tab.np("cc_coeff")[:] = ap["cc_coeff"]
        sao = synth.sao_params(w, h, seed0 + 200 + i, 7, api.SAO_CTU)
        pics.append(dict(planes=planes, tab=tab, ctu=ap["ctu"], alf_ctu=api.alf_ctu_array(ap["ctu"]), sao=sao))
    return pics


# ------------------------------------------------------------------------------------------------ reference arm
def reference_step(po, pics, threads):
    """The reference's own CPU implementation (oracle/_ref, SSE4.1/AVX2 table) of the same stages, one picture
    per host thread (the reference's frame parallelism, ovdec.h:47)."""
    import ctypes as C

    def one(p):
        t = po.AlfTables()
        C.memmove(C.byref(t), C.byref(p["tab"]), C.sizeof(t))
        if po.ref_available():
            s = po.sao_picture(p["planes"], p["sao"].astype(po.SaoCtu), 7, impl="ref", simd=1)
            po.alf_picture(s, t, p["ctu"], 7, impl="ref", simd=1)
        else:
            s = po.sao_picture(p["planes"], p["sao"].astype(po.SaoCtu), 7, impl="port")
            po.alf_picture(s, t, p["ctu"], 7, impl="port")

    ths = [threading.Thread(target=one, args=(p,)) for p in pics[:threads]]
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
    pics = make_workload(w, h, min(cores, 8))
    threads = len(pics)
    for _ in range(max(1, min(args.warmup, 1))):
        reference_step(po, pics, threads)
    tot, n = 0.0, 0
    for _ in range(args.steps):
        dt, k = reference_step(po, pics, threads)
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
    print(json.dumps(line))


def workload_config(w, h, pics_per_step):
    return {"workload": "%dx%d 10-bit 4:2:0 CTU128 synthetic RA picture, rcn stages: %s (in-loop filters; "
                        "remaining rcn stages not yet in the timed path)" % (w, h, "+".join(STAGES_IMPLEMENTED)),
            "pictures_per_step": pics_per_step, "parallelism": "frame-parallel, one picture stream per GPU",
            "l2": "inputs larger than L2: %d pictures x (src+mid+dst) per step" % pics_per_step}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="4k", choices=sorted(SIZES))
    ap.add_argument("--pics", type=int, default=8, help="pictures per step per GPU")
    args = ap.parse_args()
    w, h = SIZES[args.size]

    if args.impl == "reference":
        run_reference(args, w, h)
        return

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
    pics = make_workload(w, h, P, seed0=5000 + 1000 * rank)
    pbytes = w * h * 2 * 3 // 2

    # device-resident state: per picture src (deblocked), mid (post-SAO), dst (post-ALF) + parameters
    dev = []
    for p in pics:
        d = dict(src=ctx.picture(w, h), mid=ctx.picture(w, h), dst=ctx.picture(w, h),
                 alf=api.AlfParams(ctx, w, h, 7), sao=api.SaoParams(ctx, w, h, 7))
        d["src"].upload(p["planes"])
        d["alf"].upload(p["tab"], p["alf_ctu"], (p["ctu"]["cc_cb"], p["ctu"]["cc_cr"]))
        d["sao"].upload(p["sao"])
        dev.append(d)
    ctx.sync()

    def step_device():
        for d in dev:
            d["sao"].run(d["mid"], d["src"])
            d["alf"].run(d["dst"], d["mid"])

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
    ctx.profile(True); ctx.profile_reset()
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - n0
    prof = {s: ctx.profile_get(s) for s in ("alf_luma", "alf_chroma", "sao")}
    ctx.profile(False)

    # ---- e2e: host buffers (pinned), H2D planes + params, stages, D2H result, every step
    pinned_in = [[torch.from_numpy(pl).pin_memory() for pl in p["planes"]] for p in pics]
    pinned_out = [[torch.empty_like(t).pin_memory() for t in pin] for pin in pinned_in]
    np_in = [[t.numpy() for t in pin] for pin in pinned_in]
    np_out = [[t.numpy() for t in pin] for pin in pinned_out]

    def step_e2e():
        for i, d in enumerate(dev):
            p = pics[i]
            d["src"].upload(np_in[i])
            d["sao"].upload(p["sao"])
            d["alf"].upload(p["tab"], p["alf_ctu"], (p["ctu"]["cc_cb"], p["ctu"]["cc_cr"]))
            d["sao"].run(d["mid"], d["src"])
            d["alf"].run(d["dst"], d["mid"])
            api._check(ctx.h, api.lib().ovgpu_picture_download(ctx.h, d["dst"].p, *api._host_planes(np_out[i])), "download")
        ctx.sync()

    step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        fps = world * P * args.steps / (ms * 1e-3)
        e2e_fps = world * P * e2e_steps / e2e_s
        peak, peak_src = peaks()
        luma_ms, luma_n = prof["alf_luma"]
        luma_bytes = 2 * w * h * 2                       # read + write of the luma plane
        achieved = luma_bytes / (luma_ms / luma_n * 1e-3) / 1e9 if luma_n else None
        tab_bytes = 24 * 100 * 36 * 4 + 8 * (((w + 127) // 128) * ((h + 127) // 128))
        sao_bytes = pics[0]["sao"].nbytes
        line = {"metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u16/int32", "data": "synthetic",
                "config": workload_config(w, h, P),
                "gpu_launches": launches,
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": P * (pbytes + tab_bytes + sao_bytes),
                        "d2h_bytes_per_step": P * pbytes, "steps": e2e_steps},
                "roofline": {"kernel": "AlfLumaK (classification + 7x7 luma ALF)", "bound": "hbm", "achieved": achieved,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": achieved / peak if achieved else None, "traffic": None,
                             "algorithmic_bytes_per_launch": luma_bytes, "avg_launch_ms": luma_ms / luma_n if luma_n else None},
                "stage_ms_per_picture": {s: (prof[s][0] / prof[s][1] if prof[s][1] else None) for s in prof},
                "clocks": clocks}
        # CPU baseline on the box's host cores (N=1 only): the compiled reference on a bounded sample
        if world == 1:
            from oracle import pyoracle as po
            cores = min(os.cpu_count() or 1, P)
            reference_step(po, pics, cores)
            tot, n = 0.0, 0
            while tot < 10.0 and n < 64:
                dt, k = reference_step(po, pics, cores)
                tot += dt; n += k
            line["cpu_baseline"] = {"value": n / tot, "unit": "frames/s", "cores": cores,
                                    "kind": "reference" if po.ref_available() else "port",
                                    "sample": "%d pictures (%d per round, one per host thread), same stages" % (n, cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
