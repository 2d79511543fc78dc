"""Development aid: per-stage device times of the benchmark workload, without the end-to-end and CPU-baseline legs of bench.py.

    python tools/stage_bench.py [--size 4k] [--pics 8] [--reps 5] [--lib path/to/variant.so] [--md5]

Prints one JSON line: per-stage microseconds per picture (CUDA events around each stage, one stream), the multi-stream
device-resident pictures/s, and with --md5 the MD5 of every output picture (to compare library variants bit for bit).
Nothing here is a benchmark result: bench.py is.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="4k")
    ap.add_argument("--pics", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--md5", action="store_true")
    ap.add_argument("--all-intra", action="store_true")
    ap.add_argument("--threads", default="", help="e.g. 1,2,4: also time the multi-stream run issued from that many host threads (is the one issuing thread the limit?)")
    ap.add_argument("--libs", default="", help="comma list of library variants (\"default\" = the product build) timed one after the other on ONE workload, one JSON line each")
    args = ap.parse_args()
    import bench
    from openvvc_b200 import api
    w, h = bench.SIZES[args.size]
    wl = bench.make_workload(w, h, args.pics, seed0=5000, all_intra=args.all_intra)
    default_path, default_build = api.LIB_PATH, api._build.build
    for lib in ([v for v in args.libs.split(",") if v] or [args.lib]):
        args.lib = None if lib in (None, "default") else lib
        api._lib = None
        api.LIB_PATH, api._build.build = default_path, default_build
        run_one(args, wl, w, h)


def run_one(args, wl, w, h):
    import torch
    import bench
    from openvvc_b200 import api
    if args.lib:
        api.LIB_PATH = os.path.abspath(args.lib)
        api._build.build = lambda *a, **k: api.LIB_PATH
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = api.Ctx(0, stream.cuda_stream)
    refs = [[ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
    for lst, src in zip(refs, wl["refs"]):
        for pic, planes in zip(lst, src):
            pic.upload(planes)
    lmcs = api.Lmcs(ctx, wl["lmcs"])
    lmcs.set_data(wl["lmcs_data"])
    dev = [bench.make_dev(api, ctx, p, w, h) for p in wl["pics"]]
    ctx.sync()
    n_ds = max(1, min(args.streams, args.pics))
    streams = [stream] + [torch.cuda.Stream() for _ in range(n_ds - 1)]
    ctxs = [ctx] + [api.Ctx(0, s.cuda_stream) for s in streams[1:]]

    def step(parallel):
        if not parallel:
            for d in dev:
                bench.run_stages(api, ctx, d, refs, wl["pocs"], lmcs)
            return
        ev0 = torch.cuda.Event(); ev0.record(stream)
        for s in streams[1:]:
            s.wait_event(ev0)
        for i, d in enumerate(dev):
            bench.run_stages(api, ctxs[i % n_ds], d, refs, wl["pocs"], lmcs)
        for s in streams[1:]:
            ev = torch.cuda.Event(); ev.record(s); stream.wait_event(ev)

    for _ in range(3):
        step(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.reps * 2):
        step(True)
    e1.record(stream)
    torch.cuda.synchronize()
    fps = args.pics * args.reps * 2 / (e0.elapsed_time(e1) * 1e-3)
    e0.record(stream)
    for _ in range(args.reps):
        step(False)
    e1.record(stream)
    torch.cuda.synchronize()
    fps1 = args.pics * args.reps / (e0.elapsed_time(e1) * 1e-3)
    # host threads: thread t issues the pictures of the streams j with j % T == t, nothing else changes (wall clock around issue + sync)
    fps_threads, issue_ms = {}, {}
    for T in [int(v) for v in args.threads.split(",") if v]:
        import threading
        import time
        torch.cuda.synchronize()
        n_rounds = args.reps * 2

        def worker(t, T=T):
            for _ in range(n_rounds):
                for i, d in enumerate(dev):
                    if (i % n_ds) % T == t:
                        bench.run_stages(api, ctxs[i % n_ds], d, refs, wl["pocs"], lmcs)
        th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        fps_threads[T] = round(args.pics * n_rounds / (t2 - t0), 1)
        issue_ms[T] = round((t1 - t0) * 1e3 / (args.pics * n_rounds), 4)        # host time per picture until everything is issued
    names = ("mc", "residual", "intra", "lmcs", "dbf_v", "dbf_h", "sao", "alf_luma", "alf_chroma")
    ctx.profile(True); ctx.profile_reset()
    for _ in range(args.reps):
        step(False)
    prof = {s: ctx.profile_get(s) for s in names}
    ctx.profile(False)
    out = {"lib": args.lib or "default", "size": args.size, "fps_multi_stream": round(fps, 1), "fps_one_stream": round(fps1, 1),
           "us_per_picture": {s: round(1e3 * prof[s][0] / (args.reps * args.pics), 1) for s in names if prof[s][1]}}
    out["us_sum"] = round(sum(out["us_per_picture"].values()), 1)
    if fps_threads:
        out["fps_by_host_threads"] = fps_threads
        out["host_issue_ms_per_picture"] = issue_ms
    if args.md5:
        import helpers
        out["md5"] = [helpers.md5(*d["dst"].download()) for d in dev]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
