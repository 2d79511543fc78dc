"""Driver for ncu captures at the benchmark's configuration: the whole stage chain of ONE benchmark picture between
cudaProfilerStart / Stop, after warm-up pictures and an L2 flush (so that every kernel's DRAM traffic is what it moves when the
working set of the 8-picture step has evicted its inputs, not a re-read of lines the previous launch left in the 126 MB L2).

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/x python tools/ncu_picture.py
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv python tools/ncu_picture.py

Exactly one picture is captured, so per-kernel numbers are per picture by construction (tools/ncu_summary.py records `pictures: 1`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--size", default="4k")
ap.add_argument("--stage", default="all", choices=["all", "mc"])
ap.add_argument("--all-intra", action="store_true")
ap.add_argument("--no-flush", action="store_true")
args = ap.parse_args()
import torch
import bench
from openvvc_b200 import api
w, h = bench.SIZES[args.size]
ctx = api.Ctx(0)
wl = bench.make_workload(w, h, 2, all_intra=args.all_intra)
refs = [[ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
for lst, src in zip(refs, wl["refs"]):
    for pic, planes in zip(lst, src):
        pic.upload(planes)
lmcs = api.Lmcs(ctx, wl["lmcs"]); lmcs.set_data(wl["lmcs_data"])
dev = [bench.make_dev(api, ctx, p, w, h) for p in wl["pics"]]
ctx.sync()


def run(d):
    if args.stage == "mc":
        api.recon_picture(ctx, d["rec"], mc=d["mc"], ref_lists=refs, pocs=wl["pocs"], lmcs=lmcs)
    else:
        bench.run_stages(api, ctx, d, refs, wl["pocs"], lmcs)


for _ in range(2):
    for d in dev:
        run(d)
ctx.sync()
if not args.no_flush:
    ctx.flush_l2()
    ctx.sync()
torch.cuda.profiler.start()
run(dev[0])
ctx.sync()
torch.cuda.profiler.stop()
