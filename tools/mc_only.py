"""Development aid: run the MC stage of one benchmark picture a few times (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--lib"); ap.add_argument("--size", default="4k"); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--stage", default="mc")
args = ap.parse_args()
import bench
from openvvc_b200 import api
if args.lib:
    api.LIB_PATH = os.path.abspath(args.lib); api._build.build = lambda *a, **k: api.LIB_PATH
w, h = bench.SIZES[args.size]
ctx = api.Ctx(0)
wl = bench.make_workload(w, h, 1)
refs = [[ctx.picture(w, h) for _ in lst] for lst in wl["refs"]]
for lst, src in zip(refs, wl["refs"]):
    for pic, planes in zip(lst, src):
        pic.upload(planes)
lmcs = api.Lmcs(ctx, wl["lmcs"]); lmcs.set_data(wl["lmcs_data"])
d = bench.make_dev(api, ctx, wl["pics"][0], w, h)
ctx.sync()
for _ in range(args.reps):
    if args.stage == "mc":
        api.recon_picture(ctx, d["rec"], mc=d["mc"], ref_lists=refs, pocs=wl["pocs"], lmcs=lmcs)
    else:
        bench.run_stages(api, ctx, d, refs, wl["pocs"], lmcs)
ctx.sync()
