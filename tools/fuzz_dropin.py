#!/usr/bin/env python3
"""Differential fuzzing of the drop-in over the emulated kernels (no GPU needed): fabricate random streams with tools/vvc_stream.py,
decode them with the stock scalar-C decoder (oracle/_ref/dectest_c) and with build/dectest_emu (shim + C ABI over the CPU emulation of
the kernel sources) at several frame / entry thread counts, report every stream whose output differs.

    python tools/fuzz_dropin.py --mode tiles --seconds 300 --seed 1      random tile grids (pictures of several rect entries)
    python tools/fuzz_dropin.py --mode sizes --seconds 300               single-tile pictures of odd sizes (multiples of 8), incl. one CTU row
    python tools/fuzz_dropin.py --mode rpr   --seconds 300               reference picture resampling with random picture sizes

Streams the fabricator cannot write (its own asserts on the size combination, or the reference's generative run crashing on one- and
two-CTU-wide pictures) are counted as `unfabricated`, not as failures of the back end."""
import argparse
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vvc_stream as vs  # noqa: E402
import test_dropin as td  # noqa: E402


def split(rnd, n):
    parts, left = [], n
    while left > 0:
        p = rnd.randint(1, min(3, left))
        parts.append(p)
        left -= p
    return parts


def draw(rnd, mode):
    kw = dict(gop=rnd.choice([1, 4, 8]), seed=rnd.randint(1000, 99999), preset="ctc")
    if mode == "tiles":
        w, h = rnd.choice([384, 416, 512, 640, 832]), rnd.choice([128, 136, 240, 256, 264, 384, 480])
        cols, rows = split(rnd, (w + 127) // 128), split(rnd, (h + 127) // 128)
        if len(cols) * len(rows) == 1:
            return None
        kw["tiles"] = (cols, rows)
        return w, h, rnd.choice([3, 5]), kw, ((1, 1), (4, 4))
    if mode == "sizes":
        w, h = 8 * rnd.randint(48, 80), 8 * rnd.randint(8, 48)
        return w, h, rnd.choice([3, 5]), kw, ((1, 1), (4, 1))
    w, h = 8 * rnd.randint(48, 80), 8 * rnd.randint(17, 48)
    kw["preset"] = "rpr"
    kw["rpr_sizes"] = [(8 * rnd.randint(max(17, w // 16 + 1), min(80, w * 2 // 8)), 8 * rnd.randint(max(17, h // 16 + 1), min(60, h * 2 // 8)))
                       for _ in range(rnd.randint(1, 3))]
    return w, h, rnd.choice([4, 6]), kw, ((1, 1), (4, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["tiles", "sizes", "rpr"], default="tiles")
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    emu, ref = os.path.join(td.BUILD, "dectest_emu"), os.path.join(td.REF, "dectest_c")
    t0, streams, bad, unfab = time.time(), 0, 0, 0
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "f.266")
        while time.time() - t0 < a.seconds:
            case = draw(rnd, a.mode)
            if case is None:
                continue
            w, h, n, kw, configs = case
            try:
                vs.make_stream(path, w, h, n, verify=True, **kw)
            except Exception:
                unfab += 1
                continue
            want = td.decode(ref, path, os.path.join(tmp, "c.yuv"))[0]
            for threads, entries in configs:
                try:
                    got = td.decode(emu, path, os.path.join(tmp, "e.yuv"), threads=threads, entries=entries, backend="cuda")[0]
                except AssertionError as e:
                    got = "error: " + str(e)[-200:]
                if got != want:
                    bad += 1
                    print("MISMATCH", w, h, n, kw, "threads", threads, "entries", entries, got[:200], flush=True)
            streams += 1
    print({"mode": a.mode, "seed": a.seed, "streams": streams, "mismatches": bad, "unfabricated": unfab, "seconds": round(time.time() - t0)})
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
