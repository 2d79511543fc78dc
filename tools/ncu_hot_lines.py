"""Top source lines of one kernel launch of an .ncu-rep captured with --import-source on (-lineinfo builds): share of
executed instructions and of stall samples per CUDA source line.
Usage: python tools/ncu_hot_lines.py report.ncu-rep <kernel-name-filter> <launch-skip> [top]"""
import csv
import subprocess
import sys


def main(rep, name, skip, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", name,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, agg, fn = None, None, {}, ""
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            fn = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[2] == "-":
            try:
                agg[(cur, int(r[0]))] = (int(r[4]), int(r[7]), r[1].strip()[:110])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(fn, "samples", tot, "warp instructions", toti)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%-20s %4d inst %4.1f%% stall %4.1f%% | %s" % (k[0], k[1], 100 * v[1] / toti, 100 * v[0] / tot, v[2]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 25)
