"""SASS opcode summary of the shipped library, per kernel: `cuobjdump -sass openvvc_b200/libovvc_b200.so`, mnemonics counted.
Usage: python tools/sass_summary.py [out.json]   (prints a table; the JSON goes under profiles/ as evidence of what the kernels are made of)"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openvvc_b200", "libovvc_b200.so")
KEY = ["LDGSTS", "UTMALDG", "UTMASTG", "SYNCS", "SHFL", "BAR", "IDP", "VIADDMNMX", "VIMNMX", "PRMT", "IMAD", "LDG", "LDG.128", "LDG.64", "LDG.U16", "STG", "STG.128", "STG.64",
       "LDS", "LDS.128", "LDS.64", "STS", "ATOM/RED", "LDC"]


def main(out=None):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ovd_kernel<(.*)>\(.*$", r"\1", name)
            name = re.sub(r"\(.*$", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            base = op.split(".")[0]
            cur["total"] += 1
            cur[base] += 1
            if base in ("LDG", "STG", "LDS", "STS"):
                w = "128" if ".128" in op else "64" if ".64" in op else "U16" if ".U16" in op or ".S16" in op else None
                if w:
                    cur[base + "." + w] += 1
            if base in ("ATOM", "ATOMG", "ATOMS", "RED"):
                cur["ATOM/RED"] += 1
    doc = {"library": os.path.relpath(LIB, ROOT), "command": "cuobjdump -sass", "kernels": {k: {o: c[o] for o in ["total"] + KEY if c[o]} for k, c in kernels.items()}}
    print("%-28s %6s " % ("kernel", "total") + " ".join("%7s" % k[:7] for k in KEY[:12]))
    for k, c in kernels.items():
        print("%-28s %6d " % (k[:28], c["total"]) + " ".join("%7d" % c[o] for o in KEY[:12]))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    doc["all_kernels"] = {o: tot[o] for o in ["total"] + KEY if tot[o]}
    if out:
        json.dump(doc, open(out, "w"), indent=1)
        print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
