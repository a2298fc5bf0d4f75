#!/usr/bin/env python
"""Opcode mix of the hot kernels from `cuobjdump -sass chroma_b200/_obj/engine_d.o` -> profiles/r02_sass_summary.json
(+ the full listings, gzipped).  Runs here (no GPU needed)."""
import collections
import gzip
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "chroma_b200", "_obj", "engine_d.o")
WANT = [r"dslash_kernel<double, \(int\)[1-4], \(bool\)0, \(int\)128, \(int\)0>", r"dslash_halo_kernel<double, \(int\)[14], \(bool\)0",
        r"dslash_mrhs_kernel<double, \(int\)[124], \(bool\)0, \(int\)6, \(int\)0>", r"dslash_finish_kernel<double, \(int\)4>",
        r"cg_update_kernel<double>"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    names = {}
    blocks = re.split(r"\n\s+Function : ", sass)[1:]
    out = {"source": "cuobjdump -sass chroma_b200/_obj/engine_d.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo), scripts/sass_summary.py; "
                     "full listings of these kernels: profiles/r02_sass_dslash_kernels.txt.gz",
           "reading": "the stencil kernels are 128-bit global loads (LDG.E.128[.CONSTANT]) + DFMA; no tensor-core instruction is expected or present; "
                      "the batched kernel stages links / clover / prefetched spinors with cp.async (LDGSTS), not TMA, and its site decode is "
                      "division-free (1 MUFU.RCP left, on the cold boundary-box path)",
           "kernels": {}}
    listing = []
    for b in blocks:
        mangled = b.split("\n", 1)[0].strip()
        dem = names.get(mangled) or subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip()
        if not any(re.search(w, dem) for w in WANT):
            continue
        ops = collections.Counter(); ldg = collections.Counter(); n = 0
        for line in b.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            n += 1
            full = m.group(1)
            ops[full.split(".")[0]] += 1
            if full.startswith("LDG."):
                ldg[full] += 1
        w128 = sum(v for k, v in ldg.items() if ".128" in k); w64 = sum(v for k, v in ldg.items() if ".64" in k)
        out["kernels"][dem] = {"instructions": n, "opcodes": dict(ops.most_common()), "LDGSTS (cp.async)": ops.get("LDGSTS", 0),
                               "UTMALDG (TMA)": ops.get("UTMALDG", 0),
                               "UTCMMA/HMMA/DMMA (tensor)": sum(ops.get(k, 0) for k in ("UTCMMA", "HMMA", "DMMA", "IMMA", "QMMA")),
                               "R2UR": ops.get("R2UR", 0), "MUFU (integer-division sequences)": ops.get("MUFU", 0),
                               "global_loads_by_width": {"128-bit": w128, "64-bit": w64, "32-bit/other": sum(ldg.values()) - w128 - w64},
                               "LDG_variants": dict(ldg)}
        listing.append("Function : " + dem + "\n" + b)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_sass_summary.json"), "w"), indent=1)
    with gzip.open(os.path.join(ROOT, "profiles", "r02_sass_dslash_kernels.txt.gz"), "wt") as f:
        f.write("\n\n".join(listing))
    for k, v in out["kernels"].items():
        sys.stdout.write("%5d instr  DFMA %4d  LDGSTS %3d  MUFU %2d  %s\n" % (v["instructions"], v["opcodes"].get("DFMA", 0), v["LDGSTS (cp.async)"], v["MUFU (integer-division sequences)"], k[:100]))


if __name__ == "__main__":
    main()
