#!/usr/bin/env python
"""Compile ggs_raster.cu under a grid of flags and report, per build, registers, spills and the
number of register copies / special-register reads ptxas left in the composite loop (no GPU
needed).  ptxas' allocation of the packed accumulators is chaotic; this finds the clean builds."""
import itertools, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "genetic-gaussian-splats_b200", "csrc", "ggs_raster.cu")
BASE = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
        "-I", os.path.join(ROOT, "include"), "-I", os.path.dirname(SRC)]

def analyse(obj):
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    prod = [f for f in funcs if "raster_kernelILb0ELb0E" in f.split("\n", 1)[0]][0]
    ops = [m.group(1) for m in (re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", l) for l in prod.splitlines()) if m]
    first = next(i for i, o in enumerate(ops) if "MUFU.EX2" in o)
    end = next(i for i in range(first, len(ops)) if re.match(r"(@!?U?P\d+\s+)?BRA\b", ops[i]))
    block = ops[first:end]
    copies = sum(bool(re.match(r"(IMAD\.MOV|MOV)\b", o)) for o in block)
    pro = ops[max(0, first - 45):first]
    remat = sum(bool(re.search(r"\b(S2R|S2UR|I2FP|LDL)\b", o)) for o in pro)
    return copies, remat, len(block)

def main(extra_sets):
    for extra in extra_sets:
        obj = "/tmp/ggs/scan.o"
        r = subprocess.run(BASE + extra + ["-Xptxas", "-v", "-c", SRC, "-o", obj], capture_output=True, text=True)
        info = [l for l in r.stderr.splitlines() if "Used" in l]
        if r.returncode != 0 or not info:
            print(" ".join(extra), "-> compile failed"); continue
        regs = re.search(r"Used (\d+) registers", info[0]).group(1)
        spill = "spill" if "stack" in info[0] else ""
        copies, remat, blk = analyse(obj)
        print(f"{' '.join(extra):70s} regs {regs:>3s} {spill:5s} copies {copies:2d}  remat-in-prologue {remat}  block {blk}")

if __name__ == "__main__":
    defs = sys.argv[1:] or [""]
    sets = []
    for d in defs:
        for opt in ("-O1", "-O2", "-O3"):
            for mb in (6, 7, 8):
                sets.append([x for x in d.split() if x] + ["-Xptxas", opt, f"-DGGS_MIN_BLOCKS={mb}"])
    main(sets)
