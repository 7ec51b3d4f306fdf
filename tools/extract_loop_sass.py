#!/usr/bin/env python
"""SASS of the raster kernel's composite loop (between the barrier that publishes the staged
list and the BAR.RED that ends a flush), for profiles/rNN_raster_sass_composite_loop.txt.
    python tools/extract_loop_sass.py > profiles/r01_raster_sass_composite_loop.txt"""
import os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "genetic-gaussian-splats_b200", "build", "ggs_raster.o")
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
fn, keep = [], False
for line in out.splitlines():
    if "Function :" in line:
        keep = "raster_kernelILb0ELb0E" in line
    if keep and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        fn.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line))
first_mufu = next(i for i, l in enumerate(fn) if "MUFU.EX2" in l)
start = max(i for i in range(first_mufu) if "BAR.SYNC" in fn[i]) + 1
end = next(i for i in range(first_mufu, len(fn)) if "BAR.RED" in fn[i])
body = fn[start:end]
n = lambda pat: sum(1 for l in body if re.search(pat, l))
print("# SASS of ggs::raster_kernel<false> (sm_100a, nvcc 12.9, -O3), composite loop only (extracted by")
print("# tools/extract_loop_sass.py).  Per list entry: band byte (PRMT/ISETP/BRA), lane mask (LOP3 -> predicate),")
print("# per-thread setup, then the recurrence path (4 MUFU.EX2 + packed FMUL2/FFMA2/FADD2 in place on the named")
print("# accumulator registers) or the exact path (per-pixel MUFU.EX2, rows selected by uniform branches).")
print(f"# In this extract: FFMA2 {n('FFMA2')}, FMUL2 {n('FMUL2')}, FADD2 {n('FADD2')}, MUFU.EX2 {n('MUFU.EX2')}, "
      f"LDS.128 {n('LDS.128')}, MOV/IMAD.MOV {n(r'[^U]MOV|IMAD.MOV')}, S2R {n(r' S2R ')}, LDL/STL {n('LDL|STL')}.")
print("# (CS2R Rn, SRZ in the exact path zeroes a falloff pair; UMOV / IMAD.MOV sit in the set-up ahead of the loop head.)")
print("# Full dump: cuobjdump -sass genetic-gaussian-splats_b200/build/ggs_raster.o\n")
print("\n".join(body))
