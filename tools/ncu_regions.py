#!/usr/bin/env python
"""Where the warps of a raster launch spend their time: stall samples of an `ncu --set full
--import-source on` report summed over the kernel's three phases (prologue + scan + staging,
composite loop, epilogue + finish), with the top stall reasons of each.
    python tools/ncu_regions.py report.ncu-rep"""
import csv, io, subprocess, sys

def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, v = r[0], r[-1]
    pick = ("gpu__time_duration.sum", "launch__grid_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
    print("kernel:", v[h.index("Kernel Name")][:60])
    for k, val in zip(h, v):
        if k in pick:
            print(f"  {k:66s} {val}")
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    for hi, row in enumerate(rows):
        if 'Source' in row and 'Instructions Executed' in row:
            break
    hdr = rows[hi]
    c = {x: i for i, x in enumerate(hdr)}
    stall_cols = [(i, x) for i, x in enumerate(hdr) if x.startswith('stall_') and 'Not Issued' not in x]
    data = [x for x in rows[hi + 1:] if len(x) == len(hdr)]
    tot = sum(int(x[c['# Samples']] or 0) for x in data)
    src = [x[c['Source']] for x in data]
    first_mufu = next(k for k, s in enumerate(src) if 'MUFU.EX2' in s)
    a = max(k for k in range(first_mufu) if 'BAR.SYNC' in src[k])
    ends = [k for k in range(first_mufu, len(src)) if 'BAR.RED' in src[k] or 'BAR.SYNC' in src[k]]
    b = ends[0]

    def summ(lo, hi, name):
        n = sum(int(x[c['# Samples']] or 0) for x in data[lo:hi])
        ex = sum(int(x[c['Instructions Executed']] or 0) for x in data[lo:hi])
        st = {}
        for x in data[lo:hi]:
            for i, hname in stall_cols:
                st[hname[6:]] = st.get(hname[6:], 0) + int(x[i] or 0)
        top = ", ".join(f"{k} {v_}" for k, v_ in sorted(st.items(), key=lambda kv: -kv[1])[:5])
        print(f"  {name:26s} {100 * n / max(tot, 1):5.1f} % of samples, {ex:10d} warp instructions; stalls: {top}")
    print(f"  stall samples: {tot}")
    summ(0, a + 1, "prologue + scan + staging")
    summ(a + 1, b + 1, "composite loop")
    summ(b + 1, len(data), "epilogue + finish")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
