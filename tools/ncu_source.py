#!/usr/bin/env python
"""Per-SASS-instruction view of an .ncu-rep: executed count, stall samples and top stall
reason.   python tools/ncu_source.py rep.ncu-rep [min_samples]"""
import csv, io, subprocess, sys

def main(path, min_samples=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    for hi, r in enumerate(rows):
        if 'Source' in r and 'Instructions Executed' in r:
            break
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_')]
    total = sum(int(r[c['# Samples']] or 0) for r in rows[hi+1:] if len(r) == len(hdr))
    print(f"total samples {total}")
    for r in rows[hi+1:]:
        if len(r) != len(hdr): continue
        n = int(r[c['# Samples']] or 0)
        if n < min_samples: continue
        st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        print(f"{r[c['Address']][-5:]} {int(r[c['Instructions Executed']] or 0):>11d} {n:>6d} {100*n/total:5.1f}%  {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {r[c['Source']][:70]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
