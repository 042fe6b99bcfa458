#!/usr/bin/env python3
"""Per-SASS-instruction view of an .ncu-rep captured with --import-source on: the instructions that carry the most
stall samples / shared-memory wavefronts.   python tools/ncu_source.py rep [--top 40] [--sort samples|wavefronts|excess]"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--top", type=int, default=40)
ap.add_argument("--sort", default="samples"); ap.add_argument("--all", action="store_true")
a = ap.parse_args()
import os
raw = subprocess.run(["ncu", "-i", os.path.abspath(a.rep), "--page", "source", "--csv"], capture_output=True, text=True, cwd="/tmp").stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[col[n]].replace(",", ""))
    except Exception: return 0.0
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(f(r, "# Samples") for r in body)
key = {"samples": "# Samples", "wavefronts": "L1 Wavefronts Shared", "excess": "L1 Wavefronts Shared Excessive"}[a.sort]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
print(f"total samples {tot:.0f}, instructions {len(body)}, wavefronts {sum(f(r,'L1 Wavefronts Shared') for r in body):.0f} "
      f"excess {sum(f(r,'L1 Wavefronts Shared Excessive') for r in body):.0f} executed {sum(f(r,'Instructions Executed') for r in body):.0f}")
print("stall totals:", {n[6:]: int(sum(f(r, n) for r in body)) for n in stalls if sum(f(r, n) for r in body) > 0.01 * tot})
sel = body if a.all else sorted(body, key=lambda r: -f(r, key))[:a.top]
for r in sel:
    top = sorted(((f(r, n), n[6:]) for n in stalls), reverse=True)[:2]
    print(f"{r[col['Address']][-5:]} {f(r,'# Samples'):6.0f} ex{f(r,'Instructions Executed'):9.0f} wf{f(r,'L1 Wavefronts Shared'):9.0f}/{f(r,'L1 Wavefronts Shared Ideal'):9.0f} "
          f"{top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}  {r[col['Source']][:90]}")
