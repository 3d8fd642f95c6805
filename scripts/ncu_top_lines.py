"""Summarise an ncu `--page source --csv` export: top source lines / SASS by warp-stall samples.
python scripts/ncu_top_lines.py <source.csv> [n]"""
import csv
import sys
from collections import defaultdict

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path, newline="")))
hdr = None
for i, r in enumerate(rows):
    if any("Sampling" in c for c in r):
        hdr, start = r, i + 1
        break
if hdr is None:
    print("no header found; first rows:", rows[:3])
    sys.exit(0)
print("columns:", hdr)
col = {c: i for i, c in enumerate(hdr)}
samp = next((c for c in hdr if c.startswith("# Samples") or "Warp Stall Sampling (All" in c), None)
src = next((c for c in hdr if c == "Source"), None)
print("sampling column:", samp)
tot = 0
items = []
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[col[samp]].replace(",", ""))
    except ValueError:
        continue
    tot += v
    items.append((v, r))
items.sort(key=lambda t: -t[0])
for v, r in items[:n]:
    extra = " | ".join(f"{c}={r[col[c]]}" for c in hdr if c not in (samp, src) and ("Stall" in c or c in ("Address", "Instructions Executed", "Warp Stall Sampling (Not-issued Cycles)")) and r[col[c]] not in ("0", ""))
    print(f"{v:8.0f} {100 * v / max(tot, 1):5.1f}%  {r[col[src]][:110] if src else ''}   {extra[:200]}")
