"""Top CUDA source lines by warp-stall samples per kernel from an ncu report captured with --import-source on (-lineinfo build).
Usage: python tools/ncu_hot_lines.py report.ncu-rep [kernel substring] [top n]"""
import csv
import subprocess
import sys

path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kern, hdr, rows = None, None, {}


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kern, hdr = r[1], None
        rows.setdefault(kern, [])
        continue
    if r[0] == "Line No" and kern is not None and len(r) > 4:
        hdr = r
        continue
    if kern is not None and hdr is not None and r[0].isdigit():
        r[1] = f"[{fpath[:14]}] " + r[1].strip()
        rows[kern].append(r)
for k, v in rows.items():
    if filt not in k or not v:
        continue
    i_s = hdr.index("Warp Stall Sampling (All Samples)")
    i_x = hdr.index("Instructions Executed")
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[i_s]) for r in v)
    totx = sum(num(r[i_x]) for r in v)
    print("==", k[:110], "| samples", tot, "| warp instructions", totx)
    for r in sorted(v, key=lambda r: -num(r[i_s]))[:topn]:
        stalls = sorted(((num(r[hdr.index(n)]), n[6:]) for n in names), reverse=True)[:3]
        why = " ".join(f"{n}:{c}" for c, n in stalls if c)
        print(f"  {100 * num(r[i_s]) / max(tot, 1):5.1f}%  x{100 * num(r[i_x]) / max(totx, 1):4.1f}%  L{r[0]:>4s}  {r[1].strip()[:90]:90s} {why}")
