"""Top stall-sample SASS instructions per kernel from an ncu report (source page, sass view).
Usage: python tools/ncu_hot_sass.py report.ncu-rep [kernel substring] [top n]"""
import csv, subprocess, sys
path = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; hdr = None; continue
    if kern and hdr is None:
        hdr = r; data[kern].append(r); continue
    if kern: data[kern].append(r)
for k, v in data.items():
    if filt not in k: continue
    hdr = v[0]; i_src = hdr.index("Source"); i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_x = hdr.index("Instructions Executed")
    rs = [(int(r[i_s] or 0), idx, r[i_src].strip(), r[i_x]) for idx, r in enumerate(v[1:]) if len(r) > i_s and (r[i_s] or "0").isdigit()]
    tot = sum(x[0] for x in rs)
    print("==", k[:100], "instructions", len(rs), "total samples", tot)
    for s_, idx, src, ex in sorted(rs, reverse=True)[:topn]:
        print(f"  {100 * s_ / max(tot, 1):5.1f}%  #{idx:5d} x{ex:>8s}  {src[:100]}")
