"""Print the hottest SASS lines (warp-stall samples) of each kernel in an .ncu-rep:
   python scripts/ncu_source.py prof.ncu-rep [kernel-substring] [top N]"""
import csv, subprocess, sys
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
for b in blocks:
    if want not in b["name"]:
        continue
    h = b["hdr"]
    i_src, i_all, i_ni, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Warp Stall Sampling (Not-issued Samples)"), h.index("Instructions Executed")
    tot = sum(int(r[i_all]) for r in b["rows"])
    print(f"== {b['name'][:90]}  samples={tot} sass_lines={len(b['rows'])}")
    idx = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][i_all]))[:top]
    for i in sorted(idx):
        r = b["rows"][i]
        print(f"{i:5d} {100*int(r[i_all])/max(tot,1):5.1f}% ni={int(r[i_ni]):6d} ex={int(r[i_ex]):9d}  {r[i_src].strip()[:110]}")
