"""Per-source-line instruction counts of one kernel from an ncu report (read here, no GPU):
python tools/ncu_lines.py REP.ncu-rep KERNEL_SUBSTR [top]"""
import csv, subprocess, sys, collections, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], stdout=subprocess.PIPE, text=True).stdout
cur_file = cur_fn = None
hdr = None
agg = collections.OrderedDict()
tot = 0
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": cur_fn = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or kern not in (cur_fn or ""): continue
    d = dict(zip(hdr, row))
    if not d["Line No"]: continue
    try:
        n = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except Exception: continue
    key = (cur_file, int(d["Line No"]), d["Source"][:90])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += n; a[1] += smp; a[2] += int(d.get("Thread Instructions Executed", 0) or 0)
    tot += n
print(f"total warp instructions {tot}")
for (f, ln, src), (n, smp, tn) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.1f}%  {n:12d}  thr/inst {tn/max(n,1):5.1f}  smp {smp:6d}  {f}:{ln}  {src}")
