"""Group tools/ncu_lines.py output by code region: python tools/ncu_regions.py REP KERNEL FILE 'name:lo-hi,name:lo-hi,...' [steps]"""
import re, subprocess, sys
rep, kern, fname, spec = sys.argv[1:5]
steps = float(sys.argv[5]) if len(sys.argv) > 5 else None
regs = []
for part in spec.split(","):
    n, r = part.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, kern, "100000"], stdout=subprocess.PIPE, text=True).stdout
agg = {}; tot = 0
for line in out.splitlines()[1:]:
    m = re.match(r"\s*([\d.]+)%\s+(\d+)\s+thr/inst\s+([\d.]+)\s+smp\s+(\d+)\s+(\S+):(\d+)", line)
    if not m: continue
    n = int(m.group(2)); f = m.group(5); ln = int(m.group(6)); smp = int(m.group(4))
    k = f
    if f == fname:
        k = "other " + fname
        for name, lo, hi in regs:
            if lo <= ln <= hi: k = name; break
    a = agg.setdefault(k, [0, 0]); a[0] += n; a[1] += smp; tot += n
print("total", tot, (f"= {tot/steps:.0f} per step" if steps else ""))
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {100*n/tot:5.1f}%  " + (f"{n/steps:6.0f}/step  " if steps else "") + f"smp {s}")
