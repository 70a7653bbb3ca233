"""DRAM traffic per pipeline stage from an `ncu --set full` report of one bench.py run (read here, no GPU):

    python tools/ncu_traffic.py REP.ncu-rep OUT.json "source note"

Takes the LAST complete step in the report (a step starts with the split kernel and ends with the merge), sums
dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum over the launches of every stage, and writes
what bench.py reads as profiles/traffic.json (`roofline.traffic`)."""
import csv
import json
import subprocess
import sys

STAGES = [("split", ("fz_split",)), ("encode", ("fz_hist2", "fz_group_code", "fz_emit2", "fz_zero_hist2")),
          ("layout", ("fz_layout",)), ("gather", ("fz_gather",)), ("walk", ("fz_walk",)), ("markers", ("fz_marker",)),
          ("classify", ("fz_classify",)), ("inflate_fast", ("fz_inflate_prep", "fz_inflate_lean", "fz_inflate_group")),
          ("inflate_blockpar", ("fz_bp_",)), ("inflate_general", ("fz_inflate_general",)), ("merge", ("fz_merge",)),
          ("rawcopy", ("fz_rawcopy",))]


def stage_of(name: str):
    for st, keys in STAGES:
        if any(k in name for k in keys):
            return st
    return None


def to_bytes(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_ms(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    return x * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[unit]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        launches.append((name, to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) +
                         to_bytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"]),
                         to_ms(d["gpu__time_duration.sum"], units["gpu__time_duration.sum"])))
    starts = [i for i, (n, _, _) in enumerate(launches) if "fz_split_kernel" in n]
    ends = [i for i, (n, _, _) in enumerate(launches) if "fz_merge" in n]
    if not starts or not ends:
        sys.exit("no complete step in the report")
    end = ends[-1]
    start = max(i for i in starts if i < end)
    traffic, ms, kernels = {}, {}, []
    for name, b, t in launches[start:end + 1]:
        st = stage_of(name)
        if st is None:
            continue
        traffic[st] = traffic.get(st, 0.0) + b
        ms[st] = ms.get(st, 0.0) + t
        kernels.append({"kernel": name, "stage": st, "dram_bytes": b, "ms_under_ncu": round(t, 4)})
    res = dict(traffic)
    res["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per step (every launch of the stage's kernels in the last "
                    "step of the report), ncu --set full --clock-control none")
    res["_source"] = note
    res["_ms_under_ncu"] = {k: round(v, 4) for k, v in ms.items()}
    res["_kernels"] = kernels
    json.dump(res, open(out, "w"), indent=1)
    for k in traffic:
        print(f"{k:18s} {traffic[k] / 1e9:8.3f} GB   {ms[k]:7.3f} ms under ncu")


if __name__ == "__main__":
    main()
