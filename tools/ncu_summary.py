#!/usr/bin/env python
"""Summaries of ncu exports (how profiles/r01_summary.md's tables were produced).

  launches  FILE.csv[.gz]            aggregate a `--metrics gpu__time_duration.sum --csv` launch list per kernel
  traffic   RAW.csv EVENTS           per-kernel DRAM bytes / bytes per row / lanes per instruction of a --set full capture, as JSON
  source    RAW.csv SRC.csv[.gz] [N] key metrics of `ncu -i rep --page raw --csv` plus the N hottest source lines of
                                     `ncu -i rep --page source --csv --print-source cuda,sass`: share of executed
                                     warp instructions, share of stall samples, average active lanes per line

Typical capture on the GPU box (one kernel, one launch; keep the exports, drop the 40 MB report):
  SOFA_KINDS=1 SOFA_ANALYZE=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \\
      -k regex:"k_emit_fused<\\(int\\)1, \\(int\\)1>" --launch-count 1 -o /tmp/p -f python tools/profile_step.py 4000000 1
  ncu -i /tmp/p.ncu-rep --page raw --csv > gpurun_out/raw.csv
  ncu -i /tmp/p.ncu-rep --page source --csv --print-source cuda,sass | gzip > gpurun_out/src.csv.gz
"""
import collections
import csv
import gzip
import io
import re
import sys


def _open(path):
    return io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)


def launches(path):
    rows = list(csv.reader(_open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("sofa::", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("kernel,launches,total_ms,share_pct,avg_ms   # %d launches, %.1f ms" % (sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.3f,%.1f,%.4f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


RAW_KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def source(raw_path, src_path, top=24):
    raw = list(csv.reader(_open(raw_path)))
    h = raw[0]
    for r in raw[2:]:
        print({w: r[h.index(w)][:48] for w in RAW_KEYS if w in h})
    rows = list(csv.reader(_open(src_path)))
    cur = hdr = None
    inst, tinst, stall, text = collections.Counter(), collections.Counter(), collections.Counter(), {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ii, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            si = hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr and cur and len(r) > ti:
            try:
                ln, a, b, c = int(r[0]), int(r[ii] or 0), int(r[ti] or 0), int(r[si] or 0)
            except ValueError:
                continue
            inst[(cur, ln)] += a
            tinst[(cur, ln)] += b
            stall[(cur, ln)] += c
            text[(cur, ln)] = r[1].strip()[:90]
    tot, ts = sum(inst.values()), max(1, sum(stall.values()))
    print("warp instructions %d, average active lanes %.1f" % (tot, sum(tinst.values()) / max(1, tot)))
    for kk, v in inst.most_common(top):
        print("  inst %4.1f%%  stall %4.1f%%  lanes %4.1f  %s:%d  %s" % (100 * v / tot, 100 * stall[kk] / ts, tinst[kk] / max(v, 1),
                                                                       kk[0], kk[1], text[kk]))


def traffic(raw_path, events):
    """`ncu -i rep --page raw --csv` of one `tools/profile_step.py EVENTS 1` capture of the hot kernels ->
    JSON {kernel: dram bytes, bytes per row, duration, lanes per instruction, issue-active %}: what bench.py's
    roofline.traffic is computed from (profiles/r02_ncu_traffic.json).  Rows per source follow bench.py's 60/35/5 mix."""
    import json
    rows_of = {"0": int(events * 0.60), "1": int(events * 0.35), "2": events - int(events * 0.60) - int(events * 0.35)}
    raw = list(csv.reader(_open(raw_path)))
    h = raw[0]
    col = {w: h.index(w) for w in RAW_KEYS if w in h}
    out = {}
    for r in raw[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"(k_parse_tma|k_parse|k_emit_fused)<\(?(?:int\))?(\d)(?:, \(?(?:(?:int|bool)\))?(\w+))?(?:, \(?(?:(?:int|bool)\))?\w+)?>", name)
        if not m:
            continue
        if m.group(1) == "k_parse" and m.group(3) in ("1", "true"):
            continue                        # trace-CSV instantiation
        if m.group(2) not in rows_of:
            continue                        # side-file sources
        key = "%s<%s%s>" % (m.group(1), m.group(2), "," + m.group(3) if (m.group(3) and m.group(1) == "k_emit_fused") else "")
        if key in out:                      # first launch of a kind = the main series (filter series come later)
            continue

        def num(k):
            return float(r[col[k]].replace(",", "")) if k in col and r[col[k]] else None
        units = raw[1]
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
        rd *= scale.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr *= scale.get(units[col["dram__bytes_write.sum"]], 1.0)
        t = num("gpu__time_duration.sum") * {"us": 1e-6, "ns": 1e-9, "ms": 1e-3, "s": 1.0}.get(units[col["gpu__time_duration.sum"]], 1e-6)
        nrows = rows_of[m.group(2)]
        out[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_row": (rd + wr) / nrows, "rows": nrows,
                    "duration_s": t, "dram_GBps": (rd + wr) / t / 1e9,
                    "lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "registers": num("launch__registers_per_thread"), "warp_instructions": num("smsp__inst_executed.sum")}
    print(json.dumps({"capture": "ncu --set full --clock-control none, tools/profile_step.py %d 1 (first launch of each kernel = main series)" % events,
                      "events": events, "kernels": out}, indent=1))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]))
    elif len(sys.argv) >= 4 and sys.argv[1] == "source":
        source(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 24)
    else:
        sys.exit(__doc__)
