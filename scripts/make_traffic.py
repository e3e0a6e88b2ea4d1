#!/usr/bin/env python
"""profiles/traffic.json from ncu captures of the count family (scripts/gpu_final.sh writes them).

    python scripts/make_traffic.py gpurun_out/prof_count_<tag>.ncu-rep [more.ncu-rep ...] --source "<where the capture came from>"

Reads every report with `ncu -i <rep> --page raw --csv`, takes dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum of each
launch (ncu prints a unit row: bytes come as byte / Kbyte / Mbyte / Gbyte), keeps the launches of ONE step (the first occurrence of a kernel
name is level 2 or 3, the next one the level after; gpu_final.sh captures two steps, the first whole one is kept), and ties the sum to the sha of the kernel
sources the capture was taken from (bench.py: kernel_sources_sha) so a later edit of the kernels makes the number stale instead of wrong.
"""
import argparse
import csv
import importlib.util
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def launches_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    res = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue

        def val(metric, table):
            i = col[metric]
            return float(r[i].replace(",", "")) * table[units[i]]

        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
        name = name[5:] if name.startswith("void ") else name
        res.append({"kernel": name, "ms_under_ncu": round(val("gpu__time_duration.sum", TIME), 4),
                    "dram_read_gb": round(val("dram__bytes_read.sum", UNIT) / 1e9, 4), "dram_write_gb": round(val("dram__bytes_write.sum", UNIT) / 1e9, 4)})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--source", required=True)
    ap.add_argument("--kernels", default="count family of one step (levels 2..5); the small scan kernels left out")
    ap.add_argument("--first", default="make_id1_hist", help="kernel that opens a step ('' = keep every captured launch)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    a = ap.parse_args()
    launches = []
    for rep in a.reports:
        launches += launches_of(rep)
    if a.first:  # one whole step: from the first launch of the step's first kernel up to its next launch
        at = [i for i, l in enumerate(launches) if a.first in l["kernel"]]
        if not at:
            sys.exit("no launch of %s in the capture" % a.first)
        launches = launches[at[0] : at[1] if len(at) > 1 else len(launches)]
    seen = {}
    for l in launches:  # label repeated kernels by the level they belong to
        fam = l["kernel"].split("<")[0]
        k = seen[fam] = seen.get(fam, 0) + 1
        if l["kernel"].startswith(("ngram_filter", "count_ngrams", "filter_to_bitmap")):
            l["kernel"] += " L%d" % (k + 2)
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rd = sum(l["dram_read_gb"] for l in launches) * 1e9
    wr = sum(l["dram_write_gb"] for l in launches) * 1e9
    doc = {"source": a.source, "kernels": a.kernels + ": %d launches" % len(launches), "launches": launches, "dram_bytes_read_per_step": rd,
           "dram_bytes_write_per_step": wr, "dram_bytes_per_step": rd + wr, "kernel_sources_sha": bench.kernel_sources_sha()}
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1)
    print("%s: %d launches, %.3f ms under ncu, %.3f GB per step, sha %s" % (a.out, len(launches), sum(l["ms_under_ncu"] for l in launches), (rd + wr) / 1e9, doc["kernel_sources_sha"]), file=sys.stderr)


if __name__ == "__main__":
    main()
