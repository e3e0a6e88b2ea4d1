# partition-path iteration (development aid): parity of the part paths, bench line, per-kernel times, optional full ncu capture
mkdir -p gpurun_out
T=${TAG:-partx}
Q="--no-cpu-baseline --no-e2e --no-extra"
if [ -z "$NOTEST" ]; then
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "part" > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log | cut -c1-300
fi
timeout 300 python bench.py $Q > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$T.json"))
print(round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phase_ms_per_step"].items()})
print({k:(round(v["count_ms"],3),v["capacity"]) for k,v in d["levels_last_step"].items()}, d["parity"].get("digest_ok"), d["parity"].get("checksum_ok"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"part_|prune_dense|compact" -s ${LSKIP:-36} -c ${LCOUNT:-36} --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 $Q --no-digest > gpurun_out/ncu_$T.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${T}_launches.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; kn=hdr.index("Kernel Name"); mn=hdr.index("Metric Name"); mv=hdr.index("Metric Value"); idc=hdr.index("ID")
cur={}
for r in rows[h+1:]:
    if len(r)>mv: cur.setdefault((r[idc], r[kn].split("(")[0][-28:]),{})[r[mn]]=r[mv]
for k,v in cur.items():
    print("%-30s %8.1f us  R %7.1f MB  W %7.1f MB" % (k[1], float(v["gpu__time_duration.sum"])/1e3, float(v["dram__bytes_read.sum"])/1e6, float(v["dram__bytes_write.sum"])/1e6))
PY
if [ -n "$FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"part_hist|part_split|part_count" -s ${FSKIP:-9} -c ${FCOUNT:-8} -f -o gpurun_out/prof_$T python bench.py --steps 1 --warmup 1 $Q --no-digest > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log | cut -c1-200
fi
