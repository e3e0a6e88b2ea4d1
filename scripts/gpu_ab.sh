# A/B of tuning knobs on the bench corpus: `VARIANTS="A=1 B=2|C=3" bash scripts/gpu_ab.sh` (variants separated by |, each a list of env assignments)
mkdir -p gpurun_out
T=${TAG:-ab}
Q="--no-cpu-baseline --no-e2e --no-extra --no-digest ${BENCH_ARGS:-}"
IFS='|' read -ra VS <<< "${VARIANTS:-_=0}"
for V in "${VS[@]}"; do
  echo "== $V"; env $V timeout 200 python bench.py $Q 2> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phase_ms_per_step'].items()}, {k:(round(v['count_ms'],3),v['items'],v['singletons'],v['capacity']) for k,v in d['levels_last_step'].items()}, d['parity'].get('checksum_ok'))"
done 2>&1 | tee gpurun_out/ab_$T.log
