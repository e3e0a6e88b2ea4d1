# The C++ drop-in, file -> file, on the bench corpus (100 M tokens): `colibri-patternmodeller -f corpus -u -t 2 -l 5 -o model` with the host timing
# line the CLI prints (read files / device calls / result to host / map view / write).  Results: gpurun_out/cli_timing.log
mkdir -p gpurun_out /tmp/cli
python - <<'PY'
import colibri_core_b200 as cb
c = cb.Corpus.synthetic(100000000, vocab=100000, seed=1, mean_sentence=22)
body = c.download()
with open("/tmp/cli/zipf100m.colibri.dat", "wb") as f:
    f.write(b"\xa2\x02")
    f.write(body.tobytes())
print("corpus bytes", len(body) + 2)
PY
TIMEFORMAT="wall %R s (user %U, sys %S)"
for i in 1 2 3; do
  { time colibri-core_b200/bin/colibri-patternmodeller -f /tmp/cli/zipf100m.colibri.dat -u -t 2 -l 5 -o /tmp/cli/model.$i 2>&1 | grep -v "^Counting\|^ Found\|^Training pattern"; } 2>&1
done > gpurun_out/cli_timing.log 2>&1
ls -l /tmp/cli/model.1 >> gpurun_out/cli_timing.log
python - <<'PY' >> gpurun_out/cli_timing.log 2>&1
import hashlib, sys
sys.path.insert(0, ".")
import oracle
m = oracle.parse_modelfile(open("/tmp/cli/model.1", "rb").read())
print("patterns", len(m), "digest", m.digest())
PY
cat gpurun_out/cli_timing.log
