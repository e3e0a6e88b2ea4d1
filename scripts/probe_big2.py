"""1 B-token single-GPU probes (development aid): unindexed training, then two-stage (load with DORESET + constrained in-place rebuild)."""
import json
import sys
import time

sys.path.insert(0, ".")
import colibri_core_b200 as cb

c = cb.Corpus.synthetic(int(1e9), vocab=1000000, seed=2)
best = None
stage1 = None
for i in range(3):
    t0 = time.time()
    m = cb.train(c, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
    w = time.time() - t0
    r = {"name": "zipf-1B unindexed n<=5 t=2", "wall_ms": round(w * 1e3, 1), "device_ms": round(m.timings()["total"], 1), "patterns": len(m), "Gtok_s": round(m.tokens() / w / 1e9, 2),
         "peak_GB": round(m.counters()["peak_device_bytes"] / 1e9, 1), "timings": {k: round(v, 1) for k, v in m.timings().items()},
         "levels": {n: round(m.level(n)["count_ms"], 1) for n in range(2, 6)}}
    if best is None or r["wall_ms"] < best["wall_ms"]:
        best = r
    if stage1 is not None:
        stage1.close()
    stage1 = m
print(json.dumps(best), flush=True)
keys, off, counts, _ = stage1.export()
t0 = time.time()
loaded = cb.Model.from_flat(keys, off, None, tokens=stage1.tokens(), types=stage1.types())
print(json.dumps({"upload_ms": round((time.time() - t0) * 1e3, 1), "patterns": len(loaded)}), flush=True)
stage1.close()
for indexed in (0, 1):
    best = None
    for i in range(3):
        t0 = time.time()
        m = cb.train_constrained(c, loaded, inplace=True, MINTOKENS=2, MAXLENGTH=5, model_type=20 if indexed else 10, streamed=0, QUIET=1)
        w = time.time() - t0
        r = {"name": "zipf-1B constrained in-place rebuild, indexed=%d" % indexed, "wall_ms": round(w * 1e3, 1), "device_ms": round(m.timings()["total"], 1), "patterns": len(m),
             "Gtok_s": round(m.tokens() / w / 1e9, 2), "peak_GB": round(m.counters()["peak_device_bytes"] / 1e9, 1), "timings": {k: round(v, 1) for k, v in m.timings().items()},
             "levels": {n: round(m.level(n)["count_ms"], 1) for n in range(1, 6)}}
        m.close()
        if best is None or r["wall_ms"] < best["wall_ms"]:
            best = r
    print(json.dumps(best), flush=True)
