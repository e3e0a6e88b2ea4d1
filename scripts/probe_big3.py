"""1 B-token two-stage probe, unindexed only (development aid)."""
import json
import sys
import time

sys.path.insert(0, ".")
import colibri_core_b200 as cb

c = cb.Corpus.synthetic(int(1e9), vocab=1000000, seed=2)
stage1 = cb.train(c, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
keys, off, counts, _ = stage1.export()
loaded = cb.Model.from_flat(keys, off, None, tokens=stage1.tokens(), types=stage1.types())
stage1.close()
for i in range(2):
    t0 = time.time()
    m = cb.train_constrained(c, loaded, inplace=True, MINTOKENS=2, MAXLENGTH=5, streamed=0, QUIET=1)
    w = time.time() - t0
    print(json.dumps({"wall_ms": round(w * 1e3, 1), "device_ms": round(m.timings()["total"], 1), "patterns": len(m), "levels": {n: round(m.level(n)["count_ms"], 1) for n in range(1, 6)},
                      "timings": {k: round(v, 1) for k, v in m.timings().items()}}), flush=True)
    m.close()
