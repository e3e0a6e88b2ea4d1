"""Device timing probe (development aid): two-stage building on a synthetic corpus -- stage 1 unindexed train(), model file
round trip through colibri_b200_model_load (DORESET), constrained in-place rebuild (unindexed and indexed), batch lookups."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import colibri_core_b200 as cb

ntok = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
vocab = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
corpus = cb.Corpus.synthetic(ntok, vocab=vocab, seed=1)
stage1 = cb.train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
blob = stage1.to_bytes()
print(json.dumps({"stage1_patterns": len(stage1), "file_bytes": len(blob)}), flush=True)
for indexed in (0, 1):
    for it in range(3):
        t0 = time.time()
        loaded = cb.load_model(blob, MINTOKENS=2, MAXLENGTH=5, DORESET=1, model_type=20 if indexed else 10, QUIET=1)
        t1 = time.time()
        m = cb.train_constrained(corpus, loaded, inplace=True, MINTOKENS=2, MAXLENGTH=5, model_type=20 if indexed else 10, streamed=0, QUIET=1)
        t2 = time.time()
        tm = m.timings()
        print(json.dumps({"indexed": indexed, "iter": it, "load_ms": round((t1 - t0) * 1e3, 1), "train_wall_ms": round((t2 - t1) * 1e3, 1), "patterns": len(m),
                          "Mtok_s_device": round(m.tokens() / max(tm["total"], 1e-9) / 1e3, 1), "timings_ms": {k: round(v, 3) for k, v in tm.items()},
                          "levels_ms": {n: round(m.level(n)["count_ms"], 3) for n in range(1, 6)}, "counters": m.counters()}), flush=True)
        m.close()
        loaded.close()
keys, off, counts, _ = stage1.export()
rng = np.random.default_rng(1)
pick = rng.integers(0, len(counts), 1_000_000)
q = [keys[int(off[i]):int(off[i + 1])].tobytes() for i in pick]
for it in range(3):
    t0 = time.time()
    c, idx = stage1.lookup_batch(q)
    dt = time.time() - t0
    assert np.array_equal(c, counts[pick]) and np.array_equal(idx, pick)
    print(json.dumps({"lookup_batch": len(q), "wall_ms": round(dt * 1e3, 1)}), flush=True)
