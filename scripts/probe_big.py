"""1 B-token single-GPU probe + config-3/4 shapes at 100 M (development aid)."""
import json, sys, time
sys.path.insert(0, ".")
import colibri_core_b200 as cb
def run(name, ntok, vocab, **kw):
    c = cb.Corpus.synthetic(int(ntok), vocab=vocab, seed=2 if ntok >= 1e9 else 1)
    best = None
    for i in range(3):
        t0 = time.time(); m = cb.train(c, QUIET=1, **kw); w = time.time() - t0
        r = {"name": name, "wall_ms": round(w * 1e3, 1), "device_ms": round(m.timings()["total"], 1), "patterns": len(m), "tokens": m.tokens(), "Gtok_s": round(m.tokens() / w / 1e9, 2),
             "passes": m.passes(), "peak_GB": round(m.counters()["peak_device_bytes"] / 1e9, 1), "timings": {k: round(v, 1) for k, v in m.timings().items()}}
        m.close()
        if best is None or r["wall_ms"] < best["wall_ms"]: best = r
    print(json.dumps(best), flush=True)
    c.close()
run("zipf-1B unindexed n<=5 t=2", 1e9, 1000000, MINTOKENS=2, MAXLENGTH=5)
run("zipf-100M skipgrams (config 3 shape)", 1e8, 100000, MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0)
run("zipf-100M indexed (config 4 shape)", 1e8, 100000, MINTOKENS=2, MAXLENGTH=5, model_type=20, streamed=0)
run("zipf-1B indexed", 1e9, 1000000, MINTOKENS=2, MAXLENGTH=5, model_type=20, streamed=0)
