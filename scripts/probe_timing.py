"""Quick device timing probe (development aid): phase timings of train_corpus on a synthetic corpus."""
import json
import sys
import time

sys.path.insert(0, ".")
import colibri_core_b200 as cb

ntok = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
vocab = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
corpus = cb.Corpus.synthetic(ntok, vocab=vocab, seed=1)
print("synth %.2fs, %d bytes" % (time.time() - t0, corpus.nbytes), flush=True)
for it in range(4):
    t0 = time.time()
    m = cb.train(corpus, MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=skip, streamed=0 if skip else 1, QUIET=1)
    wall = time.time() - t0
    tm = m.timings()
    print(json.dumps({"iter": it, "wall_ms": round(wall * 1e3, 2), "patterns": len(m), "tokens": m.tokens(), "Mtok_s": round(m.tokens() / wall / 1e6, 1),
                      "timings_ms": {k: round(v, 3) for k, v in tm.items()}, "counters": m.counters(), "passes": m.passes(),
                      "levels": {n: m.level(n) for n in range(2, m.maxlength() + 1)}}), flush=True)
    m.close()
