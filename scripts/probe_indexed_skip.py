"""Timing probe: indexed model + trainskipgrams on a synthetic corpus (not a bench value)."""
import os, sys, time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import colibri_core_b200 as cb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000000
c = cb.Corpus.synthetic(ntokens=n, vocab=1000000, seed=42, mean_sentence=22, phrase_permille=100, nphrases=20000)
for rep in range(3):
    t0 = time.time()
    m = cb.train(c, MINTOKENS=2, MAXLENGTH=5, model_type=20, DOSKIPGRAMS=1, MINSKIPTYPES=2, streamed=0, QUIET=1)
    dt = time.time() - t0
    print("tokens", n, "patterns", len(m), "passes", m.passes(), "wall_ms %.1f" % (dt * 1e3), flush=True)
    print({k: round(v, 2) for k, v in m.timings().items() if v}, flush=True)
    del m
