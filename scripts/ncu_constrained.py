"""One two-stage run for profiling: stage 1 (unindexed train), model load (DORESET), constrained in-place rebuild."""
import sys

sys.path.insert(0, ".")
import colibri_core_b200 as cb

ntok = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
corpus = cb.Corpus.synthetic(ntok, vocab=100000, seed=1)
stage1 = cb.train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
loaded = cb.load_model(stage1.to_bytes(), MINTOKENS=2, MAXLENGTH=5, DORESET=1, QUIET=1)
for _ in range(2):
    m = cb.train_constrained(corpus, loaded, inplace=True, MINTOKENS=2, MAXLENGTH=5, streamed=0, QUIET=1)
    print(len(m), m.timings())
    m.close()
