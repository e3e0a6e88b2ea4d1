import sys, time
sys.path.insert(0, ".")
import colibri_core_b200 as cb
c = cb.Corpus.synthetic(100_000_000, vocab=100000, seed=1)
last = None
for i in range(8):
    t0 = time.time()
    m = cb.train(c, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
    w = time.time() - t0
    tm = m.timings()
    print(i, "wall %.1f total %.1f export %.2f" % (w * 1e3, tm["total"], tm["export"]), "peak", m.counters()["peak_device_bytes"] >> 20, flush=True)
    if last is not None:
        last.close()
    last = m
