#!/usr/bin/env python
"""Pin the BENCHMARK corpus (BASELINE.json configs[1]) to the UNMODIFIED reference: tests/golden/golden_bench.json.

The corpus is the one bench.py times: oracle.synth_corpus(1e8, vocab=100000, seed=1, mean_sentence=22) (the counter-based
integer generator; Corpus.synthetic on the device produces the identical bytes, see test_device_generator_is_bit_identical).
The reference (oracle/_ref/ref_train = PatternModel<uint32_t>::train + write of /root/reference, compiled by `make -C oracle ref`)
runs `-u -t 2 -l 5` on it: ~12 minutes on one core and ~3 GB for the 100 M-token case, so this script runs only in the development
container and its output is committed.  Also pins two shorter prefixes of the same stream (cheap enough for the oracle at test time
too) and the exhaustive-skipgram variant (config 3 shape) on the 10 M prefix.

What is recorded per case: tokens, types, pattern count, per-pass (found, pruned), per-length (patterns, sum of counts), the size of
the written model file, the canonical digest (oracle.FlatModel.digest: sha256 over the bytewise-sorted (key, count) stream) and the
order-independent checksum (tests/checksum.py = colibri_b200_model_checksum) that bench.py compares at every GPU count.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from checksum import flat_checksum  # noqa: E402

CASES = [
    # name, ntokens, vocab, seed, skipgrams
    ("zipf10m", 10_000_000, 100000, 1, False),
    ("zipf10m_skip", 10_000_000, 100000, 1, True),
    ("zipf100m", 100_000_000, 100000, 1, False),
]


def per_length(fm):
    lens = np.diff(fm.key_off.astype(np.int64))
    off = fm.key_off.astype(np.int64)
    # tokens of a key = bytes < 128 (skip markers 0x03 included)
    small = (fm.keys[: int(off[-1])] < 128).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(small)])
    ntok = csum[off[1:]] - csum[off[:-1]]
    out = {}
    for n in np.unique(ntok):
        sel = ntok == n
        out[int(n)] = [int(sel.sum()), int(fm.counts[sel].astype(np.int64).sum())]
    del lens
    return out


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/ref_train missing: run `make -C oracle ref` (needs /root/reference)")
    only = set(sys.argv[1:])
    path_out = os.path.join(HERE, "golden_bench.json")
    result = json.load(open(path_out)) if os.path.exists(path_out) else {}
    for name, ntok, vocab, seed, skip in CASES:
        if only and name not in only:
            continue
        body = oracle.synth_corpus(ntok, vocab=vocab, seed=seed, mean_sentence=22)
        with tempfile.TemporaryDirectory() as td:
            cpath, mpath = os.path.join(td, "c.colibri.dat"), os.path.join(td, "m.colibri.patternmodel")
            with open(cpath, "wb") as f:
                f.write(b"\xa2\x02")
                f.write(body.tobytes())
            t0 = time.time()
            st, err = oracle.ref_train(cpath, mpath, unindexed=True, skipgrams=skip, t=2, l=5)
            wall = time.time() - t0
            blob = open(mpath, "rb").read()
        fm = oracle.parse_modelfile(blob)
        passes = oracle.parse_ref_passes(err)
        result[name] = {
            "generator": {"ntokens": ntok, "vocab": vocab, "seed": seed, "mean_sentence": 22}, "corpus_bytes": int(body.size),
            "cli": "-u -t 2 -l 5" + (" -s" if skip else ""), "tokens": int(fm.tokens), "types": int(fm.types), "patterns": len(fm),
            "passes_found_skip_pruned_kept": passes, "per_length_patterns_occurrences": per_length(fm), "modelfile_bytes": len(blob),
            "digest": fm.digest(), "checksum": flat_checksum(fm), "reference_train_seconds": st.get("train_seconds"), "reference_wall_seconds": round(wall, 1),
            "host": "development container, 1 core (the reference is single-threaded)",
        }
        print(name, json.dumps(result[name])[:400], flush=True)
        with open(path_out, "w") as f:
            json.dump(result, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
