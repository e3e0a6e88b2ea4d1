#!/usr/bin/env python
"""Regenerate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built by `make -C oracle ref`
out of /root/reference).  Runs only in the development container; the JSON and the two small corpus fixtures
are committed so that the tests need neither /root/reference nor oracle/_ref.

Fixtures:
  hamlet.colibri.dat/.cls  -- written by the reference's own colibri-test (src/test.cpp:1165-1175: the poem at
                              :57-96 through colibri-classencode); 354 tokens, 40 sentences.
  republic.colibri.dat     -- colibri-classencode on the reference's exp/republic.txt (251 527 tokens).
Everything else is generated (quirk corpora inline, synthetic corpora from oracle.synth_corpus parameters).
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

CLI2OPT = {"t": "mintokens", "l": "maxlength", "m": "minlength", "b": "maxbackofflength", "y": "mintokens_skipgrams", "T": "minskiptypes", "W": "mintokens_unigrams"}

# corpus specs: ("file", name) | ("hex", hexbytes-of-body) | ("synth", {params})
CORPORA = {
    "hamlet": ("file", "hamlet.colibri.dat"),
    "republic": ("file", "republic.colibri.dat"),
    "noeos": ("hex", bytes([6, 7, 8, 0, 6, 7, 8, 0, 6, 7, 9]).hex()),
    "noeos_multibyte": ("hex", bytes([6, 7, 0xAC, 2, 0, 6, 7, 0xAC, 2, 0, 6, 2, 6, 7, 0xAC, 2]).hex()),
    "empty_sentences": ("hex", bytes([0, 0, 6, 7, 8, 0, 0, 6, 7, 8, 0, 0, 0, 9, 0]).hex()),
    "single_token": ("hex", bytes([6, 0]).hex()),
    "only_delims": ("hex", bytes([0, 0, 0]).hex()),
    "unknown_class": ("hex", bytes([2, 6, 2, 0, 2, 6, 2, 0, 7, 2, 6, 0]).hex()),
    # SURVEY 8a tiny skipgram KAT: a b c / a d c / a b / b c / a d / d c
    "skipkat": ("hex", bytes([6, 7, 8, 0, 6, 9, 8, 0, 6, 7, 0, 7, 8, 0, 6, 9, 0, 9, 8, 0]).hex()),
    "threebyte": ("hex", oracle.encode_corpus([[6, 300, 20000, 2097151, 70000], [6, 300, 20000, 2097151, 70000], [300, 20000, 5000000, 6], [300, 20000, 5000000, 6]]).hex()),
    "zipf200k": ("synth", dict(ntokens=200000, vocab=5000, seed=1, mean_sentence=22)),
    "zipf300k_phr": ("synth", dict(ntokens=300000, vocab=20000, seed=7, mean_sentence=15, phrase_permille=300, nphrases=2000)),
    "zipf2m": ("synth", dict(ntokens=2000000, vocab=100000, seed=1, mean_sentence=22)),
}

# (corpus, unindexed, skipgrams, {cli options})
CASES = [
    ("hamlet", True, False, dict(t=2, l=3)),       # BASELINE.json configs[0]
    ("hamlet", True, False, dict(t=2, l=5)),
    ("hamlet", True, False, dict()),               # src/test.cpp:1224-1232 (111/186/354)
    ("hamlet", True, True, dict()),                # src/test.cpp:1261-1283 (385)
    ("hamlet", True, True, dict(l=5, T=1)),
    ("hamlet", True, True, dict(l=6, y=3)),
    ("hamlet", False, False, dict(l=5)),
    ("hamlet", False, False, dict()),
    ("hamlet", True, False, dict(t=1, l=4)),
    ("hamlet", True, False, dict(t=3, l=8, m=2)),
    ("hamlet", True, False, dict(t=2, l=8, m=3)),
    ("hamlet", True, False, dict(t=2, W=4, l=6)),
    ("hamlet", True, False, dict(t=2, b=2, l=6)),
    ("hamlet", True, True, dict(t=1, l=4)),
    ("hamlet", False, False, dict(t=1, l=3)),
    ("noeos", True, False, dict(t=1, l=3)),
    ("noeos", True, False, dict(t=2, l=3)),
    ("noeos", False, False, dict(t=1, l=3)),
    ("noeos_multibyte", True, False, dict(t=1, l=3)),
    ("noeos_multibyte", True, False, dict(t=2, l=3)),
    ("noeos_multibyte", False, False, dict(t=2, l=3)),
    ("empty_sentences", True, False, dict(t=2, l=3)),
    ("empty_sentences", False, False, dict(t=2, l=3)),
    ("single_token", True, False, dict(t=1, l=3)),
    ("single_token", True, False, dict(t=2, l=3)),
    ("only_delims", True, False, dict(t=2, l=3)),
    ("unknown_class", True, False, dict(t=2, l=3)),
    ("skipkat", True, True, dict(t=2, l=3)),
    ("skipkat", True, True, dict(t=2, l=3, y=3, T=1)),
    ("skipkat", True, True, dict(t=2, l=3, y=3, T=2)),
    ("threebyte", True, False, dict(t=2, l=5)),
    ("threebyte", True, True, dict(t=2, l=5)),
    ("threebyte", False, False, dict(t=2, l=5)),
    ("republic", True, False, dict(t=2, l=5)),
    ("republic", True, True, dict(t=2, l=5)),
    ("republic", False, False, dict(t=2, l=5)),
    ("republic", True, True, dict(t=3, l=4, y=5)),
    ("republic", True, False, dict(t=2, l=8)),
    ("zipf200k", True, False, dict(t=2, l=5)),
    ("zipf200k", True, True, dict(t=2, l=5)),
    ("zipf200k", False, False, dict(t=2, l=5)),
    ("zipf300k_phr", True, False, dict(t=2, l=5)),
    ("zipf300k_phr", True, True, dict(t=2, l=5)),
    ("zipf300k_phr", False, False, dict(t=2, l=6)),
    ("zipf300k_phr", True, False, dict(t=5, l=7)),
    ("zipf2m", True, False, dict(t=2, l=5)),
    # indexed models with skipgrams from n-grams (IndexedPatternModel::trainskipgrams, skip-type pruning)
    ("hamlet", False, True, dict()),               # src/test.cpp:1321-1337, test.py:289-291 (133 patterns)
    ("hamlet", False, True, dict(l=5, T=1)),
    ("hamlet", False, True, dict(l=5, T=2)),
    ("hamlet", False, True, dict(l=6, T=3, t=2)),
    ("skipkat", False, True, dict(t=2, l=3, T=1)),
    # NOTE: ("threebyte", indexed, skipgrams) is deliberately NOT a golden case.  trainskipgrams() inserts skipgrams into the
    # std::unordered_map it is iterating over (patternmodel.h:2986-2990); on that 21-pattern model the insertions of the n=4
    # loop trigger a rehash (29 buckets), libstdc++ relinks the node list, and the running iterator never reaches one of the
    # three 4-grams: the reference emits 6 instead of 9 4-skipgrams.  That is undefined behaviour of the reference, not a
    # semantics to reproduce; the cases below (and the reference's own KAT, 133 patterns) are free of it.
    ("republic", False, True, dict(t=2, l=4, T=2)),
    ("republic", False, True, dict(t=3, l=5, T=1)),
    ("zipf300k_phr", False, True, dict(t=2, l=5, T=2)),
]


def corpus_body(spec):
    kind, arg = spec
    if kind == "file":
        return open(os.path.join(HERE, arg), "rb").read()[2:]
    if kind == "hex":
        return bytes.fromhex(arg)
    return oracle.synth_corpus(**arg).tobytes()


def kat_messages():
    """Deterministic messages of every length 0..191 (the SpookyV2 Short range) plus pattern-like ones."""
    msgs = []
    x = 0x243F6A8885A308D3
    for n in range(0, 192):
        b = bytearray()
        for _ in range(n):
            x = (x * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            b.append((x >> 33) & 0xFF)
        msgs.append(bytes(b))
    msgs += [bytes([6]), bytes([6, 7]), bytes.fromhex("860107808001"), bytes(range(1, 16)), bytes(range(1, 17)), bytes.fromhex("0a03030d0e")]
    return msgs


def main():
    assert oracle.build_ref(), "reference not built (need /root/reference)"
    out = {"reference": "proycon/colibri-core v2.5.9 (4c07c5a), built by oracle/Makefile", "corpora": {}, "cases": [], "spooky": [], "masks": {}}
    for name, spec in CORPORA.items():
        out["corpora"][name] = {"kind": spec[0], "arg": spec[1]}
    msgs = kat_messages()
    nonempty = [m for m in msgs if len(m) > 0]
    r = subprocess.run([oracle.REF_TRAIN, "--spooky"] + [m.hex() for m in nonempty], capture_output=True, text=True, check=True)
    hashes = [int(x) for x in r.stdout.split()]
    assert len(hashes) == len(nonempty)
    out["spooky"] = [[m.hex(), h] for m, h in zip(nonempty, hashes)]
    for n in range(3, 10):
        for ms in (1, 2, 3):
            r = subprocess.run([oracle.REF_TRAIN, "--masks", str(n), str(ms)], capture_output=True, text=True, check=True)
            out["masks"]["%d,%d" % (n, ms)] = [int(x) for x in r.stdout.split()]
    with tempfile.TemporaryDirectory() as td:
        for cname, unindexed, skipgrams, cli in CASES:
            body = corpus_body(CORPORA[cname])
            cpath = os.path.join(td, cname + ".colibri.dat")
            with open(cpath, "wb") as f:
                f.write(b"\xa2\x02" + body)
            mpath = os.path.join(td, "m.patternmodel")
            st, err = oracle.ref_train(cpath, mpath, unindexed=unindexed, skipgrams=skipgrams, **cli)
            ref = oracle.parse_modelfile(open(mpath, "rb").read())
            opts = {CLI2OPT[k]: v for k, v in cli.items()}
            opts["indexed"] = 0 if unindexed else 1
            opts["doskipgrams_exhaustive"] = 1 if (skipgrams and unindexed) else 0
            opts["doskipgrams"] = 1 if (skipgrams and not unindexed) else 0
            # the CLI streams the file only for unindexed models without skipgrams (src/patternmodeller.cpp:721-754)
            opts["streamed"] = 1 if (unindexed and not skipgrams) else 0
            case = {
                "corpus": cname, "options": opts, "cli": cli, "unindexed": unindexed, "skipgrams": skipgrams,
                "tokens": st["tokens"], "types": st["types"], "patterns": st["patterns"], "maxn": st["maxn"], "minn": st["minn"], "hasskipgrams": st["hasskipgrams"],
                "passes": [list(p) for p in oracle.parse_ref_passes(err)], "occurrences": int(ref.counts.sum()), "digest": ref.digest(),
            }
            if len(ref) <= 120:
                c = ref.canonical()
                case["model"] = [[c.key(i).hex(), int(c.counts[i])] + ([c.refs(i)] if c.ref_off is not None else []) for i in range(len(c))]
            out["cases"].append(case)
            print(cname, cli, "u" if unindexed else "i", "s" if skipgrams else "-", st["patterns"], case["digest"][:12])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
