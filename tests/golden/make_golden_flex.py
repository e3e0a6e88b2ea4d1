#!/usr/bin/env python
"""Regenerate tests/golden/golden_flex.json from the UNMODIFIED reference CLI (oracle/_ref/colibri-patternmodeller -s -F S): flexgrams
abstracted from the skipgrams of an indexed model (IndexedPatternModel::computeflexgrams_fromskipgrams, include/patternmodel.h:3724-3744;
SURVEY.md 8(f)-4; reference KAT src/test.cpp:1440-1443: 22 flexgrams, 155 patterns).

The reference inserts the new flexgrams into the unordered_map it is iterating over.  Whenever those insertions trigger a rehash the
running iterator skips some skipgrams and visits others twice (probed: hamlet -l 5 -T 1 yields 66 of 98 flexgrams, 56 of them with
doubled occurrence lists).  That is undefined behaviour of the reference, not a semantics to reproduce: the cases below are the ones where
no rehash strikes (the bucket array is still large from the skipgram candidates), and there the reference equals the clean iteration.
Occurrence lists of flexgrams are compared as multisets (the reference never sorts them; we emit them ascending).
"""
import json
import os
import re
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from make_golden import CORPORA, corpus_body  # noqa: E402

CLI2OPT = {"t": "mintokens", "l": "maxlength", "T": "minskiptypes", "y": "mintokens_skipgrams"}
CASES = [
    ("hamlet", dict()),                      # src/test.cpp:1440-1443: 22 found, 155 patterns
    ("hamlet", dict(l=6, T=3, t=2)),         # no skipgram survives -> no flexgram
    ("skipkat", dict(t=2, l=3, T=1)),
    ("republic", dict(t=2, l=4, T=2)),
    ("republic", dict(t=3, l=5, T=1)),
    ("zipf300k_phr", dict(t=2, l=5, T=2)),
]


def main():
    assert oracle.build_ref(), "reference not built (need /root/reference)"
    out = {"reference": "proycon/colibri-core v2.5.9 (4c07c5a), built by oracle/Makefile", "cases": []}
    with tempfile.TemporaryDirectory() as td:
        for cname, cli in CASES:
            body = corpus_body(CORPORA[cname])
            cpath, mpath = os.path.join(td, cname + ".colibri.dat"), os.path.join(td, "m.patternmodel")
            with open(cpath, "wb") as f:
                f.write(b"\xa2\x02" + body)
            args = ["-f", cpath, "-s", "-F", "S", "-o", mpath]
            for k, v in cli.items():
                args += ["-" + k, v]
            rc, err = oracle.ref_cli(args)
            assert rc == 0, err
            ref = oracle.parse_modelfile(open(mpath, "rb").read()).sorted_refs()
            found = int(re.search(r"(\d+) flexgrams found", err).group(1))
            case = {"corpus": cname, "cli": cli, "options": dict({CLI2OPT[k]: v for k, v in cli.items()}, indexed=1, doskipgrams=1, streamed=0), "flexfound": found,
                    "patterns": len(ref), "tokens": ref.tokens, "types": ref.types, "occurrences": int(ref.counts.sum()), "digest_sorted_refs": ref.digest()}
            out["cases"].append(case)
            print(cname, cli, found, len(ref), case["digest_sorted_refs"][:12])
    with open(os.path.join(HERE, "golden_flex.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
