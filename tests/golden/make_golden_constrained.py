#!/usr/bin/env python
"""Regenerate tests/golden/golden_constrained.json from the UNMODIFIED reference CLI (oracle/_ref/colibri-patternmodeller,
built by `make -C oracle ref` out of /root/reference): training under a constraint model, SURVEY.md 8(f)-2.

  mode "j": colibri-patternmodeller -f CORPUS -j STAGE1 [-u] -t .. -l .. -m .. -o OUT      (src/patternmodeller.cpp:717-721, :316-319)
  mode "I": colibri-patternmodeller -f CORPUS -i STAGE1 -I [-u] -t .. -l .. -m .. -o OUT   (in-place rebuild, :777-852; stage 2 of -2)

STAGE1 is the model file the oracle writes for (stage1 corpus, stage1 options) -- the same bytes in the tests -- so the
fixture needs no binary blobs.  Runs only in the development container; the JSON is committed.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from make_golden import CORPORA, corpus_body  # noqa: E402

# (corpus, stage-1 corpus, stage-1 oracle options, mode, unindexed, {cli options})
CASES = [
    ("hamlet", "hamlet", dict(mintokens=2, maxlength=3), "j", True, dict(t=2, l=5)),
    ("hamlet", "hamlet", dict(mintokens=2, maxlength=4), "I", False, dict(t=2, l=4)),          # stage 2 of `-2 -t 2 -l 4`
    ("hamlet", "hamlet", dict(mintokens=1, maxlength=3), "j", True, dict(t=3, l=3)),
    ("hamlet", "hamlet", dict(mintokens=1, maxlength=3), "I", True, dict(t=3, l=2)),           # -l widened to the model's 3
    ("hamlet", "hamlet", dict(mintokens=2, maxlength=5), "j", False, dict(t=2, l=5, m=2)),
    ("hamlet", "hamlet", dict(mintokens=1, maxlength=4), "I", True, dict(t=1, l=4, m=2)),      # types() computed lazily at write
    ("hamlet", "hamlet", dict(mintokens=1, maxlength=4, minlength=2), "I", False, dict(t=1, l=4)),
    ("hamlet", "republic", dict(mintokens=2, maxlength=4), "j", True, dict(t=1, l=4)),
    ("noeos", "noeos", dict(mintokens=1, maxlength=3), "j", True, dict(t=1, l=3)),             # streamed source repeats the last byte
    ("noeos", "noeos", dict(mintokens=1, maxlength=3), "I", True, dict(t=1, l=3)),             # preloaded source does not
    ("noeos_multibyte", "noeos_multibyte", dict(mintokens=1, maxlength=3), "j", False, dict(t=1, l=3)),
    ("empty_sentences", "empty_sentences", dict(mintokens=2, maxlength=3), "I", False, dict(t=2, l=3)),
    ("only_delims", "hamlet", dict(mintokens=2, maxlength=3), "j", True, dict(t=2, l=3)),
    ("threebyte", "threebyte", dict(mintokens=2, maxlength=5), "I", False, dict(t=2, l=5)),
    ("republic", "republic", dict(mintokens=2, maxlength=5), "I", False, dict(t=2, l=5)),      # stage 2 of `-2 -t 2 -l 5`
    ("republic", "republic", dict(mintokens=2, maxlength=5), "j", True, dict(t=5, l=4)),
    ("republic", "hamlet", dict(mintokens=1, maxlength=5), "j", True, dict(t=2, l=5)),
    ("zipf300k_phr", "zipf300k_phr", dict(mintokens=2, maxlength=5), "I", True, dict(t=3, l=5)),
    ("zipf300k_phr", "zipf200k", dict(mintokens=2, maxlength=3), "j", False, dict(t=2, l=3)),
    ("zipf2m", "zipf2m", dict(mintokens=2, maxlength=5), "I", True, dict(t=2, l=5)),
]


def oracle_side(body, stage1_blob, mode, unindexed, cli):
    """The oracle's account of one case: how the CLI loads the model and which options train() finally sees."""
    t, l, m = cli.get("t", -1), cli.get("l", 100), cli.get("m", 1)
    if mode == "I":
        # src/patternmodeller.cpp:804-821 / :828-845: load with the options as filters + DORESET, widen MAXLENGTH/MINLENGTH to the model's;
        # the corpus was preloaded (LOADCORPUS stays true with -I, :728-737)
        cm = oracle.load_model(stage1_blob, mintokens=t, minlength=m, maxlength=l, doreset=1, indexed=0 if unindexed else 1)
        f = cm.flat()
        return cm, dict(mintokens=t, maxlength=max(l, f.maxn), minlength=min(m, f.minn), indexed=0 if unindexed else 1, streamed=0), True
    # :717-721 PatternSetModel(inputmodelfile2, constrainoptions); unindexed output streams the corpus file (:728-737)
    cm = oracle.load_model(stage1_blob, mintokens=t, minlength=m, maxlength=l, indexed=0)
    return cm, dict(mintokens=t, maxlength=l, minlength=m, indexed=0 if unindexed else 1, streamed=1 if unindexed else 0), False


def main():
    assert oracle.build_ref(), "reference not built (need /root/reference)"
    out = {"reference": "proycon/colibri-core v2.5.9 (4c07c5a), built by oracle/Makefile", "cases": []}
    with tempfile.TemporaryDirectory() as td:
        for cname, s1name, s1opts, mode, unindexed, cli in CASES:
            body = corpus_body(CORPORA[cname])
            cpath = os.path.join(td, cname + ".colibri.dat")
            with open(cpath, "wb") as f:
                f.write(b"\xa2\x02" + body)
            s1 = oracle.train_to_modelfile(corpus_body(CORPORA[s1name]), **s1opts)
            s1path = os.path.join(td, "stage1.patternmodel")
            with open(s1path, "wb") as f:
                f.write(s1)
            mpath = os.path.join(td, "m.patternmodel")
            if os.path.exists(mpath):
                os.remove(mpath)
            args = ["-f", cpath] + (["-i", s1path, "-I"] if mode == "I" else ["-j", s1path]) + ([] if not unindexed else ["-u"]) + ["-o", mpath]
            for k, v in cli.items():
                args += ["-" + k, v]
            rc, err = oracle.ref_cli(args)
            assert rc == 0, err
            ref = oracle.parse_modelfile(open(mpath, "rb").read())
            case = {
                "corpus": cname, "stage1_corpus": s1name, "stage1_options": s1opts, "mode": mode, "unindexed": unindexed, "cli": cli,
                "tokens": ref.tokens, "types": ref.types, "patterns": len(ref), "passes": [list(p) for p in oracle.parse_ref_passes(err)],
                "occurrences": int(ref.counts.sum()), "digest": ref.digest(),
            }
            if len(ref) <= 100:
                c = ref.canonical()
                case["model"] = [[c.key(i).hex(), int(c.counts[i])] + ([c.refs(i)] if c.ref_off is not None else []) for i in range(len(c))]
            out["cases"].append(case)
            print(cname, s1name, mode, "u" if unindexed else "i", cli, len(ref), ref.tokens, ref.types, case["passes"], case["digest"][:12])
    with open(os.path.join(HERE, "golden_constrained.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
