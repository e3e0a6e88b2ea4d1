"""The oracle's restatement of training under a constraint model (PatternModel::train with constrainbymodel != NULL,
include/patternmodel.h:880-1345; the CLI's -j / -I / stage 2 of -2) and of model loading with the options as filters
(PatternModel::load :781-861, PatternMapStore::read patternstore.h:555-619) against golden vectors written by the
UNMODIFIED reference CLI (tests/golden/golden_constrained.json).  CPU only."""
import pytest

import oracle
from conftest import cli_load_and_train_options, constrained_case_id, corpus_body, load_constrained_cases

CASES = load_constrained_cases()


def oracle_run(golden, case):
    body = corpus_body(golden, case["corpus"])
    stage1 = oracle.train_to_modelfile(corpus_body(golden, case["stage1_corpus"]), **case["stage1_options"])
    load_kw, train_kw, inplace = cli_load_and_train_options(case)
    cm = oracle.load_model(stage1, **load_kw)
    if inplace:  # src/patternmodeller.cpp:814-817: MAXLENGTH / MINLENGTH widened to the loaded model's
        f = cm.flat()
        train_kw["maxlength"] = max(train_kw["maxlength"], f.maxn)
        train_kw["minlength"] = min(train_kw["minlength"], f.minn)
    return oracle.train_constrained(body, cm, inplace=inplace, **train_kw)


@pytest.mark.parametrize("case", CASES, ids=[constrained_case_id(c) for c in CASES])
def test_oracle_constrained_matches_reference_cli(golden, case):
    m = oracle_run(golden, case)
    assert (m.tokens, m.types, len(m)) == (case["tokens"], case["types"], case["patterns"])
    assert int(m.counts.sum()) == case["occurrences"]
    assert [(p[1], p[3]) for p in m.passes] == [(p[0], p[2]) for p in case["passes"]]
    assert m.digest() == case["digest"]
    if "model" in case:
        c = m.canonical()
        got = [[c.key(i).hex(), int(c.counts[i])] + ([[list(r) for r in c.refs(i)]] if c.ref_off is not None else []) for i in range(len(c))]
        assert got == case["model"]


def test_load_filters(golden):
    """PatternMapStore::read: count >= MINTOKENS, MINLENGTH <= n <= MAXLENGTH, category switches, DORESET, constraint."""
    body = corpus_body(golden, "hamlet")
    blob = oracle.train_to_modelfile(body, mintokens=2, maxlength=5, doskipgrams_exhaustive=1, streamed=0)
    full = oracle.parse_modelfile(blob)
    allm = oracle.load_model(blob).flat()
    assert allm.same_patterns(full) and (allm.tokens, allm.types) == (full.tokens, full.types)
    assert allm.hasskipgrams and (allm.maxn, allm.minn) == (5, 1)

    def shape(k):
        toks, cat, start = 0, 0, True
        for b in k:
            if b < 128:
                if start and cat == 0 and b in (3, 4):
                    cat = b - 2
                toks += 1
                start = True
            else:
                start = False
        return toks, cat

    d = full.as_dict()
    f = oracle.load_model(blob, mintokens=3, minlength=2, maxlength=4, doskipgrams=0).flat()
    want = {k: v for k, v in d.items() if v >= 3 and 2 <= shape(k)[0] <= 4 and shape(k)[1] == 0}
    assert f.as_dict() == want and not f.hasskipgrams and (f.maxn, f.minn) == (max(shape(k)[0] for k in want), min(shape(k)[0] for k in want))
    g = oracle.load_model(blob, dongrams=0).flat()
    assert g.as_dict() == {k: v for k, v in d.items() if shape(k)[1] == 1}
    r = oracle.load_model(blob, doreset=1).flat()
    assert set(r.as_dict()) == set(d) and int(r.counts.sum()) == 0
    small = oracle.load_model(oracle.train_to_modelfile(body, mintokens=2, maxlength=2))
    c = oracle.load_model(blob, constrain=small).flat()
    assert c.as_dict() == {k: v for k, v in d.items() if k in small.flat().as_dict()}
    # indexed file read as unindexed keeps the counts; unindexed read as indexed loses them (patternmodel.h:827-837)
    iblob = oracle.train_to_modelfile(body, mintokens=2, maxlength=3, indexed=1, streamed=0)
    assert oracle.load_model(iblob, indexed=0).flat().same_patterns(oracle.train(body, mintokens=2, maxlength=3))
    assert oracle.load_model(iblob, indexed=1).flat().same_patterns(oracle.train(body, mintokens=2, maxlength=3, indexed=1, streamed=0))
    lost = oracle.load_model(oracle.train_to_modelfile(body, mintokens=2, maxlength=3), indexed=1).flat()
    assert len(lost) == 81 and int(lost.counts.sum()) == 0


# ---- flexgrams abstracted from skipgrams (first piece): tests/golden/golden_flex.json, written by make_golden_flex.py
import json  # noqa: E402
import os  # noqa: E402

from conftest import GOLDEN_DIR  # noqa: E402

FLEX_CASES = json.load(open(os.path.join(GOLDEN_DIR, "golden_flex.json")))["cases"]


@pytest.mark.parametrize("case", FLEX_CASES, ids=["%s-%s" % (c["corpus"], "".join("%s%s" % kv for kv in sorted(c["cli"].items())) or "default") for c in FLEX_CASES])
def test_oracle_flexgrams_fromskipgrams_match_reference_cli(golden, case):
    """computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744); reference KAT src/test.cpp:1440-1443: 22 found, 155 patterns."""
    m = oracle.train(corpus_body(golden, case["corpus"]), flexfromskip=1, **case["options"])
    assert (m.flexfound, len(m), m.tokens, m.types, int(m.counts.sum())) == (case["flexfound"], case["patterns"], case["tokens"], case["types"], case["occurrences"])
    assert m.sorted_refs().digest() == case["digest_sorted_refs"]
