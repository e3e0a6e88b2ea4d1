"""The reference binding's names on the B200 library (colibricore_b200 = colibri-core_b200/pybinding.py): the model part of the reference's test.py
(/root/reference/test.py:230-311, the same expected numbers), the reverse index and the co-occurrence relations against the committed answers of
the unmodified reference (tests/golden/golden_relations.json) and against the oracle."""
import json
import os
import sys

import pytest

import oracle
from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

sys.path.insert(0, GOLDEN_DIR)
from make_golden_relations import corpus as relations_corpus  # noqa: E402

HAMLET = os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")
CLS = os.path.join(GOLDEN_DIR, "hamlet.colibri.cls")
with open(os.path.join(GOLDEN_DIR, "golden_relations.json")) as f:
    REL_CASES = json.load(f)["cases"]


def cc():
    import colibricore_b200

    return colibricore_b200


def test_reference_python_test_model_section(tmp_path):
    colibricore = cc()
    encoder, decoder = colibricore.ClassEncoder(CLS), colibricore.ClassDecoder(CLS)
    options = colibricore.PatternModelOptions(doskipgrams_exhaustive=True)
    unindexedmodel = colibricore.UnindexedPatternModel()
    unindexedmodel.train(HAMLET, options)
    assert (len(unindexedmodel), unindexedmodel.types(), unindexedmodel.tokens()) == (385, 186, 354)  # test.py:237-239
    unindexedmodel.printmodel(decoder)
    unindexedmodel.report()
    unindexedmodel.histogram()
    out = str(tmp_path / "data.colibri.patternmodel")
    unindexedmodel.write(out)
    unindexedmodel = colibricore.UnindexedPatternModel(out)
    assert sum(1 for _ in unindexedmodel) == 385
    assert sum(1 for _pattern, _value in unindexedmodel.items()) == 385
    assert unindexedmodel[encoder.buildpattern("to be")] == 2  # test.py:272
    assert encoder.buildpattern("to be") in unindexedmodel and encoder.buildpattern("be to") not in unindexedmodel

    corpus = colibricore.IndexedCorpus(HAMLET)
    assert corpus.sentencecount() == 40  # test.py:278
    assert sum(1 for _ in corpus.sentences()) == 40
    assert len(corpus) == 354
    options = colibricore.PatternModelOptions(doskipgrams=True)
    indexedmodel = colibricore.IndexedPatternModel(reverseindex=corpus)
    indexedmodel.train(HAMLET, options)
    assert (len(indexedmodel), indexedmodel.types(), indexedmodel.tokens()) == (133, 186, 354)  # test.py:290-292
    out = str(tmp_path / "data.colibri.indexedpatternmodel")
    indexedmodel.write(out)
    indexedmodel = colibricore.IndexedPatternModel(out)
    assert sum(1 for _p, v in indexedmodel.items() if len(v) >= 1) == 133
    assert len(indexedmodel[encoder.buildpattern("to be")]) == 2  # test.py:311
    assert decoder.decode(bytes(encoder.buildpattern("to be"))) == "to be"


@pytest.mark.parametrize("case", REL_CASES, ids=["%s-t%d-l%d" % (c["corpus"], c["t"], c["l"]) for c in REL_CASES])
def test_reverse_index_and_cooccurrence_equal_reference(case, tmp_path):
    colibricore = cc()
    body = relations_corpus(case["corpus"])
    path = str(tmp_path / "c.colibri.dat")
    with open(path, "wb") as f:
        f.write(b"\xa2\x02" + body)
    corpus = colibricore.IndexedCorpus(path)
    model = colibricore.IndexedPatternModel(reverseindex=corpus)
    model.train(path, colibricore.PatternModelOptions(mintokens=case["t"], maxlength=case["l"]))
    assert len(model) == int(case["header"]["patterns"]) and corpus.sentencecount() == int(case["header"]["sentences"])
    assert model.totaloccurrencesingroup(0, 0) == int(case["header"]["total"])
    # getreverseindex of every position of the corpus, in one batch and one by one for a few
    refs = [(s, t) for s, t, _ in case["G"]]
    got = model.getreverseindex_batch(refs)
    assert [sorted(bytes(p).hex() for p in g) for g in got] == [sorted(v) for _, _, v in case["G"]]
    for s, t, v in case["G"][:25]:
        assert sorted(bytes(p).hex() for p in model.getreverseindex((s, t))) == sorted(v)
    assert list(model.getreverseindex((corpus.sentencecount() + 5, 0))) == [] and list(model.getreverseindex((1, 60000))) == []
    # getrightcooc / getleftcooc of every pattern
    for tag, fn in (("R", model.getrightcooc), ("L", model.getleftcooc)):
        got = sorted([bytes(p).hex(), bytes(q).hex(), j] for p in model for q, j in fn(p))
        assert got == sorted(case[tag]), tag
    # getcooc (both directions, no overlap) of every pattern: defaults, then occurrencethreshold 2 + ordersignificant; one size filter against the oracle
    got = sorted([bytes(p).hex(), bytes(q).hex(), j] for p in model for q, j in model.getcooc(p))
    assert got == sorted(case["C"])
    got = sorted([bytes(p).hex(), bytes(q).hex(), j] for p in model for q, j in model.getcooc(p, occurrencethreshold=2, ordersignificant=True))
    assert got == sorted(case["O"])
    pats = {bytes(p): len(v) for p, v in model.items()}
    want = oracle.cooc_both(body, pats, size=2)
    got = {(bytes(p), bytes(q)): j for p in model for q, j in model.getcooc(p, size=2)}
    assert got == want
    # npmi: the same doubles
    got_n = sorted([bytes(p).hex(), bytes(q).hex(), v] for p, rel in model.computenpmi(case["threshold"], right=True, left=False).items() for q, v in rel.items())
    want_n = sorted(case["N"])
    assert [(a, b) for a, b, _ in got_n] == [(a, b) for a, b, _ in want_n]
    assert all(x[2] == y[2] for x, y in zip(got_n, want_n))
    # flexgrams from co-occurrence: the oracle's clean iteration (pinned to the reference where the reference itself is reproducible)
    patterns = {bytes(p): len(v) for p, v in model.items()}
    want_found, want_flex = oracle.flexgrams_fromcooc(body, patterns, case["threshold"])
    found = model.computeflexgrams_fromcooc(case["threshold"])
    assert found == want_found
    flex = {bytes(p): len(v) for p, v in model.items() if p.category() == colibricore.FLEXGRAM}
    assert flex == want_flex
    if case["flex_check"] == "subset":
        ref = {k: v for k, v in case["X"]}
        assert all(ref.get(k.hex()) == v for k, v in flex.items())
    assert len(model) == len(patterns) + found


def test_cooccurrence_at_scale_matches_oracle_properties():
    """2 M tokens: the relation sums obey what the definition implies -- joint(P, P) = sum over the occurrences of P of the positions right of it,
    and every relation joins a pattern with one of its own prefixes or extensions."""
    import colibri_core_b200 as cb

    corpus = cb.Corpus.synthetic(2000000, vocab=50000, seed=3)
    m = cb.train(corpus, MINTOKENS=2, MAXLENGTH=4, model_type=20, streamed=0, QUIET=1)
    ri = cb.ReverseIndex(m, corpus, streamed=0)
    p, q, j = ri.cooc(left=False)
    keys, off, counts, refs = m.export()
    kb = keys.tobytes()
    key = lambda i: kb[int(off[i]):int(off[i + 1])]  # noqa: E731
    assert len(p) > 0
    for a, b in list(zip(p.tolist(), q.tolist()))[:20000]:
        ka, kq = key(a), key(b)
        assert ka.startswith(kq) or kq.startswith(ka)
    starts = ri.sentence_starts()
    import numpy as np

    rs, rt, ro = refs
    self_rel = {int(a): int(c) for a, b, c in zip(p.tolist(), q.tolist(), j.tolist()) if a == b}
    for i in list(self_rel)[:300]:
        n = len([x for x in key(i) if x < 128])
        sl = (starts[rs[int(ro[i]):int(ro[i + 1])]] - 1 - starts[rs[int(ro[i]):int(ro[i + 1])] - 1]).astype(np.int64)
        w = np.maximum(0, sl - 1 - (rt[int(ro[i]):int(ro[i + 1])].astype(np.int64) + n))
        assert int(w.sum()) == self_rel[i]
    ri.close()


with open(os.path.join(GOLDEN_DIR, "golden_stats.json")) as f:
    STATS_CASES = json.load(f)["cases"]


@pytest.mark.parametrize("case", STATS_CASES, ids=["%s-t%d-l%d-s%d" % (c["corpus"], c["t"], c["l"], c["skipgrams"]) for c in STATS_CASES])
def test_group_statistics_equal_reference(case, tmp_path):
    """totaloccurrencesingroup / totalpatternsingroup / totalwordtypesingroup of every (category, n) of a model trained on the device equal the
    unmodified reference's (tests/golden/make_golden_stats.py)."""
    colibricore = cc()
    path = str(tmp_path / "c.colibri.dat")
    with open(path, "wb") as f:
        f.write(b"\xa2\x02" + relations_corpus(case["corpus"]))
    corpus = colibricore.IndexedCorpus(path)
    model = colibricore.IndexedPatternModel(reverseindex=corpus)
    model.train(path, colibricore.PatternModelOptions(mintokens=case["t"], maxlength=case["l"], doskipgrams=bool(case["skipgrams"])))
    for c, n, occ, npat, wtypes in case["S"]:
        assert (model.totaloccurrencesingroup(c, n), model.totalpatternsingroup(c, n), model.totalwordtypesingroup(c, n)) == (occ, npat, wtypes), (c, n)


@pytest.mark.parametrize("seed", range(12))
def test_relations_equal_oracle_on_random_input(seed, tmp_path):
    """Random corpora (the ones tests/test_oracle_vs_ref_random.py pins the oracle to the reference with): reverse index, right / left
    co-occurrence, getcooc with and without its filters, group statistics -- device against oracle."""
    import random

    colibricore = cc()
    rng = random.Random(5000 + seed)
    body = b""
    while not body:
        vocab = rng.choice([3, 6, 15, 60])
        sentences = [[6 + min(int(rng.paretovariate(1.1)) - 1, vocab - 1) for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 13, 21]))] for _ in range(rng.randint(1, 40))]
        body = bytes(oracle.encode_corpus(sentences))
    t, l = rng.choice([2, 2, 3]), rng.choice([2, 3, 4, 5])
    patterns = oracle.train(body, mintokens=t, maxlength=l, indexed=1, streamed=0).as_dict()
    if not patterns:
        pytest.skip("no pattern survives")
    path = str(tmp_path / "c.colibri.dat")
    with open(path, "wb") as f:
        f.write(b"\xa2\x02" + body)
    corpus = colibricore.IndexedCorpus(path)
    model = colibricore.IndexedPatternModel(reverseindex=corpus)
    model.train(path, colibricore.PatternModelOptions(mintokens=t, maxlength=l))
    assert {bytes(p): len(v) for p, v in model.items()} == patterns
    rindex = oracle.reverse_index(body, patterns)
    refs = sorted(rindex)
    got = model.getreverseindex_batch(refs)
    assert [sorted(bytes(p) for p in g) for g in got] == [sorted(rindex[r]) for r in refs]
    for left, fn in ((False, model.getrightcooc), (True, model.getleftcooc)):
        assert {(bytes(p), bytes(q)): j for p in model for q, j in fn(p)} == oracle.cooc(body, patterns, left=left)
    assert {(bytes(p), bytes(q)): j for p in model for q, j in model.getcooc(p)} == oracle.cooc_both(body, patterns)
    assert {(bytes(p), bytes(q)): j for p in model for q, j in model.getcooc(p, occurrencethreshold=2, size=1, ordersignificant=True)} == \
        oracle.cooc_both(body, patterns, occurrencethreshold=2, size=1, ordersignificant=True)
    for (c, n), (occ, npat, wtypes) in oracle.group_stats(patterns).items():
        assert (model.totaloccurrencesingroup(c, n), model.totalpatternsingroup(c, n), model.totalwordtypesingroup(c, n)) == (occ, npat, wtypes), (c, n)
