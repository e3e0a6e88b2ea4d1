"""GPU parity of the paths around train() that SURVEY.md 8(f) ranks next: training under a constraint model (8f-2: the CLI's
-j / -I / stage 2 of -2), model loading with the options as filters and device-side queries (8f-3).  Everything goes through
the C ABI (include/colibri_b200.h); the checker is the oracle, itself pinned to the unmodified reference CLI
(tests/golden/golden_constrained.json).  Bit-exact: patterns, counts, occurrence lists, header numbers, pass statistics."""
import random

import numpy as np
import pytest

import oracle
from conftest import cli_load_and_train_options, constrained_case_id, corpus_body, load_constrained_cases
from test_gpu_parity import cb, to_flat

pytestmark = pytest.mark.gpu

CASES = load_constrained_cases()


def gpu_load(blob, constrain=None, mintokens=-1, minlength=1, maxlength=100, dongrams=1, doskipgrams=1, doflexgrams=1, doreset=0, indexed=0):
    o = cb().PatternModelOptions(MINTOKENS=mintokens, MINLENGTH=minlength, MAXLENGTH=maxlength, DOREMOVENGRAMS=not dongrams, DOREMOVESKIPGRAMS=not doskipgrams,
                                 DOREMOVEFLEXGRAMS=not doflexgrams, DORESET=doreset, model_type=20 if indexed else 10, QUIET=1)
    return cb().load_model(blob, o, constrain=constrain)


def gpu_train_constrained(body, cm, inplace, mintokens=-1, maxlength=100, minlength=1, indexed=0, streamed=1):
    o = cb().PatternModelOptions(MINTOKENS=mintokens, MAXLENGTH=maxlength, MINLENGTH=minlength, model_type=20 if indexed else 10, streamed=streamed, QUIET=1)
    return cb().train_constrained(body, cm, inplace=inplace, options=o)


def run_both(body, stage1, load_kw, train_kw, inplace):
    """The same load + constrained train on the device and in the oracle."""
    ocm = oracle.load_model(stage1, **load_kw)
    gcm = gpu_load(stage1, **load_kw)
    of = ocm.flat()
    gf = to_flat(gcm)
    assert gf.same_patterns(of) and (gf.tokens, gf.types, gf.maxn, gf.minn, gf.hasskipgrams) == (of.tokens, of.types, of.maxn, of.minn, of.hasskipgrams)
    train_kw = dict(train_kw)
    if inplace:  # src/patternmodeller.cpp:814-817
        train_kw["maxlength"] = max(train_kw["maxlength"], gcm.maxlength())
        train_kw["minlength"] = min(train_kw["minlength"], gcm.minlength())
    want = oracle.train_constrained(body, ocm, inplace=inplace, **train_kw)
    got = gpu_train_constrained(body, gcm, inplace, **train_kw)
    return got, want


def assert_same_model(got, want):
    flat = to_flat(got)
    assert (got.tokens(), got.types(), len(got)) == (want.tokens, want.types, len(want))
    assert (got.maxlength(), got.minlength()) == (want.maxn, want.minn)
    assert got.passes() == want.passes
    assert flat.same_patterns(want)


@pytest.mark.parametrize("case", CASES, ids=[constrained_case_id(c) for c in CASES])
def test_gpu_constrained_matches_reference_golden(golden, case):
    body = corpus_body(golden, case["corpus"])
    stage1 = oracle.train_to_modelfile(corpus_body(golden, case["stage1_corpus"]), **case["stage1_options"])
    load_kw, train_kw, inplace = cli_load_and_train_options(case)
    got, want = run_both(body, stage1, load_kw, train_kw, inplace)
    assert (got.tokens(), got.types(), len(got)) == (case["tokens"], case["types"], case["patterns"])
    assert [(p[1], p[3]) for p in got.passes()] == [(p[0], p[2]) for p in case["passes"]]
    flat = to_flat(got)
    assert int(flat.counts.sum()) == case["occurrences"]
    assert flat.digest() == case["digest"]
    assert_same_model(got, want)
    # the model file the device path writes parses back to the reference's digest
    assert oracle.parse_modelfile(got.to_bytes()).digest() == case["digest"]


def _rand_corpus(rng, nsent, vocab, maxlen):
    sents = []
    for _ in range(nsent):
        n = rng.randint(0, maxlen)
        sents.append([6 + min(int(rng.paretovariate(1.1)) - 1, vocab - 1) if rng.random() < 0.9 else 6 + rng.randrange(vocab) for _ in range(n)])
    return oracle.encode_corpus(sents)


@pytest.mark.parametrize("seed", range(40))
def test_gpu_constrained_random_vs_oracle(seed):
    """Random corpora / stage-1 models / options, the recipe that pinned the oracle to the reference CLI on 400 seeds."""
    rng = random.Random(1000 + seed)
    body = _rand_corpus(rng, rng.randint(1, 60), rng.choice([5, 30, 200, 20000, 3000000]), rng.choice([4, 12, 30]))
    if not body.strip(b"\0"):
        body = bytes([6, 7, 0])
    body2 = body if rng.random() < 0.5 else _rand_corpus(rng, rng.randint(1, 60), rng.choice([5, 30, 200]), rng.choice([4, 12, 30]))
    if not body2.strip(b"\0"):
        body2 = body
    if rng.random() < 0.2:
        body = body.rstrip(b"\0") or body  # missing final delimiter: streamed and preloaded sources differ
    t1, l1, m1, idx1 = rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 3, 5, 8]), rng.choice([1, 1, 1, 2]), rng.random() < 0.4
    stage1 = oracle.train_to_modelfile(body2, mintokens=t1, maxlength=l1, minlength=m1, indexed=int(idx1), streamed=0 if idx1 else 1)
    t, l, m = rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 3, 5, 8, 100]), rng.choice([1, 1, 1, 2, 3])
    indexed, inplace = rng.random() < 0.5, rng.random() < 0.5
    if inplace:
        load_kw = dict(mintokens=t, minlength=m, maxlength=l, doreset=1, indexed=int(indexed))
        train_kw = dict(mintokens=t, maxlength=l, minlength=m, indexed=int(indexed), streamed=0)
    else:
        load_kw = dict(mintokens=t, minlength=m, maxlength=l, indexed=0)
        train_kw = dict(mintokens=t, maxlength=l, minlength=m, indexed=int(indexed), streamed=0 if indexed else 1)
    got, want = run_both(body, stage1, load_kw, train_kw, inplace)
    assert_same_model(got, want)


def test_gpu_load_filters_match_oracle(golden):
    """PatternMapStore::read filters on the device: threshold, length window, category switches, DORESET, constraint, type conversions."""
    body = corpus_body(golden, "hamlet")
    blob = oracle.train_to_modelfile(body, mintokens=2, maxlength=5, doskipgrams_exhaustive=1, streamed=0)
    iblob = oracle.train_to_modelfile(body, mintokens=2, maxlength=4, indexed=1, streamed=0)
    for kw in (dict(), dict(mintokens=3, minlength=2, maxlength=4, doskipgrams=0), dict(dongrams=0), dict(doreset=1), dict(mintokens=100), dict(minlength=9)):
        g, o = to_flat(gpu_load(blob, **kw)), oracle.load_model(blob, **kw).flat()
        assert g.same_patterns(o), kw
        assert (g.tokens, g.types, g.maxn, g.minn, g.hasskipgrams) == (o.tokens, o.types, o.maxn, o.minn, o.hasskipgrams), kw
    for kw in (dict(indexed=0), dict(indexed=1), dict(indexed=1, mintokens=3, maxlength=2), dict(indexed=1, doreset=1)):
        g, o = to_flat(gpu_load(iblob, **kw)), oracle.load_model(iblob, **kw).flat()
        assert g.same_patterns(o), kw
    g, o = to_flat(gpu_load(blob, indexed=1)), oracle.load_model(blob, indexed=1).flat()  # unindexed read as indexed: patterns without counts
    assert g.same_patterns(o) and int(g.counts.sum()) == 0
    small = oracle.train_to_modelfile(body, mintokens=2, maxlength=2)
    g = to_flat(gpu_load(blob, constrain=gpu_load(small)))
    o = oracle.load_model(blob, constrain=oracle.load_model(small)).flat()
    assert g.same_patterns(o) and len(g) == len(oracle.parse_modelfile(small))
    # malformed files are refused, not guessed at
    for bad in (b"", b"\x01" * 40, blob[:40], bytes([0, 30, 2]) + blob[3:]):
        with pytest.raises(cb().ColibriError):
            cb().load_model(bad)


def test_gpu_lookup_batch_matches_model(golden):
    """occurrencecount()/has() on the device (SpookyV2 of the pattern bytes into the HBM index) against the exported model."""
    corpus = cb().Corpus.synthetic(300000, vocab=5000, seed=4)
    m = cb().train(corpus, MINTOKENS=2, MAXLENGTH=4, QUIET=1)
    flat = to_flat(m)
    d = flat.as_dict()
    rng = random.Random(5)
    present = rng.sample(sorted(d), 2000)
    absent = []
    while len(absent) < 2000:
        k = b"".join(oracle.inttobytes(6 + rng.randrange(6000)) for _ in range(rng.randint(1, 5)))
        if k not in d:
            absent.append(k)
    keys = present + absent + [b"\x03", bytes(range(6, 100)) * 3]  # a lone skip token, a key beyond the Short-hash range
    counts, index = m.lookup_batch(keys)
    assert counts[:2000].tolist() == [d[k] for k in present]
    assert not counts[2000:].any() and (index[2000:] == -1).all()
    assert all(flat.key(int(i)) == k for i, k in zip(index[:2000], present))
    assert m.occurrencecount(present[0]) == d[present[0]] and m.has(present[1]) and not m.has(absent[0])
    # an uploaded set answers the same queries
    up = cb().Model.from_flat(flat.keys, flat.key_off, flat.counts, tokens=flat.tokens, types=flat.types)
    c2, _ = up.lookup_batch(keys)
    assert np.array_equal(c2, counts)
    assert to_flat(up).same_patterns(flat) and (up.maxlength(), up.minlength()) == (m.maxlength(), m.minlength())


def test_gpu_two_stage_equals_direct_indexed_training():
    """The point of the CLI's -2: stage 1 (unindexed) + constrained in-place rebuild (indexed) yields the patterns and occurrence
    lists of a direct indexed training run (only totaltypes differs: the rebuild reports size(), patternmodel.h:1199-1201)."""
    corpus = cb().Corpus.synthetic(2000000, vocab=20000, seed=11, phrase_permille=200, nphrases=3000)
    stage1 = cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
    loaded = cb().load_model(stage1.to_bytes(), MINTOKENS=2, MAXLENGTH=5, DORESET=1, model_type=20, QUIET=1)
    assert len(loaded) == len(stage1)
    stage2 = cb().train_constrained(corpus, loaded, inplace=True, MINTOKENS=2, MAXLENGTH=5, model_type=20, streamed=0, QUIET=1)
    direct = cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, model_type=20, streamed=0, QUIET=1)
    assert to_flat(stage2).same_patterns(to_flat(direct))
    assert stage2.tokens() == direct.tokens() and stage2.types() == len(stage1)
    assert stage2.counters()["kernel_launches"] > 0


def test_gpu_constrained_refusals():
    m = cb().Model.from_flat(np.array([6, 7], dtype=np.uint8), np.array([0, 1, 2], dtype=np.uint64), np.array([2, 2], dtype=np.uint32))
    body = bytes([6, 7, 0])
    for kw in (dict(DOSKIPGRAMS_EXHAUSTIVE=1), dict(MINTOKENS_UNIGRAMS=5), dict(MAXLENGTH=300)):
        with pytest.raises(cb().ColibriError) as ei:
            cb().train_constrained(body, m, **kw)
        assert ei.value.code == 2
    with pytest.raises(cb().ColibriError):  # duplicate patterns are not a set
        cb().Model.from_flat(np.array([6, 6], dtype=np.uint8), np.array([0, 1, 2], dtype=np.uint64)).lookup_batch([b"\x06"])


# ---- flexgrams abstracted from skipgrams (first piece)
import json  # noqa: E402
import os  # noqa: E402

from conftest import GOLDEN_DIR  # noqa: E402

FLEX_CASES = json.load(open(os.path.join(GOLDEN_DIR, "golden_flex.json")))["cases"]


def gpu_train_indexed_skipgrams(body, o):
    opts = cb().PatternModelOptions(MINTOKENS=o.get("mintokens", -1), MAXLENGTH=o.get("maxlength", 100), MINSKIPTYPES=o.get("minskiptypes", 2),
                                    MINTOKENS_SKIPGRAMS=o.get("mintokens_skipgrams", -1), DOSKIPGRAMS=1, model_type=20, streamed=0, QUIET=1)
    return cb().train(body, opts)


@pytest.mark.parametrize("case", FLEX_CASES, ids=["%s-%s" % (c["corpus"], "".join("%s%s" % kv for kv in sorted(c["cli"].items())) or "default") for c in FLEX_CASES])
def test_gpu_flexgrams_fromskipgrams_match_reference_golden(golden, case):
    """computeflexgrams_fromskipgrams on the device against the reference CLI's `-s -F S` files (reference KAT: 22 found, 155 patterns)."""
    body = corpus_body(golden, case["corpus"])
    m = gpu_train_indexed_skipgrams(body, case["options"])
    found, fm = m.flexgrams_fromskipgrams()
    assert (found, len(fm), fm.tokens(), fm.types()) == (case["flexfound"], case["patterns"], case["tokens"], case["types"])
    assert fm.hasflexgrams == (found > 0) and fm.hasskipgrams == m.hasskipgrams
    flat = to_flat(fm)
    assert int(flat.counts.sum()) == case["occurrences"]
    assert flat.sorted_refs().digest() == case["digest_sorted_refs"]
    want = oracle.train(body, flexfromskip=1, **case["options"])
    assert flat.same_patterns(want)  # both emit every occurrence list ascending
    assert oracle.parse_modelfile(fm.to_bytes()).sorted_refs().digest() == case["digest_sorted_refs"]
    assert len(m) == case["patterns"] - case["flexfound"]  # the source model is untouched


@pytest.mark.parametrize("seed", range(25))
def test_gpu_flexgrams_random_vs_oracle(seed):
    """Random corpora: the clean iteration (every skipgram once), which is what the oracle restates; the reference itself is only
    comparable where its insert-while-iterating does not rehash (tests/golden/make_golden_flex.py)."""
    rng = random.Random(7000 + seed)
    body = _rand_corpus(rng, rng.randint(1, 60), rng.choice([5, 30, 200]), rng.choice([6, 12, 30]))
    if not body.strip(b"\0"):
        body = bytes([6, 7, 8, 0, 6, 9, 8, 0])
    o = dict(mintokens=rng.choice([2, 2, 3]), maxlength=rng.choice([3, 4, 5, 6]), minskiptypes=rng.choice([1, 2]), indexed=1, doskipgrams=1, streamed=0)
    want = oracle.train(body, flexfromskip=1, **o)
    found, fm = gpu_train_indexed_skipgrams(body, o).flexgrams_fromskipgrams()
    assert found == want.flexfound and len(fm) == len(want)
    assert to_flat(fm).same_patterns(want)


def test_gpu_flexgrams_refusals_and_cli(tmp_path):
    body = open(os.path.join(GOLDEN_DIR, "hamlet.colibri.dat"), "rb").read()[2:]
    with pytest.raises(cb().ColibriError) as ei:
        cb().train(body, MINTOKENS=2, MAXLENGTH=3, QUIET=1).flexgrams_fromskipgrams()  # unindexed
    assert ei.value.code == 1
    found, fm = gpu_train_indexed_skipgrams(body, {}).flexgrams_fromskipgrams()
    with pytest.raises(cb().ColibriError) as ei:
        fm.flexgrams_fromskipgrams()  # already holds flexgrams
    assert ei.value.code == 2
    import subprocess

    from conftest import ROOT

    cli = os.path.join(ROOT, "colibri-core_b200", "bin", "colibri-patternmodeller")
    subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True)
    out = str(tmp_path / "flex.colibri.patternmodel")
    r = subprocess.run([cli, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat"), "-F", "S", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "22 flexgrams found" in r.stderr
    got = oracle.parse_modelfile(open(out, "rb").read())
    assert len(got) == 155 and got.sorted_refs().digest() == FLEX_CASES[0]["digest_sorted_refs"]
