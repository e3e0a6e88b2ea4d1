"""Parity of the CUDA path (through the C ABI, include/colibri_b200.h) against the oracle and the committed golden
vectors of the unmodified reference.  Bit-exact: patterns, counts, header numbers and per-pass statistics."""
import os
import tempfile

import numpy as np
import pytest

import oracle
from conftest import FORCED_PATHS, case_id, corpus_body, force_path, load_cases

pytestmark = pytest.mark.gpu

CASES = load_cases()


def cb():
    import colibri_core_b200

    return colibri_core_b200


def to_flat(model):
    """Wrap the product's flat export into the oracle's comparison form (checker side only)."""
    keys, key_off, counts, refs = model.export()
    rs = rt = ro = None
    if refs is not None:
        rs, rt, ro = refs
    return oracle.FlatModel(keys, key_off, counts, model.tokens(), model.types(), model.maxlength(), model.minlength(), model.hasskipgrams, model.model_type, rs, rt, ro,
                            model.passes())


def gpu_options(case):
    o = case["options"]
    return cb().PatternModelOptions(
        MINTOKENS=o.get("mintokens", -1), MINTOKENS_SKIPGRAMS=o.get("mintokens_skipgrams", -1), MINTOKENS_UNIGRAMS=o.get("mintokens_unigrams", 1), MINLENGTH=o.get("minlength", 1),
        MAXLENGTH=o.get("maxlength", 100), MAXBACKOFFLENGTH=o.get("maxbackofflength", 100), MINSKIPTYPES=o.get("minskiptypes", 2), DOSKIPGRAMS_EXHAUSTIVE=o["doskipgrams_exhaustive"], DOSKIPGRAMS=o.get("doskipgrams", 0),
        model_type=20 if o["indexed"] else 10, streamed=o["streamed"], QUIET=1)


def supported(case):
    o = case["options"]
    if o.get("mintokens", 2) > 1 and o.get("maxbackofflength", 100) < o.get("maxlength", 100) - 1:
        return False
    return True


@pytest.mark.parametrize("path", list(FORCED_PATHS))
@pytest.mark.parametrize("case", CASES, ids=[case_id(c) for c in CASES])
def test_gpu_matches_reference_golden(golden, case, path, monkeypatch):
    """Every golden case of the unmodified reference, through every code path of the level loop (conftest.FORCED_PATHS: the occurrence
    filter with colliding buckets, the hot-key cache, the dense pair slots, the list mode -- what bench.py runs at 100 M tokens)."""
    force_path(monkeypatch, path)
    body = corpus_body(golden, case["corpus"])
    if not supported(case):
        with pytest.raises(cb().ColibriError) as ei:
            cb().train(body, gpu_options(case))
        assert ei.value.code == 2  # COLIBRI_E_UNSUPPORTED: loud, never a silent fallback
        return
    if len(body) == 0:
        pytest.skip("empty corpus")
    m = cb().train(body, gpu_options(case))
    assert m.tokens() == case["tokens"]
    assert m.types() == case["types"]
    assert len(m) == case["patterns"]
    assert (m.maxlength(), m.minlength()) == (case["maxn"], case["minn"])
    assert int(m.hasskipgrams) == case["hasskipgrams"]
    assert [(p[1], p[3]) for p in m.passes()] == [(p[0], p[2]) for p in case["passes"]]
    flat = to_flat(m)
    assert int(flat.counts.sum()) == case["occurrences"]
    assert flat.digest() == case["digest"]
    # and against the oracle run on the same bytes
    assert flat.same_patterns(oracle.train(body, **case["options"]))


def test_device_spooky_matches_reference_kats(golden):
    """Pattern::hash on the device (SpookyV2 Short, every length 1..191) against hashes produced by the reference library."""
    msgs = [bytes.fromhex(h) for h, _ in golden["spooky"]]
    want = np.array([v for _, v in golden["spooky"]], dtype=np.uint64)
    got = cb().hash64_batch(msgs)
    assert np.array_equal(got, want)
    assert cb().hash64_batch([b""])[0] == 0


def test_device_tokeniser_matches_codec(golden):
    for name in ("hamlet", "republic", "threebyte", "empty_sentences", "zipf300k_phr"):
        body = corpus_body(golden, name)
        c = cb().Corpus.from_bytes(body)
        got = c.tokens()
        want, i = [], 0
        while i < len(body):
            v, ln = oracle.bytestoint(body[i:i + 8])
            want.append(v)
            i += ln
        assert got.tolist() == want, name


def test_device_synth_corpus_is_bit_identical_to_oracle():
    for kw in (dict(ntokens=100000, vocab=5000, seed=1, mean_sentence=22), dict(ntokens=250000, vocab=300000, seed=9, mean_sentence=7, phrase_permille=250, nphrases=1000)):
        dev = cb().Corpus.synthetic(**kw).download()
        host = oracle.synth_corpus(**kw)
        assert dev.size == host.size and np.array_equal(dev, host)


def test_modelfile_written_by_gpu_is_reference_layout(golden):
    body = corpus_body(golden, "hamlet")
    m = cb().train(body, MINTOKENS=2, MAXLENGTH=3, QUIET=1)
    blob = m.to_bytes()
    assert len(blob) == 563 and blob[:3] == bytes([0, 10, 2])
    assert oracle.parse_modelfile(blob).same_patterns(oracle.train(body, mintokens=2, maxlength=3))
    assert m.occurrencecount(bytes([6])) == oracle.train(body, mintokens=2, maxlength=3).as_dict().get(bytes([6]), 0)
    assert m.occurrencecount(bytes([5, 5, 5])) == 0


@pytest.mark.parametrize("kw", [dict(ntokens=3000000, vocab=100000, seed=11, mean_sentence=22), dict(ntokens=2000000, vocab=50000, seed=12, mean_sentence=18, phrase_permille=200, nphrases=5000)])
@pytest.mark.parametrize("skip", [0, 1])
@pytest.mark.parametrize("path", list(FORCED_PATHS))
def test_gpu_matches_oracle_on_seeded_synthetic(kw, skip, path, monkeypatch):
    force_path(monkeypatch, path)
    corpus = cb().Corpus.synthetic(**kw)
    body = corpus.download()
    m = cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=skip, streamed=0 if skip else 1, QUIET=1)
    want = oracle.train(body, mintokens=2, maxlength=5, doskipgrams_exhaustive=skip, streamed=0 if skip else 1)
    got = to_flat(m)
    assert (got.tokens, got.types, len(got)) == (want.tokens, want.types, len(want))
    assert got.passes == want.passes
    assert got.same_patterns(want)


@pytest.mark.parametrize("kw,maxlength,mintokens", [(dict(ntokens=1500000, vocab=60000, seed=31, mean_sentence=22), 5, 2),
                                                    (dict(ntokens=800000, vocab=2000, seed=32, mean_sentence=9, phrase_permille=300, nphrases=300), 8, 3),
                                                    (dict(ntokens=300000, vocab=500, seed=33, mean_sentence=30), 3, 1)])
@pytest.mark.parametrize("path", ["default", "bench", "part", "part+dense+list"])
def test_indexed_model_matches_oracle(kw, maxlength, mintokens, path, monkeypatch):
    force_path(monkeypatch, path)
    _indexed_model_matches_oracle(kw, maxlength, mintokens)


def _indexed_model_matches_oracle(kw, maxlength, mintokens):
    """Config 4: IndexedPatternModel -- every pattern's sorted (sentence, token) list (datatypes.h:33-89, patternmodel.h:2699-2705, :2789-2800)."""
    corpus = cb().Corpus.synthetic(**kw)
    body = corpus.download()
    m = cb().train(corpus, MINTOKENS=mintokens, MAXLENGTH=maxlength, model_type=20, streamed=0, QUIET=1)
    want = oracle.train(body, mintokens=mintokens, maxlength=maxlength, indexed=1, streamed=0)
    got = to_flat(m)
    assert (got.tokens, got.types, len(got)) == (want.tokens, want.types, len(want))
    assert got.passes == want.passes
    assert got.same_patterns(want)  # keys, counts AND reference lists
    assert oracle.parse_modelfile(m.to_bytes()).same_patterns(want)


def test_indexed_sentence_length_limit():
    """IndexReference.token is a uint16_t (datatypes.h:36).  65 536 tokens in one sentence is the encoder's maximum
    (src/classencoder.cpp:581-588) and must work; longer ones would wrap in the reference and are refused loudly here."""
    body = oracle.encode_corpus([[6 + (i % 7) for i in range(65536)], [6, 7, 8]])
    m = cb().train(body, MINTOKENS=2, MAXLENGTH=2, model_type=20, streamed=0, QUIET=1)
    assert to_flat(m).same_patterns(oracle.train(body, mintokens=2, maxlength=2, indexed=1, streamed=0))
    with pytest.raises(cb().ColibriError) as ei:
        cb().train(oracle.encode_corpus([[6 + (i % 7) for i in range(70000)]]), MINTOKENS=2, MAXLENGTH=2, model_type=20, streamed=0, QUIET=1)
    assert ei.value.code == 2


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
def test_gpu_matches_unmodified_reference_binary():
    kw = dict(ntokens=1500000, vocab=80000, seed=21, mean_sentence=20, phrase_permille=100, nphrases=3000)
    corpus = cb().Corpus.synthetic(**kw)
    body = corpus.download()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.colibri.dat")
        with open(path, "wb") as f:
            f.write(b"\xa2\x02" + body.tobytes())
        st, err = oracle.ref_train(path, os.path.join(td, "m"), unindexed=True, t=2, l=5)
        ref = oracle.parse_modelfile(open(os.path.join(td, "m"), "rb").read())
    m = cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
    got = to_flat(m)
    assert (got.tokens, got.types, len(got)) == (st["tokens"], st["types"], st["patterns"])
    assert [(p[1], p[3]) for p in got.passes] == [(p[0], p[2]) for p in oracle.parse_ref_passes(err)]
    assert got.same_patterns(ref)


def test_large_corpus_invariants():
    """BASELINE-sized properties that need no oracle: idempotence, count conservation, monotone levels."""
    corpus = cb().Corpus.synthetic(ntokens=20000000, vocab=100000, seed=1, mean_sentence=22)
    a = to_flat(cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1))
    b = to_flat(cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, QUIET=1))
    assert a.tokens == 20000000
    assert a.digest() == b.digest()  # same corpus, same options -> same model (table order differs, canonical form does not)
    # an n-gram cannot occur more often than its prefix: occurrences per level are non-increasing
    n_of = np.add.reduceat((a.keys < 128).astype(np.int64), a.key_off[:-1].astype(np.int64)) if len(a) else np.zeros(0)
    occ = [int(a.counts[n_of == n].sum()) for n in range(1, 6)]
    assert all(x >= y for x, y in zip(occ, occ[1:]))
    assert occ[0] <= a.tokens
    assert int(a.counts.min()) >= 2
    # raising the threshold can only remove patterns, and keeps the counts of the survivors
    c = to_flat(cb().train(corpus, MINTOKENS=3, MAXLENGTH=3, QUIET=1)).canonical()
    d = to_flat(cb().train(corpus, MINTOKENS=2, MAXLENGTH=3, QUIET=1)).canonical()
    dd = dict(zip(d.padded_keys(32).tolist(), d.counts.tolist()))
    ck = c.padded_keys(32).tolist()
    assert all(dd.get(k) == v for k, v in zip(ck[:50000], c.counts.tolist()[:50000]))


def test_compact_export_agrees_with_offsets(golden):
    body = corpus_body(golden, "republic")
    m = cb().train(body, MINTOKENS=2, MAXLENGTH=5, QUIET=1)
    keys, off, counts, _ = m.export()
    ckeys, lens, ccounts = m.export_compact()
    assert np.array_equal(keys, ckeys) and np.array_equal(counts, ccounts)
    assert np.array_equal(np.diff(off.astype(np.int64)), lens.astype(np.int64))


def test_error_paths():
    lib = cb()
    with pytest.raises(lib.ColibriError):
        lib.train(b"", MINTOKENS=2)  # reference: "file is empty?" -> InternalError (src/pattern.cpp:520-523)
    with pytest.raises(lib.ColibriError) as ei:
        lib.train(bytes([6, 3, 7, 0]), MINTOKENS=1)  # reserved skip class in running text
    assert ei.value.code == 2
    with pytest.raises(lib.ColibriError) as ei:
        lib.train(bytes([6, 0x80, 0x00, 7, 0]), MINTOKENS=1)  # non-canonical varint
    assert ei.value.code == 4
    with pytest.raises(lib.ColibriError):
        lib.train(bytes([6, 7, 0]), DOSKIPGRAMS=1, DOSKIPGRAMS_EXHAUSTIVE=1)  # patternmodel.h:958-963


@pytest.mark.parametrize("seed", range(240))
def test_gpu_equals_oracle_on_random_input(seed, monkeypatch):
    """Random small corpora (empty/ragged sentences, unknown class, missing final delimiter) x random supported option sets; the forced
    code paths of conftest.FORCED_PATHS take turns (seed 0..119: default knobs, as in round 1)."""
    import random

    from test_oracle_vs_ref_random import random_case, random_corpus

    force_path(monkeypatch, "default" if seed < 120 else list(FORCED_PATHS)[1 + seed % (len(FORCED_PATHS) - 1)])
    rng = random.Random(5000 + seed % 120 + (7000 if seed >= 120 else 0))
    body = random_corpus(rng)
    if not body:
        pytest.skip("empty corpus")
    unindexed, skipgrams, cli = random_case(rng)
    names = {"t": "MINTOKENS", "l": "MAXLENGTH", "m": "MINLENGTH", "y": "MINTOKENS_SKIPGRAMS", "T": "MINSKIPTYPES", "W": "MINTOKENS_UNIGRAMS"}
    onames = {"t": "mintokens", "l": "maxlength", "m": "minlength", "y": "mintokens_skipgrams", "T": "minskiptypes", "W": "mintokens_unigrams"}
    streamed = 1 if (unindexed and not skipgrams) else 0
    okw = {onames[k]: v for k, v in cli.items()}
    okw.update(indexed=0 if unindexed else 1, doskipgrams_exhaustive=1 if skipgrams else 0, streamed=streamed)
    try:
        want = oracle.train(body, **okw)
    except RuntimeError:
        with pytest.raises(cb().ColibriError):
            cb().train(body, QUIET=1, model_type=10 if unindexed else 20, DOSKIPGRAMS_EXHAUSTIVE=int(skipgrams), streamed=streamed, **{names[k]: v for k, v in cli.items()})
        return
    try:
        m = cb().train(body, QUIET=1, model_type=10 if unindexed else 20, DOSKIPGRAMS_EXHAUSTIVE=int(skipgrams), streamed=streamed, **{names[k]: v for k, v in cli.items()})
    except cb().ColibriError as e:
        assert e.code == 2, e  # only "outside the accelerated subset" is an acceptable refusal
        pytest.skip("unsupported on the device path: %s" % e)
    got = to_flat(m)
    assert (got.tokens, got.types, len(got), got.maxn, got.minn, got.hasskipgrams) == (want.tokens, want.types, len(want), want.maxn, want.minn, want.hasskipgrams), (cli, unindexed, skipgrams)
    assert got.passes == want.passes
    assert got.same_patterns(want)


@pytest.mark.parametrize("seed", range(60))
def test_gpu_indexed_skipgrams_equal_oracle_on_random_input(seed, monkeypatch):
    """IndexedPatternModel::trainskipgrams (reference :2969-3010) + skip-type pruning on random corpora: patterns, counts, occurrence lists."""
    import random

    from test_oracle_vs_ref_random import random_corpus

    force_path(monkeypatch, list(FORCED_PATHS)[seed % len(FORCED_PATHS)])
    rng = random.Random(9000 + seed)
    body = random_corpus(rng)
    if not body:
        pytest.skip("empty corpus")
    kw = dict(MINTOKENS=rng.choice([2, 2, 3]), MAXLENGTH=rng.choice([3, 4, 5, 6, 8]), MINSKIPTYPES=rng.choice([1, 2, 2, 3]))
    if rng.random() < 0.3:
        kw["MINLENGTH"] = rng.randint(1, 3)
    okw = dict(mintokens=kw["MINTOKENS"], maxlength=kw["MAXLENGTH"], minskiptypes=kw["MINSKIPTYPES"], minlength=kw.get("MINLENGTH", 1), indexed=1, doskipgrams=1, streamed=0)
    try:
        want = oracle.train(body, **okw)
    except RuntimeError:
        with pytest.raises(cb().ColibriError):
            cb().train(body, QUIET=1, model_type=20, DOSKIPGRAMS=1, streamed=0, **kw)
        return
    m = cb().train(body, QUIET=1, model_type=20, DOSKIPGRAMS=1, streamed=0, **kw)
    got = to_flat(m)
    assert (got.tokens, got.types, len(got), got.maxn, got.minn, got.hasskipgrams) == (want.tokens, want.types, len(want), want.maxn, want.minn, want.hasskipgrams), kw
    assert got.passes == want.passes
    assert got.same_patterns(want)


@pytest.mark.parametrize("dense", [8, 40, 3000, 100000])
def test_gpu_dense_pairs_match_oracle(golden, dense, monkeypatch):
    """Level 2 with directly addressed slots for pairs of frequent classes (count_ngrams_kernel, `dense`): forced on for small corpora and
    for several sizes of the square, so that windows split between the dense and the hashed part of the table in different ways."""
    monkeypatch.setenv("COLIBRI_B200_DENSE_MIN", "0")
    monkeypatch.setenv("COLIBRI_B200_DENSE", str(dense))
    runs = [
        ("hamlet", dict(mintokens=2, maxlength=5, streamed=1)),
        ("hamlet", dict(mintokens=1, maxlength=3, streamed=1)),
        ("hamlet", dict(mintokens=2, maxlength=4, indexed=1, streamed=0)),
        ("hamlet", dict(mintokens=2, maxlength=5, doskipgrams_exhaustive=1, streamed=0)),
        ("hamlet", dict(mintokens=2, maxlength=5, indexed=1, doskipgrams=1, streamed=0)),
        ("threebyte", dict(mintokens=2, maxlength=5, streamed=1)),
        ("republic", dict(mintokens=2, maxlength=5, streamed=1)),
        ("republic", dict(mintokens=3, maxlength=4, indexed=1, streamed=0)),
        ("zipf300k_phr", dict(mintokens=2, maxlength=5, doskipgrams_exhaustive=1, streamed=0)),
        ("zipf2m", dict(mintokens=2, maxlength=5, streamed=1)),
    ]
    for name, okw in runs:
        body = corpus_body(golden, name)
        want = oracle.train(body, **okw)
        o = cb().PatternModelOptions(MINTOKENS=okw["mintokens"], MAXLENGTH=okw["maxlength"], DOSKIPGRAMS_EXHAUSTIVE=okw.get("doskipgrams_exhaustive", 0),
                                     DOSKIPGRAMS=okw.get("doskipgrams", 0), model_type=20 if okw.get("indexed") else 10, streamed=okw["streamed"], QUIET=1)
        m = cb().train(body, o)
        assert (m.tokens(), m.types(), len(m)) == (want.tokens, want.types, len(want)), (name, okw)
        assert m.passes() == want.passes, (name, okw)
        assert to_flat(m).same_patterns(want), (name, okw)


# ---- the benchmark corpus itself, pinned to the unmodified reference (tests/golden/golden_bench.json, written by make_golden_bench.py)
def _bench_golden():
    import json

    from conftest import GOLDEN_DIR

    with open(os.path.join(GOLDEN_DIR, "golden_bench.json")) as f:
        return json.load(f)


def _check_against_bench_fixture(name, skip, monkeypatch, path):
    if name not in _bench_golden():
        pytest.skip("tests/golden/golden_bench.json has no case %s yet (make_golden_bench.py)" % name)
    g = _bench_golden()[name]
    force_path(monkeypatch, path)
    gen = g["generator"]
    corpus = cb().Corpus.synthetic(gen["ntokens"], vocab=gen["vocab"], seed=gen["seed"], mean_sentence=gen["mean_sentence"])
    assert corpus.nbytes == g["corpus_bytes"]  # the device generator produced the bytes the reference was run on
    m = cb().train(corpus, MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=int(skip), streamed=0 if skip else 1, QUIET=1)
    assert (m.tokens(), m.types(), len(m)) == (g["tokens"], g["types"], g["patterns"])
    # the reference's progress lines: (found n-grams, found skipgram occurrences, pruned, kept) per pass
    assert [(p[1], p[3]) for p in m.passes()] == [(p[0], p[2]) for p in g["passes_found_skip_pruned_kept"]]
    got = to_flat(m)
    n_of = np.add.reduceat((got.keys < 128).astype(np.int64), got.key_off[:-1].astype(np.int64))
    per_len = {int(n): [int((n_of == n).sum()), int(got.counts[n_of == n].astype(np.int64).sum())] for n in np.unique(n_of)}
    assert per_len == {int(k): v for k, v in g["per_length_patterns_occurrences"].items()}
    assert len(m.to_bytes()) == g["modelfile_bytes"]
    assert got.digest() == g["digest"]  # every pattern and every count, canonical order
    return m


@pytest.mark.parametrize("path", ["default", "bench"])
def test_bench_corpus_10m_prefix_matches_reference_fixture(path, monkeypatch):
    """10 M-token prefix of the benchmark stream, -u -t 2 -l 5, against the unmodified reference's model (digest, passes, file size);
    `bench` forces filter + hot-key cache + dense pairs + list mode, which a 10 M-token corpus does not reach on its own."""
    _check_against_bench_fixture("zipf10m", False, monkeypatch, path)


def test_bench_corpus_10m_prefix_with_skipgrams_matches_reference_fixture(monkeypatch):
    """BASELINE.json configs[2] shape (exhaustive skipgrams) on the 10 M-token prefix against the unmodified reference."""
    _check_against_bench_fixture("zipf10m_skip", True, monkeypatch, "bench")


@pytest.mark.slow
def test_bench_corpus_100m_matches_reference_fixture(monkeypatch):
    """BASELINE.json configs[1]: the corpus bench.py times (100 M tokens, V = 100 000, seed 1), default knobs -- i.e. exactly the kernels and
    thresholds the benchmark runs -- bit-exact against the model the unmodified reference wrote for the same bytes (12 CPU-minutes, once)."""
    m = _check_against_bench_fixture("zipf100m", False, monkeypatch, "default")
    assert m.level(2)["singletons"] > 0  # the occurrence filter was on ...
    assert m.level(5)["items"] < m.counters()["positions"] // 4  # ... and the sparse levels ran from a position list


def test_device_checksum_matches_host_restatement(golden):
    """colibri_b200_model_checksum (what bench.py compares across GPU counts and against the fixtures) against tests/checksum.py on the same model."""
    from checksum import flat_checksum

    for name, kw in (("hamlet", dict(MINTOKENS=2, MAXLENGTH=5)), ("republic", dict(MINTOKENS=2, MAXLENGTH=5, model_type=20, streamed=0)),
                     ("zipf300k_phr", dict(MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0)), ("single_token", dict(MINTOKENS=1, MAXLENGTH=3))):
        m = cb().train(corpus_body(golden, name), QUIET=1, **kw)
        assert m.checksum() == flat_checksum(to_flat(m)), name


@pytest.mark.parametrize("path", ["default", "bench"])
def test_streamed_train_export_equals_train_then_export(golden, path, monkeypatch):
    """colibri_b200_train_export (levels copied to the host while the next one counts) delivers the same model as train() + export, for
    streamable option sets, for those where a level may still be dropped at the end (MINLENGTH > 1: exported after the last level),
    with skipgram segments, and through the grow-and-retry path of too small buffers."""
    force_path(monkeypatch, path)
    runs = [("hamlet", dict(MINTOKENS=2, MAXLENGTH=5)), ("hamlet", dict(MINTOKENS=1, MAXLENGTH=3)), ("hamlet", dict(MINTOKENS=2, MAXLENGTH=6, MINLENGTH=3)),
            ("republic", dict(MINTOKENS=2, MAXLENGTH=5)), ("zipf300k_phr", dict(MINTOKENS=2, MAXLENGTH=5, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0)),
            ("zipf2m", dict(MINTOKENS=2, MAXLENGTH=5)), ("single_token", dict(MINTOKENS=2, MAXLENGTH=3)), ("noeos", dict(MINTOKENS=1, MAXLENGTH=3))]
    for name, kw in runs:
        body = corpus_body(golden, name)
        m = cb().train(body, QUIET=1, **kw)
        want = to_flat(m)
        keys, lens, counts, sm = cb().train_export(body, QUIET=1, **kw)
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens.astype(np.uint64))
        got = oracle.FlatModel(keys, off, counts, sm["tokens"], sm["types"], sm["maxn"], sm["minn"], sm["hasskipgrams"])
        assert (got.tokens, got.types, len(got), got.maxn, got.minn, got.hasskipgrams) == (want.tokens, want.types, len(want), want.maxn, want.minn, want.hasskipgrams), (name, kw)
        assert sm["passes"] == m.passes(), (name, kw)
        assert got.same_patterns(want), (name, kw)
    with pytest.raises(cb().ColibriError) as ei:
        cb().train_export(corpus_body(golden, "hamlet"), MINTOKENS=2, MAXLENGTH=3, model_type=20, streamed=0, QUIET=1)
    assert ei.value.code == 2


@pytest.mark.parametrize("chunk", [4096, 8192, 65536])
def test_tokeniser_following_a_chunked_copy_matches_oracle(golden, chunk, monkeypatch):
    """Host corpora above four chunks are copied in chunks and tokenised chunk by chunk behind the copy (engine.cu: Trainer::tokenise, `piped`);
    forced here with tiny chunks: tokens that straddle a chunk border, a corpus without its final delimiter (streamed and preloaded), the
    dense square's spare room behind an over-sized token array."""
    monkeypatch.setenv("COLIBRI_B200_H2D_CHUNK", str(chunk))
    runs = [("republic", dict(mintokens=2, maxlength=5), {}), ("zipf300k_phr", dict(mintokens=2, maxlength=4), {}),
            ("zipf2m", dict(mintokens=2, maxlength=3), {"COLIBRI_B200_DENSE_MIN": "0", "COLIBRI_B200_DENSE": "64", "COLIBRI_B200_PART_MIN": "0"}),
            ("noeos", dict(mintokens=1, maxlength=3, streamed=1), {}), ("noeos", dict(mintokens=1, maxlength=3, streamed=0), {})]
    for name, okw, env in runs:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        body = corpus_body(golden, name)
        if name == "noeos":
            body = body * 4000  # long enough to be chunked; still without the final delimiter
        want = oracle.train(body, **okw)
        m = cb().train(body, QUIET=1, MINTOKENS=okw["mintokens"], MAXLENGTH=okw["maxlength"], streamed=okw.get("streamed", 1))
        got = to_flat(m)
        assert (got.tokens, got.types, len(got)) == (want.tokens, want.types, len(want)), (name, chunk)
        assert got.passes == want.passes and got.same_patterns(want), (name, chunk)
        keys, lens, counts, sm = cb().train_export(body, QUIET=1, MINTOKENS=okw["mintokens"], MAXLENGTH=okw["maxlength"], streamed=okw.get("streamed", 1))
        assert len(lens) == len(want) and sm["passes"] == want.passes, (name, chunk)
        for k in env:
            monkeypatch.delenv(k)


def test_partition_overflow_falls_back_to_the_table_path(monkeypatch):
    """A partition with more distinct keys than its shared-memory table holds raises a flag and the level is rerun on the HBM table (engine.cu:
    level_partitioned -> overflow).  Forced by sizing the partitions for a hundredth of the windows; with and without the dense square."""
    monkeypatch.setenv("COLIBRI_B200_PART_MIN", "0")
    monkeypatch.setenv("COLIBRI_B200_PART_ALL", "1")
    monkeypatch.setenv("COLIBRI_B200_PART_SCALE", "0.01")
    corpus = cb().Corpus.synthetic(3000000, vocab=100000, seed=11, mean_sentence=22)
    want = oracle.train(corpus.download(), mintokens=2, maxlength=4)
    for dense in ("0", "64"):
        monkeypatch.setenv("COLIBRI_B200_DENSE_MIN", "0")
        monkeypatch.setenv("COLIBRI_B200_DENSE", dense)
        m = cb().train(corpus, MINTOKENS=2, MAXLENGTH=4, QUIET=1)
        got = to_flat(m)
        assert got.passes == want.passes and got.same_patterns(want)
        assert m.level(2)["path"] == "table"  # the partitioned attempt was abandoned
