"""Randomised pinning of the oracle: small random corpora x random option sets, oracle.c vs the UNMODIFIED reference binary
(oracle/_ref/ref_train, compiled from /root/reference by `make -C oracle ref`).  Skipped where the reference build is absent."""
import os
import random
import tempfile

import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref (the compiled reference) is not present")

CLI2OPT = {"t": "mintokens", "l": "maxlength", "m": "minlength", "y": "mintokens_skipgrams", "T": "minskiptypes", "W": "mintokens_unigrams"}


def random_corpus(rng):
    vocab = rng.choice([3, 5, 12, 40, 300, 20000])
    sentences = []
    for _ in range(rng.randint(1, 60)):
        n = rng.choice([0, 0, 1, 2, 3, 5, 8, 13, 30]) if rng.random() < 0.3 else rng.randint(1, 25)
        # classes 6.. (never the reserved 0-5); a Zipf-ish skew so that patterns repeat; occasionally the unknown class 2
        sent = [2 if rng.random() < 0.02 else 6 + min(int(rng.paretovariate(1.1)) - 1, vocab - 1) for _ in range(n)]
        sentences.append(sent)
    body = oracle.encode_corpus(sentences)
    if rng.random() < 0.25 and len(body) > 1 and body[-2] != 0:
        body = body[:-1]  # drop the final end-of-sentence marker (reader quirk, src/pattern.cpp:483-587)
    return body


def random_case(rng):
    unindexed = rng.random() < 0.7
    skipgrams = unindexed and rng.random() < 0.4
    cli = {"t": rng.choice([1, 2, 2, 2, 3, 4]), "l": rng.choice([1, 2, 3, 4, 5, 6, 8])}
    if rng.random() < 0.25 and not skipgrams and cli["t"] > 1:
        cli["m"] = rng.randint(1, min(3, cli["l"]))
    if skipgrams:
        if rng.random() < 0.5:
            cli["y"] = rng.randint(1, 4)
        if rng.random() < 0.5:
            cli["T"] = rng.choice([1, 2])
    if rng.random() < 0.2 and cli["t"] > 1 and cli.get("m", 1) == 1:
        cli["W"] = rng.randint(1, 4)
    return unindexed, skipgrams, cli


@pytest.mark.parametrize("seed", range(200))
def test_oracle_equals_reference_on_random_input(seed):
    rng = random.Random(1000 + seed)
    body = random_corpus(rng)
    if not body:
        pytest.skip("empty corpus")
    unindexed, skipgrams, cli = random_case(rng)
    opts = {CLI2OPT[k]: v for k, v in cli.items()}
    opts["indexed"] = 0 if unindexed else 1
    opts["doskipgrams_exhaustive"] = 1 if skipgrams else 0
    opts["streamed"] = 1 if (unindexed and not skipgrams) else 0  # what the CLI does (src/patternmodeller.cpp:721-754)
    with tempfile.TemporaryDirectory() as td:
        cpath, mpath = os.path.join(td, "c.colibri.dat"), os.path.join(td, "m")
        with open(cpath, "wb") as f:
            f.write(b"\xa2\x02" + body)
        try:
            st, err = oracle.ref_train(cpath, mpath, unindexed=unindexed, skipgrams=skipgrams, timeout=60, **cli)
        except RuntimeError as e:
            # the reference itself rejects some inputs (e.g. a corpus consisting of one unterminated delimiter); the oracle must refuse too
            with pytest.raises(RuntimeError):
                oracle.train(body, **opts)
            return
        ref = oracle.parse_modelfile(open(mpath, "rb").read())
    try:
        mine = oracle.train(body, **opts)
    except RuntimeError as e:
        if "not restated" in str(e):
            pytest.skip(str(e))
        raise
    assert (mine.tokens, mine.types, len(mine), mine.maxn, mine.minn, int(mine.hasskipgrams)) == (st["tokens"], st["types"], st["patterns"], st["maxn"], st["minn"], st["hasskipgrams"]), (cli, unindexed, skipgrams)
    assert [(p[1], p[3]) for p in mine.passes] == [(p[0], p[2]) for p in oracle.parse_ref_passes(err)]
    assert mine.same_patterns(ref)


# ---- training under a constraint model and model loading with the options as filters (8f-3): the oracle against the
# ---- UNMODIFIED reference CLI (oracle/_ref/colibri-patternmodeller -j / -i -I)
def _plain_corpus(rng, vocab_choices):
    sentences = []
    for _ in range(rng.randint(1, 60)):
        n = rng.randint(0, rng.choice([4, 12, 30]))
        sentences.append([6 + min(int(rng.paretovariate(1.1)) - 1, rng.choice(vocab_choices) - 1) if rng.random() < 0.9 else 6 + rng.randrange(max(vocab_choices)) for _ in range(n)])
    body = oracle.encode_corpus(sentences)
    return body if body.strip(b"\0") else bytes([6, 7, 0])


@pytest.mark.parametrize("seed", range(80))
def test_oracle_constrained_equals_reference_cli_on_random_input(seed):
    rng = random.Random(5000 + seed)
    body = _plain_corpus(rng, [5, 30, 200, 20000])
    body2 = body if rng.random() < 0.5 else _plain_corpus(rng, [5, 30, 200])
    t1, l1, m1, idx1 = rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 3, 5, 8]), rng.choice([1, 1, 1, 2]), rng.random() < 0.4
    t, l, m = rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 3, 5, 8, 100]), rng.choice([1, 1, 1, 2, 3])
    indexed, inplace = rng.random() < 0.5, rng.random() < 0.5
    with tempfile.TemporaryDirectory() as td:
        cpath, c2path, s1, out = (os.path.join(td, x) for x in ("c.colibri.dat", "c2.colibri.dat", "s1.model", "out.model"))
        for path, b in ((cpath, body), (c2path, body2)):
            with open(path, "wb") as f:
                f.write(b"\xa2\x02" + b)
        rc, err = oracle.ref_cli(["-f", c2path, "-t", t1, "-l", l1, "-m", m1, "-o", s1] + ([] if idx1 else ["-u"]), timeout=60)
        assert rc == 0, err
        args = ["-f", cpath] + (["-i", s1, "-I"] if inplace else ["-j", s1]) + ["-t", t, "-l", l, "-m", m, "-o", out] + ([] if indexed else ["-u"])
        rc, err = oracle.ref_cli(args, timeout=60)
        assert rc == 0, err
        ref = oracle.parse_modelfile(open(out, "rb").read())
        stage1 = open(s1, "rb").read()
    if inplace:  # src/patternmodeller.cpp:804-821: load AS the output type with DORESET, widen the length window, corpus preloaded
        cm = oracle.load_model(stage1, mintokens=t, minlength=m, maxlength=l, doreset=1, indexed=int(indexed))
        f = cm.flat()
        mine = oracle.train_constrained(body, cm, inplace=True, mintokens=t, maxlength=max(l, f.maxn), minlength=min(m, f.minn), indexed=int(indexed), streamed=0)
    else:  # :713-721 PatternSetModel(file, options); an unindexed output model streams the corpus
        cm = oracle.load_model(stage1, mintokens=t, minlength=m, maxlength=l, indexed=0)
        mine = oracle.train_constrained(body, cm, inplace=False, mintokens=t, maxlength=l, minlength=m, indexed=int(indexed), streamed=0 if indexed else 1)
    assert (mine.tokens, mine.types, len(mine)) == (ref.tokens, ref.types, len(ref)), args
    assert [(p[1], p[3]) for p in mine.passes] == [(p[0], p[2]) for p in oracle.parse_ref_passes(err)], args
    assert mine.same_patterns(ref), args


REF_RELATIONS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_relations")


@pytest.mark.skipif(not os.path.exists(REF_RELATIONS), reason="oracle/_ref/ref_relations (the reference's relation queries) is not built")
@pytest.mark.parametrize("seed", range(40))
def test_oracle_relations_equal_reference_on_random_input(seed):
    """getreverseindex, getrightcooc / getleftcooc, getcooc (defaults and occurrencethreshold 2 + ordersignificant) and the group statistics of the
    oracle against the unmodified reference (oracle/ref_relations.cpp) on random corpora: what pins the device's relation queries."""
    import subprocess

    rng = random.Random(5000 + seed)
    body = b""
    while not body:
        vocab = rng.choice([3, 6, 15, 60])
        sentences = [[6 + min(int(rng.paretovariate(1.1)) - 1, vocab - 1) for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 13, 21]))] for _ in range(rng.randint(1, 40))]
        body = bytes(oracle.encode_corpus(sentences))
    t, l = rng.choice([2, 2, 3]), rng.choice([2, 3, 4, 5])  # (threshold 1 trips the reference's statistics cache, tests/golden/make_golden_stats.py)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.colibri.dat")
        with open(path, "wb") as f:
            f.write(b"\xa2\x02" + body)
        rel = subprocess.run([REF_RELATIONS, "-f", path, "-t", str(t), "-l", str(l)], capture_output=True, text=True, check=True).stdout
        sta = subprocess.run([REF_RELATIONS, "-f", path, "-t", str(t), "-l", str(l), "-S"], capture_output=True, text=True, check=True).stdout
    patterns = oracle.train(body, mintokens=t, maxlength=l, indexed=1, streamed=0).as_dict()
    ref = {k: [] for k in "GRLCO"}
    for line in rel.splitlines():
        p = line.split()
        if p[0] == "G":
            ref["G"].append((int(p[1]), int(p[2]), sorted(p[3:])))
        elif p[0] in "RLCO":
            ref[p[0]].append((p[1], p[2], int(p[3])))
    if not patterns:
        assert not any(ref[k] for k in "RLCO")
        return
    rindex = oracle.reverse_index(body, patterns)
    assert sorted((s, tk, sorted(k.hex() for k in v)) for (s, tk), v in rindex.items()) == sorted(ref["G"])
    flat = lambda rel: sorted((p.hex(), q.hex(), j) for (p, q), j in rel.items())  # noqa: E731
    assert flat(oracle.cooc(body, patterns, left=False)) == sorted(ref["R"])
    assert flat(oracle.cooc(body, patterns, left=True)) == sorted(ref["L"])
    assert flat(oracle.cooc_both(body, patterns)) == sorted(ref["C"])
    assert flat(oracle.cooc_both(body, patterns, occurrencethreshold=2, ordersignificant=True)) == sorted(ref["O"])
    got = oracle.group_stats(patterns)
    for line in sta.splitlines():
        p = line.split()
        if p[0] == "S":
            assert got.get((int(p[1]), int(p[2])), (0, 0, 0)) == (int(p[3]), int(p[4]), int(p[5])), line
