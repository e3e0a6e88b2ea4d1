"""The multi-GPU phases on real GPUs through NCCL.  world_size 1 (all-to-all with itself) runs on any GPU box and exercises
every shard kernel; world_size 2 needs two GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


DENSE = {"COLIBRI_B200_DENSE_MIN": "0", "COLIBRI_B200_DENSE": "64"}  # the dense square of level 2 forced onto small shards (bench.py has it at 100 M tokens per GPU)


def _run(world, per, vocab, seed, maxlength, mintokens, port, mode="nccl", extra="", env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(per), str(vocab), str(seed), str(maxlength), str(mintokens), mode] + ([extra] if extra else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_RESULT OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("per,vocab,seed,maxlength,mintokens", [(300000, 20000, 4, 5, 2), (120000, 3000, 8, 6, 3)])
def test_shard_phases_world1(per, vocab, seed, maxlength, mintokens):
    _run(1, per, vocab, seed, maxlength, mintokens, 29711)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_shard_phases_world2(mode, dense):
    import colibri_core_b200 as cb

    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, 400000, 30000, 6, 5, 2, 29713, mode, env=DENSE if dense else None)


@pytest.mark.parametrize("mode,extra", [("nccl", ""), ("p2p", ""), ("p2p", "skipgrams")])
def test_shard_dense_pairs_world1(mode, extra):
    """Level 2 with the dense square on the sharded path: pairs of frequent classes are counted locally and summed by one all-reduce instead of
    being shipped; their ids (cell + 1) live below the ids of the hashed n-grams and feed level 3 and the skipgrams."""
    _run(1, 300000, 20000, 4, 5, 2, 29725, mode, extra, env=DENSE)


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_shard_skipgrams_world1(mode):
    """Config-3 shape on the sharded path: n-grams + exhaustive skipgrams."""
    _run(1, 200000, 8000, 14, 5, 2, 29717, mode, "skipgrams")


def test_shard_skipgrams_world2():
    import colibri_core_b200 as cb

    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, 250000, 8000, 15, 5, 2, 29719, "p2p", "skipgrams")


def test_shard_phases_world1_peer_stores():
    """NVLink peer-store mode with a single rank: the kernels store into the rank's own symmetric buffers."""
    _run(1, 300000, 20000, 4, 5, 2, 29715, "p2p")


def test_constrained_sharded_world1():
    """Sharded constrained training (count / all-reduce / finish) with a single rank."""
    _run(1, 300000, 20000, 21, 5, 2, 29721, "constrained")


def test_constrained_sharded_world2():
    import colibri_core_b200 as cb

    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, 300000, 20000, 23, 5, 2, 29723, "constrained")


def _merged(shares):
    merged, head = {}, None
    for m in shares:
        keys, off, counts, _ = m.export()
        kb = keys.tobytes()
        for i in range(len(counts)):
            k = kb[int(off[i]):int(off[i + 1])]
            assert k not in merged, "pattern exported twice"
            merged[k] = int(counts[i])
        h = (m.tokens(), m.types(), m.passes(), m.maxlength(), m.minlength())
        assert head is None or head == h
        head = h
    return merged, head


@pytest.mark.parametrize("ndev", [1, 2])
@pytest.mark.parametrize("dense", [False, True])
def test_train_multi_in_process(ndev, dense, monkeypatch):
    """colibri_b200_train_multi: the shard phases driven by one host thread per device inside ONE process (what the C++ CLI's -d 0-7 uses), peer-mapped
    buffers instead of symmetric memory, host barriers.  One device exercises the whole machinery on any box; two need two GPUs."""
    import colibri_core_b200 as cb
    import oracle

    if cb.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    for k, v in (DENSE if dense else {}).items():
        monkeypatch.setenv(k, v)
    for kw, t, l, ml in ((dict(ntokens=400000, vocab=30000, seed=6, mean_sentence=15, phrase_permille=150, nphrases=500), 2, 5, 1),
                         (dict(ntokens=60000, vocab=800, seed=9, mean_sentence=7), 3, 6, 1), (dict(ntokens=30000, vocab=500, seed=10, mean_sentence=12), 1, 3, 1),
                         (dict(ntokens=80000, vocab=900, seed=11, mean_sentence=9), 2, 5, 3), (dict(ntokens=80000, vocab=900, seed=11, mean_sentence=9), 2, 4, 2)):
        body = oracle.synth_corpus(**kw).tobytes()
        want = oracle.train(body, mintokens=t, maxlength=l, minlength=ml)
        shares = cb.train_multi(body, list(range(ndev)), MINTOKENS=t, MAXLENGTH=l, MINLENGTH=ml, QUIET=1)
        merged, head = _merged(shares)
        assert merged == want.as_dict(), (kw, t, l, ml)
        assert head == (want.tokens, want.types, want.passes, want.maxn, want.minn), (kw, t, l, ml)
    with pytest.raises(cb.ColibriError) as ei:
        cb.train_multi(body, list(range(ndev)), MINTOKENS=2, MAXLENGTH=4, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0, QUIET=1)
    assert ei.value.code == 2


def test_cli_on_two_devices_writes_the_single_gpu_model(tmp_path):
    """colibri-patternmodeller -d 0-1: the model file of a two-GPU run has the digest of the one-GPU run (and of the reference golden)."""
    import colibri_core_b200 as cb
    import oracle

    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    cli = os.path.join(ROOT, "colibri-core_b200", "bin", "colibri-patternmodeller")
    corpus = str(tmp_path / "c.colibri.dat")
    body = oracle.synth_corpus(500000, vocab=20000, seed=12, mean_sentence=18).tobytes()
    with open(corpus, "wb") as f:
        f.write(b"\xa2\x02" + body)
    outs = []
    for spec in ("0", "0-1", "1,0"):
        out = str(tmp_path / ("m%s.patternmodel" % spec.replace(",", "_")))
        r = subprocess.run([cli, "-f", corpus, "-u", "-t", "2", "-l", "5", "-o", out, "-d", spec], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(oracle.parse_modelfile(open(out, "rb").read()))
    want = oracle.train(body, mintokens=2, maxlength=5)
    for got in outs:
        assert got.same_patterns(want) and (got.tokens, got.types) == (want.tokens, want.types)
