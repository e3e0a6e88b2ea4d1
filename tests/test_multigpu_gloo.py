"""The N>1 path on CPU: the real orchestration (colibri-core_b200/multigpu.py: ownership routing, variable all-to-all,
reply routing, global statistics, early stop) over torch.distributed/gloo with world_size 2 and 3, driven by a numpy
stand-in for the per-rank CUDA phases.  The merged model must equal the oracle's model of the concatenated shards."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NumpyShardEngine:
    """Test double with the interface of multigpu.CudaShardEngine; same record formats (16-byte records, 8-byte replies)."""

    def __init__(self, body, mintokens, rank, world):
        import oracle

        self.t, self.rank, self.world = mintokens, rank, world
        tok, i = [], 0
        while i < len(body):
            v, ln = oracle.bytestoint(body[i:i + 8])
            tok.append(v)
            i += ln
        tok.append(0)
        self.tok = np.array(tok, dtype=np.int64)
        self.segs = []  # (n, pos or class, count)
        self.skipsegs = []  # (n, pos, count, mask)
        self.ids_keep = {}
        self.level = 1

    def new_buffer(self, nwords):
        return torch.empty(max(int(nwords), 1), dtype=torch.int32)

    def sync(self):
        pass

    def info(self):
        return {"tokens": int((self.tok != 0).sum()), "maxclass": int(self.tok.max()), "positions": len(self.tok), "launches": 0}

    def unigram_counts(self, nclasses):
        c = np.bincount(self.tok[self.tok != 0], minlength=nclasses).astype(np.int32)
        self.nclasses = nclasses
        return torch.from_numpy(c)

    def unigram_finish(self, global_counts, global_tokens):
        c = global_counts.numpy()[: self.nclasses].astype(np.int64)
        found, kept, occ = int((c > 0).sum()), int((c >= self.t).sum()), int(c[c >= self.t].sum())
        for cls in np.nonzero(c >= self.t)[0]:
            if cls % self.world == self.rank and c[cls] > 0:
                self.segs.append((1, int(cls), int(c[cls])))
        self.prev = np.where((self.tok != 0) & (c[self.tok] >= self.t), self.tok, 0)
        self.ids_keep[1] = self.prev
        return found, kept, occ

    def _owner(self, a, b):
        return (a * 1000003 + b * 7919) % self.world

    def level_split_count(self, n):
        """valid windows grouped by owner, corpus order inside a group (same contract as the CUDA split kernels)"""
        groups = [[] for _ in range(self.world)]
        for p in range(len(self.tok) - 1):
            a, b = int(self.prev[p]), int(self.prev[p + 1])
            if a and b:
                groups[self._owner(a, b)].append(p)
        self.groups = groups
        self.send_base = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.int64)
        return [len(g) for g in groups], int(self.send_base[-1])

    def level_split_write(self, nsend):
        rec = np.zeros((max(nsend, 1), 2), dtype=np.uint32)
        j = 0
        for g in self.groups:
            for p in g:
                rec[j] = (int(self.prev[p + 1]), int(self.prev[p]))  # key = a << 32 | b, little-endian words
                j += 1
        return torch.from_numpy(rec.view(np.int32).reshape(-1))

    def level_owner(self, recv, recv_counts):
        nrecv = sum(recv_counts)
        rec = recv.numpy().view(np.uint32)[: nrecv * 2].reshape(-1, 2)
        table, rid = {}, []
        for i in range(nrecv):
            e = table.setdefault((int(rec[i, 1]), int(rec[i, 0])), [0, i, len(table)])
            e[0] += 1
            rid.append(e)
        reply = np.zeros(max(nrecv, 1), dtype=np.uint32)
        for i, e in enumerate(rid):
            if e[0] >= self.t:
                reply[i] = e[2] * self.world + self.rank + 1
        kept = [e for e in table.values() if e[0] >= self.t]
        src_base = np.concatenate([[0], np.cumsum(recv_counts)])
        self.surv = [[] for _ in range(self.world)]
        for e in kept:
            src = int(np.searchsorted(src_base, e[1], side="right") - 1)
            self.surv[src].append((e[1] - int(src_base[src]), e[0]))
        return torch.from_numpy(reply.view(np.int32)), (len(table), len(kept), sum(e[0] for e in kept)), [len(x) for x in self.surv]

    def level_owner_survivors(self, nsurv):
        rec = np.zeros((max(nsurv, 1), 2), dtype=np.uint32)
        j = 0
        for grp in self.surv:
            for idx, cnt in grp:
                rec[j] = (idx, cnt)
                j += 1
        return torch.from_numpy(rec.view(np.int32).reshape(-1))

    def level_finish(self, reply_back, surv, surv_counts):
        n = self.level + 1
        rep = reply_back.numpy().view(np.uint32)
        cur = np.zeros(len(self.tok), dtype=np.int64)
        j = 0
        for g in self.groups:
            for p in g:
                cur[p] = int(rep[j])
                j += 1
        srec = surv.numpy().view(np.uint32)[: sum(surv_counts) * 2].reshape(-1, 2)
        k = 0
        for owner, c in enumerate(surv_counts):
            for _ in range(c):
                self.segs.append((n, self.groups[owner][int(srec[k, 0])], int(srec[k, 1])))
                k += 1
        self.prev, self.level = cur, n
        self.ids_keep[n] = cur
        return int((cur != 0).sum())

    # ---- skipgrams of the level just finished (same record formats as the CUDA phases: 16-byte keys, 16-byte survivor records)
    def _skip_owner_of(self, k):
        return (k[0] * 31 + k[1] * 1000003 + k[2] * 7919 + k[3] * 104729) % self.world

    def skip_split_count(self):
        import oracle

        n = self.level
        self.ids_keep = getattr(self, "ids_keep", {})
        masks = oracle.skip_configurations(n, 3) if n >= 3 else []
        groups = [[] for _ in range(self.world)]
        prev = self.ids_keep[n - 1] if n >= 2 else None
        for p in range(len(self.tok) - 1):
            if not masks or not (prev[p] and prev[p + 1]):
                continue
            for mask in masks:
                parts, j = [], 0
                while j < n:
                    if (mask >> j) & 1:
                        j += 1
                        continue
                    k = j
                    while k < n and not (mask >> k) & 1:
                        k += 1
                    parts.append(int(self.ids_keep[k - j][p + j]))
                    j = k
                key = (mask, parts[0], parts[1], parts[2] if len(parts) > 2 else 0)
                groups[self._skip_owner_of(key)].append((key, p))
        self.sk_groups = groups
        return [len(g) for g in groups], sum(len(g) for g in groups)

    def skip_split_write(self, nsend):
        rec = np.zeros((max(nsend, 1), 4), dtype=np.uint32)
        j = 0
        for g in self.sk_groups:
            for (mask, a, b, c), _p in g:
                rec[j] = (a, mask, c, b)  # k0 = mask << 32 | a ; k1 = b << 32 | c  (little-endian words)
                j += 1
        return torch.from_numpy(rec.view(np.int32).reshape(-1))

    def skip_owner(self, recv, recv_counts):
        nrecv = sum(recv_counts)
        rec = recv.numpy().view(np.uint32)[: nrecv * 4].reshape(-1, 4)
        table = {}
        for i in range(nrecv):
            e = table.setdefault(tuple(int(x) for x in rec[i]), [0, i])
            e[0] += 1
        kept = {k: e for k, e in table.items() if e[0] >= self.t}
        src_base = np.concatenate([[0], np.cumsum(recv_counts)])
        self.sk_surv = [[] for _ in range(self.world)]
        for k, e in kept.items():
            src = int(np.searchsorted(src_base, e[1], side="right") - 1)
            self.sk_surv[src].append((e[1] - int(src_base[src]), e[0], k[1]))
        return (len(table), len(kept)), [len(x) for x in self.sk_surv]

    def skip_owner_survivors(self, nsurv):
        rec = np.zeros((max(nsurv, 1), 4), dtype=np.uint32)
        j = 0
        for grp in self.sk_surv:
            for idx, cnt, mask in grp:
                rec[j] = (idx, cnt, mask, 0)
                j += 1
        return torch.from_numpy(rec.view(np.int32).reshape(-1))

    def skip_finish(self, surv, surv_counts):
        srec = surv.numpy().view(np.uint32)[: sum(surv_counts) * 4].reshape(-1, 4)
        k = 0
        for owner, c in enumerate(surv_counts):
            for _ in range(c):
                self.skipsegs.append((self.level, self.sk_groups[owner][int(srec[k, 0])][1], int(srec[k, 1]), int(srec[k, 2])))
                k += 1

    def finish(self, passes, types, maxn, minn):
        import oracle

        out = {}
        for n, pos, cnt in self.segs:
            toks = [pos] if n == 1 else self.tok[pos:pos + n].tolist()
            out[b"".join(oracle.inttobytes(int(c)) for c in toks)] = cnt
        for n, pos, cnt, mask in self.skipsegs:
            ng = b"".join(oracle.inttobytes(int(c)) for c in self.tok[pos:pos + n].tolist())
            out[oracle.skipgram_collapse(ng, mask)] = cnt
        return out


def _worker(rank, world, port, bodies, mintokens, maxlength, results, skipgrams=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colibri_core_b200.multigpu as mg

    eng = NumpyShardEngine(bodies[rank], mintokens, rank, world)
    share, passes, head = mg.train_distributed(eng, dist, torch, mintokens, maxlength, skipgrams)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((share, passes, head), gathered, dst=0)
    if rank == 0:
        results.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mintokens,maxlength,seed,skipgrams", [(2, 2, 5, 3, False), (3, 2, 4, 5, False), (2, 3, 6, 7, False), (2, 1, 3, 9, False), (2, 2, 5, 11, True),
                                                                      (3, 2, 6, 13, True)])
def test_distributed_orchestration_matches_oracle(world, mintokens, maxlength, seed, skipgrams):
    import oracle

    per = 4000
    bodies = [oracle.synth_corpus(per, vocab=60, seed=seed, mean_sentence=9, phrase_permille=300, nphrases=20, first_token=r * per).tobytes() for r in range(world)]
    want = oracle.train(b"".join(bodies), mintokens=mintokens, maxlength=maxlength, doskipgrams_exhaustive=1 if skipgrams else 0, streamed=0 if skipgrams else 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + seed) % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, bodies, mintokens, maxlength, q, skipgrams)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for share, passes, head in gathered:
        assert not (set(share) & set(merged)), "a pattern was exported by two ranks"
        merged.update(share)
        assert passes == gathered[0][1] and head == gathered[0][2]  # every rank agrees on the global numbers
    assert merged == want.as_dict()
    _, passes, head = gathered[0]
    assert [tuple(p) for p in passes] == want.passes
    assert (head["tokens"], head["types"], head["maxn"], head["minn"]) == (want.tokens, want.types, want.maxn, want.minn)


# ---- constrained training, sharded (SURVEY 8f-2 x 8e): replicated constraint set, local counting, one all-reduce
class NumpyConstrainedEngine:
    """Test double with the interface of multigpu.CudaConstrainedEngine, backed by the oracle on this rank's shard."""

    def __init__(self, body, stage1_blob, load_kw, train_kw):
        import oracle

        self.body, self.train_kw = body, train_kw
        self.cm = oracle.load_model(stage1_blob, **load_kw)
        self.flat = self.cm.flat()  # canonical order = the pattern numbering all ranks share

    def count(self):
        import oracle

        kw = dict(self.train_kw, mintokens=1)
        local = oracle.train_constrained(self.body, self.cm, inplace=False, **kw)
        d = local.as_dict()
        counts = np.array([d.get(self.flat.key(i), 0) for i in range(len(self.flat))], dtype=np.int32)
        return torch.from_numpy(counts if len(counts) else np.zeros(1, dtype=np.int32)), local.tokens - self.flat.tokens

    def finish(self, counts, tokens, inplace):
        t = self.train_kw["mintokens"]
        c = counts.numpy()
        return {self.flat.key(i): int(c[i]) for i in range(len(self.flat)) if c[i] >= t}, tokens


def _constrained_worker(rank, world, port, bodies, stage1, load_kw, train_kw, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colibri_core_b200.multigpu as mg

    eng = NumpyConstrainedEngine(bodies[rank], stage1, load_kw, train_kw)
    model, tokens = mg.train_constrained_distributed(eng, dist, torch, inplace=False)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((model, tokens), gathered, dst=0)
    if rank == 0:
        results.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mintokens,maxlength,seed", [(2, 2, 5, 3), (3, 3, 4, 5), (2, 1, 3, 7)])
def test_distributed_constrained_orchestration_matches_oracle(world, mintokens, maxlength, seed):
    import oracle

    per = 4000
    bodies = [oracle.synth_corpus(per, vocab=60, seed=seed, mean_sentence=9, phrase_permille=300, nphrases=20, first_token=r * per).tobytes() for r in range(world)]
    stage1 = oracle.train_to_modelfile(oracle.synth_corpus(6000, vocab=60, seed=seed + 1, mean_sentence=9, phrase_permille=300, nphrases=20).tobytes(), mintokens=2, maxlength=maxlength)
    load_kw = dict(mintokens=mintokens, maxlength=maxlength)
    train_kw = dict(mintokens=mintokens, maxlength=maxlength, streamed=1)
    want = oracle.train_constrained(b"".join(bodies), oracle.load_model(stage1, **load_kw), inplace=False, **train_kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + seed) % 90
    procs = [ctx.Process(target=_constrained_worker, args=(r, world, port, bodies, stage1, load_kw, train_kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for model, tokens in gathered:  # every rank holds the same model
        assert model == want.as_dict()
        assert tokens == per * world


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_strong_scaling_cuts_fall_on_sentence_boundaries(seed, world):
    """The strong-scaling parity step cuts ONE corpus into `world` shards (multigpu.cut_at_sentences): every cut must follow a real delimiter
    (a 0x00 that is not the last byte of a multi-byte class), the shards must tile the corpus, and the sentences of the shards are the
    sentences of the corpus -- so the n-grams of the shards are exactly the n-grams of the whole (windows never cross a sentence)."""
    import colibri_core_b200.multigpu as mg
    import oracle

    rng = np.random.default_rng(seed)
    # one-, two- and three-byte classes, empty sentences among them
    sentences = [[int(rng.choice([6, 7, 127, 128, 256, 300, 16384, 16512, 40000])) for _ in range(int(rng.integers(0, 9)))] for _ in range(int(rng.integers(1, 60)))]
    body = np.frombuffer(bytes(oracle.encode_corpus(sentences)), dtype=np.uint8)
    cuts = mg.cut_at_sentences(body, world)
    assert cuts[0] == 0 and cuts[-1] == len(body) and len(cuts) == world + 1
    assert all(a <= b for a, b in zip(cuts, cuts[1:]))
    for c in cuts[1:-1]:
        if 0 < c < len(body):
            assert body[c - 1] == 0 and (c < 2 or body[c - 2] < 128), (c, body[max(0, c - 3):c + 1])
    whole = oracle.corpus_sentences(body)
    parts = [s for a, b in zip(cuts, cuts[1:]) for s in oracle.corpus_sentences(body[a:b])]
    assert [s for s in parts if s] == [s for s in whole if s]


class _FailingEngine(NumpyShardEngine):
    """Rank 1's second level fails the way a CUDA phase does (an exception out of the engine) while rank 0 is already inside the next collective."""

    def level_split_write(self, nsend):
        if self.rank == 1 and self.level >= 2:  # (level 3 of this corpus always exists)
            raise RuntimeError("receive slot overflow (injected)")
        return super().level_split_write(nsend)


def _failing_worker(rank, world, port, bodies):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colibri_core_b200.multigpu as mg

    mg.train_distributed(_FailingEngine(bodies[rank], 2, rank, world), dist, torch, 2, 4, False)


def test_a_failing_rank_leaves_the_process_instead_of_hanging_the_job():
    """multigpu.train_distributed: the rank whose phase raises prints the reason and exits the process with status 13, so that the launcher
    (torchrun, or this test) can take the other ranks down; without that they would wait in their collective forever."""
    import oracle

    bodies = [oracle.synth_corpus(3000, vocab=40, seed=5, mean_sentence=9, first_token=r * 3000).tobytes() for r in range(2)]
    ctx = mp.get_context("spawn")
    port = 29600 + (os.getpid() + 77) % 300
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, bodies)) for r in range(2)]
    for p in procs:
        p.start()
    procs[1].join(timeout=120)
    assert procs[1].exitcode == 13
    procs[0].join(timeout=5)  # rank 0 is stuck in (or failed out of) its collective: what a launcher does next
    if procs[0].is_alive():
        procs[0].terminate()
        procs[0].join(timeout=30)
    assert procs[0].exitcode != 0
