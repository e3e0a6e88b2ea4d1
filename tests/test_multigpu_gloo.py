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
        return found, kept, occ

    def _owner(self, a, b):
        return (a * 1000003 + b * 7919) % self.world

    def level_count(self, n):
        self.local, self.cur = {}, np.zeros(len(self.tok), dtype=np.int64)
        for p in range(len(self.tok) - 1):
            a, b = int(self.prev[p]), int(self.prev[p + 1])
            if a and b:
                e = self.local.setdefault((a, b), [0, p, len(self.local)])
                e[0] += 1
                self.cur[p] = e[2] + 1
        self.order = sorted(self.local.items(), key=lambda kv: (self._owner(*kv[0]), kv[1][2]))
        dest = [0] * self.world
        for (a, b), _ in self.order:
            dest[self._owner(a, b)] += 1
        return dest, int((self.cur != 0).sum()), len(self.local)

    def level_pack(self, nsend):
        rec = np.zeros((max(nsend, 1), 4), dtype=np.uint32)
        for j, ((a, b), (cnt, _pos, slot)) in enumerate(self.order):
            rec[j] = (b, a, cnt, slot)  # key = a << 32 | b, little-endian words
        return torch.from_numpy(rec.view(np.int32).reshape(-1))

    def level_merge(self, recv, nrecv):
        rec = recv.numpy().view(np.uint32)[: nrecv * 4].reshape(-1, 4)
        owner = {}
        slots = []
        for i in range(nrecv):
            key = (int(rec[i, 1]), int(rec[i, 0]))
            e = owner.setdefault(key, [0, i, len(owner)])
            e[0] += int(rec[i, 2])
            slots.append(e)
        reply = np.zeros((max(nrecv, 1), 2), dtype=np.uint32)
        for i, e in enumerate(slots):
            if e[0] >= self.t:
                reply[i] = (e[2] * self.world + self.rank + 1, e[0] if e[1] == i else 0)
        kept = [e for e in owner.values() if e[0] >= self.t]
        return torch.from_numpy(reply.view(np.int32).reshape(-1)), (len(owner), len(kept), sum(e[0] for e in kept))

    def level_finish(self, reply_back):
        n = self.level + 1
        rep = reply_back.numpy().view(np.uint32)[: len(self.order) * 2].reshape(-1, 2)
        gid = np.zeros(len(self.local) + 1, dtype=np.int64)
        for j, (_key, (_cnt, pos, slot)) in enumerate(self.order):
            gid[slot] = int(rep[j, 0])
            if rep[j, 1]:
                self.segs.append((n, pos, int(rep[j, 1])))
        self.cur = np.where(self.cur != 0, gid[np.maximum(self.cur - 1, 0)], 0)
        self.prev, self.level = self.cur, n
        return int((self.cur != 0).sum())

    def finish(self, passes, types, maxn, minn):
        import oracle

        out = {}
        for n, pos, cnt in self.segs:
            toks = [pos] if n == 1 else self.tok[pos:pos + n].tolist()
            out[b"".join(oracle.inttobytes(int(c)) for c in toks)] = cnt
        return out


def _worker(rank, world, port, bodies, mintokens, maxlength, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colibri_core_b200.multigpu as mg

    eng = NumpyShardEngine(bodies[rank], mintokens, rank, world)
    share, passes, head = mg.train_distributed(eng, dist, torch, mintokens, maxlength)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((share, passes, head), gathered, dst=0)
    if rank == 0:
        results.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mintokens,maxlength,seed", [(2, 2, 5, 3), (3, 2, 4, 5), (2, 3, 6, 7), (2, 1, 3, 9)])
def test_distributed_orchestration_matches_oracle(world, mintokens, maxlength, seed):
    import oracle

    per = 4000
    bodies = [oracle.synth_corpus(per, vocab=60, seed=seed, mean_sentence=9, phrase_permille=300, nphrases=20, first_token=r * per).tobytes() for r in range(world)]
    want = oracle.train(b"".join(bodies), mintokens=mintokens, maxlength=maxlength)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + seed) % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, bodies, mintokens, maxlength, q)) for r in range(world)]
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
