"""Multi-GPU PatternModel::train: one process per GPU, corpus sharded at sentence boundaries, model partitioned by hash.

The library (csrc/shard.cu) exposes each rank's work as phases with plain device pointers; this module is the plumbing
between them: torch.distributed collectives over NCCL/NVLink (gloo on CPU in the tests, with a stand-in engine).

Per level n >= 2, "ship the windows to their owners":
    level_split (count, write) -> all_to_all(8-byte keys) -> level_owner (filter + count + prune on the received stream)
    -> all_to_all back (4-byte global ids, same routes) + all_to_all(8-byte survivor records) -> level_finish
Level 1 is an all-reduce of the class histograms.  Every rank ends with its own share of the surviving patterns
(the rank whose window claimed a pattern's slot at the owner exports it, with the global count).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np


class CudaShardEngine:
    """Thin wrapper over the colibri_b200_shard_* phases for one rank; buffers are torch CUDA tensors."""

    def __init__(self, corpus, options, rank, world, device):
        import torch

        from . import _check, library

        self.torch, self._check, self.lib = torch, _check, library()
        self.device = torch.device("cuda", device)
        self.rank, self.world = rank, world
        self.corpus = corpus
        self._h = C.c_void_p()
        _check(self.lib.colibri_b200_shard_begin(corpus._h, C.byref(options._c), rank, world, C.byref(self._h)))

    def new_buffer(self, nwords):
        return self.torch.empty(max(int(nwords), 1), dtype=self.torch.int32, device=self.device)

    def info(self):
        out = (C.c_uint64 * 4)()
        self._check(self.lib.colibri_b200_shard_info(self._h, out))
        return {"tokens": int(out[0]), "maxclass": int(out[1]), "positions": int(out[2]), "launches": int(out[3])}

    def device_ms(self):
        return float(self.lib.colibri_b200_shard_device_ms(self._h))

    def phase_ms(self):
        out = (C.c_double * 8)()
        self._check(self.lib.colibri_b200_shard_phase_ms(self._h, out))
        return dict(zip(["tokenise", "unigrams", "split_count", "split_write", "owner", "finish", "export"], [float(x) for x in out]))

    def unigram_counts(self, nclasses):
        buf = self.new_buffer(nclasses)
        self._check(self.lib.colibri_b200_shard_unigram_counts(self._h, nclasses, buf.data_ptr()))
        return buf

    def unigram_finish(self, global_counts, global_tokens):
        st = (C.c_uint64 * 3)()
        self._check(self.lib.colibri_b200_shard_unigram_finish(self._h, global_counts.data_ptr(), global_tokens, st))
        return tuple(int(x) for x in st)

    def level_split_count(self, n):
        counts = (C.c_uint64 * self.world)()
        w = C.c_uint64()
        self._check(self.lib.colibri_b200_shard_level_split_count(self._h, n, counts, C.byref(w)))
        return [int(x) for x in counts], int(w.value)

    def level_split_write(self, nsend):
        buf = self.new_buffer(nsend * 2)
        self._check(self.lib.colibri_b200_shard_level_split_write(self._h, buf.data_ptr()))
        return buf

    def level_owner(self, recv, recv_counts):
        nrecv = sum(recv_counts)
        reply = self.new_buffer(nrecv)
        rc = (C.c_uint64 * self.world)(*recv_counts)
        st = (C.c_uint64 * 3)()
        sc = (C.c_uint64 * self.world)()
        self._check(self.lib.colibri_b200_shard_level_owner(self._h, recv.data_ptr(), rc, reply.data_ptr(), st, sc))
        return reply, tuple(int(x) for x in st), [int(x) for x in sc]

    def level_owner_survivors(self, nsurv):
        buf = self.new_buffer(nsurv * 2)
        self._check(self.lib.colibri_b200_shard_level_owner_survivors(self._h, buf.data_ptr()))
        return buf

    def level_finish(self, reply_back, surv, surv_counts):
        v = C.c_uint64()
        sc = (C.c_uint64 * self.world)(*surv_counts)
        self._check(self.lib.colibri_b200_shard_level_finish(self._h, reply_back.data_ptr(), surv.data_ptr(), sc, C.byref(v)))
        return int(v.value)

    def finish(self, passes, types, maxn, minn):
        from . import Model

        flat = (C.c_uint64 * (4 * max(len(passes), 1)))()
        for i, p in enumerate(passes):
            for j in range(4):
                flat[4 * i + j] = int(p[j])
        h = C.c_void_p()
        self._check(self.lib.colibri_b200_shard_finish(self._h, flat, len(passes), types, maxn, minn, C.byref(h)))
        return Model(h)

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    # ---- exhaustive skipgrams of the level just finished
    def skip_split_count(self):
        counts = (C.c_uint64 * self.world)()
        n = C.c_uint64()
        self._check(self.lib.colibri_b200_shard_skip_split_count(self._h, counts, C.byref(n)))
        return [int(x) for x in counts], int(n.value)

    def skip_split_write(self, nsend):
        buf = self.new_buffer(nsend * 4)
        self._check(self.lib.colibri_b200_shard_skip_split_write(self._h, buf.data_ptr()))
        return buf

    def skip_owner(self, recv, recv_counts):
        rc = (C.c_uint64 * self.world)(*recv_counts)
        st = (C.c_uint64 * 2)()
        sc = (C.c_uint64 * self.world)()
        self._check(self.lib.colibri_b200_shard_skip_owner(self._h, recv.data_ptr(), rc, st, sc))
        return tuple(int(x) for x in st), [int(x) for x in sc]

    def skip_owner_survivors(self, nsurv):
        buf = self.new_buffer(nsurv * 4)
        self._check(self.lib.colibri_b200_shard_skip_owner_survivors(self._h, buf.data_ptr()))
        return buf

    def skip_finish(self, surv, surv_counts):
        sc = (C.c_uint64 * self.world)(*surv_counts)
        self._check(self.lib.colibri_b200_shard_skip_finish(self._h, surv.data_ptr(), sc))

    def enable_dense(self, dim):
        """Level 2: pairs of classes below `dim` are counted in a square on every rank and summed by ONE all-reduce instead of being shipped."""
        self.dense_t = self.torch.zeros(dim * dim, dtype=self.torch.int32, device=self.device)
        self._check(self.lib.colibri_b200_shard_set_dense(self._h, self.dense_t.data_ptr(), dim))
        return self.dense_t

    # ---- NVLink peer-store mode
    def use_peers(self, peers: "PeerBuffers"):
        self._check(self.lib.colibri_b200_shard_set_stream(self._h, C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)))
        arr = [(C.c_uint64 * self.world)(*p) for p in peers.ptrs]
        self._check(self.lib.colibri_b200_shard_set_peers(self._h, arr[0], arr[1], arr[2], arr[3], peers.slot_cap, peers.surv_cap))
        self.peers = peers

    def p2p_split(self, n):
        w = C.c_uint64()
        self._check(self.lib.colibri_b200_shard_p2p_split(self._h, n, C.byref(w)))
        return int(w.value)

    def p2p_owner(self):
        st = (C.c_uint64 * 3)()
        self._check(self.lib.colibri_b200_shard_p2p_owner(self._h, st))
        return tuple(int(x) for x in st)

    def p2p_finish(self):
        st = (C.c_uint64 * 3)()
        v = C.c_uint64()
        self._check(self.lib.colibri_b200_shard_p2p_finish(self._h, st, C.byref(v)))
        return tuple(int(x) for x in st), int(v.value)

    def close(self):
        if self._h:
            self.lib.colibri_b200_shard_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffers:
    """Symmetric receive buffers of one rank (torch.distributed._symmetric_memory), rendezvoused once and reused by every
    step: the split / reply kernels of the other ranks store into them over NVLink.  Sized for `positions` per rank."""

    _cache = {}

    def __init__(self, dist, torch, world, positions, device):
        import torch.distributed._symmetric_memory as symm

        self.world = world
        self.slot_cap = int(1.3 * positions / world) + 65536
        self.surv_cap = self.slot_cap // 2 + 4096
        dev = torch.device("cuda", device)
        self.keys = symm.empty(world * self.slot_cap * 2, dtype=torch.int32, device=dev)
        self.reply = symm.empty(world * self.slot_cap, dtype=torch.int32, device=dev)
        self.surv = symm.empty(world * self.surv_cap * 2, dtype=torch.int32, device=dev)
        self.hdr = symm.empty(6 * world * 2 + 64, dtype=torch.int32, device=dev)
        self.hdr.zero_()
        group = dist.group.WORLD
        self.handles = [symm.rendezvous(t, group) for t in (self.keys, self.reply, self.surv, self.hdr)]
        self.ptrs = [[int(p) for p in h.buffer_ptrs] for h in self.handles]
        torch.cuda.synchronize()
        dist.barrier()

    @classmethod
    def get(cls, dist, torch, world, positions, device):
        key = (world, device, int(positions))
        if key not in cls._cache:
            cls._cache[key] = cls(dist, torch, world, positions, device)
        return cls._cache[key]

    def barrier(self):
        self.handles[3].barrier(channel=0)  # device-side barrier through the signal pads, enqueued on the current stream


def _exchange(dist, torch, engine, send, send_counts, width):
    """all-to-all of variable-length record groups; returns (recv buffer, recv_counts)."""
    world = len(send_counts)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = engine.new_buffer(sum(recv_counts) * width)
    nsend, nrecv = sum(send_counts) * width, sum(recv_counts) * width
    dist.all_to_all_single(recv[:nrecv], send[:nsend], output_split_sizes=[c * width for c in recv_counts], input_split_sizes=[c * width for c in send_counts])
    engine.sync()
    return recv, recv_counts


class _Stopwatch:
    """Wall-clock split of one distributed train (enabled with COLIBRI_B200_TRACE=1): where the non-kernel time goes."""

    def __init__(self, engine):
        self.on = bool(os.environ.get("COLIBRI_B200_TRACE"))
        self.engine, self.acc, self.t = engine, {}, time.perf_counter()

    def lap(self, name):
        if not self.on:
            return
        self.engine.sync()
        now = time.perf_counter()
        self.acc[name] = self.acc.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def _skipgram_level(engine, dist, torch, dev):
    """Exhaustive skipgrams of the level the engine just finished: keys to their owners, survivors back.  Returns (found, kept) globally."""
    send_counts, nsend = engine.skip_split_count()
    send = engine.skip_split_write(nsend)
    recv, recv_counts = _exchange(dist, torch, engine, send, send_counts, 4)
    (f, k), surv_counts = engine.skip_owner(recv, recv_counts)
    surv = engine.skip_owner_survivors(sum(surv_counts))
    surv_recv, surv_recv_counts = _exchange(dist, torch, engine, surv, surv_counts, 4)
    st = torch.tensor([f, k], dtype=torch.int64, device=dev)
    dist.all_reduce(st, op=dist.ReduceOp.SUM)
    engine.sync()
    engine.skip_finish(surv_recv, surv_recv_counts)
    return int(st[0].item()), int(st[1].item())


def train_distributed(engine, dist, torch, mintokens=2, maxlength=5, skipgrams=False):
    """Drive one rank through all levels.  Returns (local model share, global passes, global header dict).

    A rank whose phase fails (a receive slot that overflows, a table that cannot grow, a malformed shard) cannot tell the others: they are inside a
    device-side barrier or a collective by then.  So the failing rank reports and leaves the PROCESS; torchrun (and any launcher that watches its
    workers) then takes the other ranks down instead of letting them wait forever.  COLIBRI_B200_NO_ABORT=1 re-raises instead (single-rank tests)."""
    try:
        return _train_distributed(engine, dist, torch, mintokens, maxlength, skipgrams)
    except Exception as e:
        if dist.get_world_size() > 1 and not os.environ.get("COLIBRI_B200_NO_ABORT"):
            print("colibri_b200: rank %d failed, aborting the job: %r" % (dist.get_rank(), e), file=sys.stderr, flush=True)
            os._exit(13)
        raise


def _train_distributed(engine, dist, torch, mintokens=2, maxlength=5, skipgrams=False):
    world = dist.get_world_size()
    sw = _Stopwatch(engine)
    info = engine.info()
    dev = engine.new_buffer(1).device
    head = torch.tensor([info["tokens"], 0], dtype=torch.int64, device=dev)
    dist.all_reduce(head[:1], op=dist.ReduceOp.SUM)
    mx = torch.tensor([info["maxclass"]], dtype=torch.int64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    global_tokens, nclasses = int(head[0].item()), int(mx.item()) + 1

    sw.lap("head_allreduce")
    counts = engine.unigram_counts(nclasses)
    sw.lap("unigram_counts")
    dist.all_reduce(counts[:nclasses], op=dist.ReduceOp.SUM)  # u32 counts carried as int32 bit patterns: exact below 2^31 occurrences per class
    engine.sync()
    sw.lap("unigram_allreduce")
    found, kept, kept_occ = engine.unigram_finish(counts, global_tokens)
    sw.lap("unigram_finish")
    passes, maxn, minn, types = [], 0, 999, found
    if found:
        passes.append((1, found, 0, found - kept))
        maxn = minn = 1
    prev_kept = kept
    n = 2
    peers = getattr(engine, "peers", None)
    # dense pairs of level 2: decided from global quantities only, so that every rank decides alike (the ids of level 2 depend on it)
    dense_t = None
    dense_dim = min(int(os.environ.get("COLIBRI_B200_DENSE", "3072")), nclasses, 16384)
    if hasattr(engine, "enable_dense") and dense_dim >= 2 and global_tokens // world >= int(os.environ.get("COLIBRI_B200_DENSE_MIN", str(1 << 25))):
        dense_t = engine.enable_dense(dense_dim)
    while peers is not None and found and n <= maxlength and prev_kept > 0:
        # NVLink peer-store mode: keys and replies are stored into the peers' symmetric buffers by the kernels themselves
        engine.p2p_split(n)
        if n == 2 and dense_t is not None:
            dist.all_reduce(dense_t)  # same stream as the phases: summed before the owner phase reads it
        sw.lap("p2p_split")
        peers.barrier()
        sw.lap("barrier")
        engine.p2p_owner()
        sw.lap("p2p_owner")
        peers.barrier()
        sw.lap("barrier")
        (gf, gk, _gocc), _valid = engine.p2p_finish()
        sw.lap("p2p_finish")
        if gf == 0:
            break
        sf = sk = 0
        if skipgrams and n >= 3:
            sf, sk = _skipgram_level(engine, dist, torch, dev)
            sw.lap("skipgrams")
        passes.append((n, gf, sf, (gf - gk) + (sf - sk)))
        maxn, minn = max(maxn, n), min(minn, n)
        prev_kept = gk
        n += 1
    while peers is None and found and n <= maxlength and prev_kept > 0:
        send_counts, nsend = engine.level_split_count(n)
        if n == 2 and dense_t is not None:
            dist.all_reduce(dense_t)  # (the engine.sync() of the key exchange below covers it)
        sw.lap("split_count")
        send = engine.level_split_write(nsend)
        sw.lap("split_write")
        recv, recv_counts = _exchange(dist, torch, engine, send, send_counts, 2)
        sw.lap("a2a_keys")
        reply, (f, k, occ), surv_counts = engine.level_owner(recv, recv_counts)
        surv = engine.level_owner_survivors(sum(surv_counts))
        sw.lap("owner")
        # replies travel back along the same routes: what I received from rank r goes back to r
        nrep = sum(recv_counts)
        back = engine.new_buffer(nsend)
        dist.all_to_all_single(back[:nsend], reply[:nrep], output_split_sizes=send_counts, input_split_sizes=recv_counts)
        surv_recv, surv_recv_counts = _exchange(dist, torch, engine, surv, surv_counts, 2)
        st = torch.tensor([f, k, occ], dtype=torch.int64, device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
        engine.sync()
        sw.lap("a2a_replies")
        engine.level_finish(back, surv_recv, surv_recv_counts)
        sw.lap("level_finish")
        gf, gk, _gocc = (int(x) for x in st.tolist())
        if gf == 0:
            break  # "None found" (reference include/patternmodel.h:1189-1194)
        sf = sk = 0
        if skipgrams and n >= 3:
            sf, sk = _skipgram_level(engine, dist, torch, dev)
            sw.lap("skipgrams")
        passes.append((n, gf, sf, (gf - gk) + (sf - sk)))
        maxn, minn = max(maxn, n), min(minn, n)
        prev_kept = gk
        n += 1
    if mintokens == 1 and passes:  # the reference reports one pass when every length is extracted in a single scan
        passes = [(1, sum(p[1] for p in passes), sum(p[2] for p in passes), sum(p[3] for p in passes))]
    model = engine.finish(passes, types, maxn, minn)
    sw.lap("export")
    if sw.on:
        print("TRACE rank%d ms:" % dist.get_rank(), {k: round(v, 2) for k, v in sw.acc.items()}, "device phases:", {k: round(v, 2) for k, v in engine.phase_ms().items()}, flush=True)
    return model, passes, {"tokens": global_tokens, "types": types, "maxn": maxn, "minn": minn}


# --------------------------------------------------------------------------------------------- constrained training, sharded
class CudaConstrainedEngine:
    """Per-rank side of a sharded constrained run (PatternModel::train with constrainbymodel, SURVEY 8f-2 x 8e): the constraint set is
    replicated, the corpus is cut at sentence boundaries, one shard per rank.  Counting needs no communication at all."""

    def __init__(self, shard, constrain, options, local, torch):
        self.shard, self.constrain, self.options, self.local, self.torch = shard, constrain, options, local, torch

    def count(self):
        counts = self.torch.zeros(max(len(self.constrain), 1), dtype=self.torch.int32, device="cuda:%d" % self.local)
        self.torch.cuda.synchronize(self.local)
        tokens, self.launches = _cb().constrained_count(self.shard, self.constrain, self.options, counts.data_ptr())
        return counts, tokens

    def finish(self, counts, tokens, inplace):
        self.torch.cuda.synchronize(self.local)
        return _cb().constrained_finish(self.constrain, self.options, counts.data_ptr(), tokens, inplace)


def _cb():
    import colibri_core_b200 as cb

    return cb


def train_constrained_distributed(engine, dist, torch, inplace=False):
    """count (local) -> all-reduce SUM of the per-pattern counters and of the token counts -> finish (the same model on every rank).
    The counters are uint32 carried in int32 tensors: two's-complement sums wrap the same way."""
    counts, tokens = engine.count()
    tok = torch.tensor([tokens], dtype=torch.int64, device=counts.device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts)
        dist.all_reduce(tok)
    return engine.finish(counts, int(tok.item()), inplace)


def cut_at_sentences(body: np.ndarray, world: int):
    """Byte offsets that cut a class-encoded corpus into `world` shards of about equal size at sentence boundaries (a delimiter is a 0x00 byte
    that is not the last byte of a multi-byte class: the byte before it is below 128)."""
    cuts = [0]
    for r in range(1, world):
        j = max(len(body) * r // world, cuts[-1])
        while j < len(body) and not (body[j] == 0 and (j == 0 or body[j - 1] < 128)):
            j += 1
        cuts.append(min(j + 1, len(body)))
    cuts.append(len(body))
    return cuts


def strong_scaling_parity(a, dist, torch, cb, rank, world, local, opts, peers):
    """One untimed STRONG-scaling step: the 1-GPU benchmark corpus (the one tests/golden/golden_bench.json pins to the unmodified reference) cut
    into `world` shards at sentence boundaries and trained through the same sharded path; the order-independent checksums of the ranks' shares are
    combined (sum, xor, occurrences, patterns) and compared with the reference model's -- the same model from 1, 2, 4 or 8 GPUs."""
    if (int(a.tokens), a.vocab, a.seed, a.maxlength, a.mintokens, int(a.skipgrams)) != (100000000, 100000, 1, 5, 2, 0):
        return {"checked": False, "why": "the committed fixture pins the default workload only"}
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_bench.json")) as f:
            g = json.load(f)["zipf100m"]
    except Exception as e:
        return {"checked": False, "why": "no fixture: %r" % (e,)}
    full = cb.Corpus.synthetic(int(a.tokens), vocab=a.vocab, seed=a.seed, device=local)
    body = full.download()
    full.close()
    cuts = cut_at_sentences(body, world)
    shard = np.ascontiguousarray(body[cuts[rank]:cuts[rank + 1]])
    c = cb.Corpus.from_host_pointer(shard.ctypes.data, shard.size, device=local)
    eng = CudaShardEngine(c, opts, rank, world, local)
    if peers is not None:
        eng.use_peers(peers)
    model, passes, head = train_distributed(eng, dist, torch, a.mintokens, a.maxlength, False)
    cs = model.checksum()
    n_local = len(model)
    model.close()
    eng.close()
    c.close()

    def i64(x):  # u64 bit pattern as a signed value: two's-complement sums wrap the same way
        return x - (1 << 64) if x >= (1 << 63) else x

    tot = torch.tensor([i64(cs["sum"]), cs["occurrences"], cs["patterns"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    xors = [None] * world
    dist.all_gather_object(xors, cs["xor"])
    x = 0
    for v in xors:
        x ^= v
    got = {"sum": int(tot[0].item()) & ((1 << 64) - 1), "xor": x, "occurrences": int(tot[1].item()), "patterns": int(tot[2].item())}
    want = g["checksum"]
    return {"checked": True, "what": "strong scaling: the 100 M-token benchmark corpus cut into %d shards at sentence boundaries, trained on %d GPUs" % (world, world),
            "fixture": "tests/golden/golden_bench.json: zipf100m (unmodified reference, %s)" % g["cli"], "shard_bytes": [cuts[i + 1] - cuts[i] for i in range(world)],
            "checksum": got, "checksum_ok": all(got[k] == want[k] for k in ("sum", "xor", "occurrences", "patterns")),
            "passes_ok": [(p[1], p[3]) for p in passes] == [(p[0], p[2]) for p in g["passes_found_skip_pruned_kept"]],
            "header_ok": (head["tokens"], head["types"]) == (g["tokens"], g["types"]), "patterns_on_rank0": n_local}


def bench(a, dist, rank, world, local, metric, unit, workload, ClockSampler, measured_peaks):
    """bench.py --gpus N under torchrun: weak scaling, every rank trains its own `--tokens` shard of one global stream."""
    import torch

    import colibri_core_b200 as cb

    ntok = int(a.tokens)
    opts = cb.PatternModelOptions(MINTOKENS=a.mintokens, MAXLENGTH=a.maxlength, DOSKIPGRAMS_EXHAUSTIVE=a.skipgrams, streamed=0 if a.skipgrams else 1, QUIET=1, device=local)
    strong = getattr(a, "scaling", "weak") == "strong"
    if strong:
        # one corpus of --tokens for the whole job, cut at sentence boundaries: every rank trains 1/N of it (the model is the 1-GPU model)
        full = cb.Corpus.synthetic(ntok, vocab=a.vocab, seed=a.seed, device=local)
        body = full.download()
        full.close()
        cuts = cut_at_sentences(body, world)
        shard_bytes = np.ascontiguousarray(body[cuts[rank]:cuts[rank + 1]])
        corpus = cb.Corpus.from_host_pointer(shard_bytes.ctypes.data, shard_bytes.size, device=local)
        ntok = max(ntok // world, 1)
    else:
        corpus = cb.Corpus.synthetic(ntok, vocab=a.vocab, seed=a.seed, device=local, first_token=rank * ntok)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    peers = None
    if not os.environ.get("COLIBRI_B200_NO_P2P"):
        try:
            peers = PeerBuffers.get(dist, torch, world, int(ntok * 1.06) + 1024, local)
        except Exception as e:  # symmetric memory unavailable: the NCCL all-to-all path is used instead (still all on the GPUs)
            if rank == 0:
                print("note: symmetric memory unavailable (%r); using NCCL all-to-all" % (e,), file=sys.stderr, flush=True)

    def step():
        eng = CudaShardEngine(corpus, opts, rank, world, local)
        if peers is not None:
            eng.use_peers(peers)
        model, passes, head = train_distributed(eng, dist, torch, a.mintokens, a.maxlength, bool(a.skipgrams))
        out = (len(model), head, passes, eng.device_ms(), eng.info()["launches"], eng.phase_ms())
        model.close()
        eng.close()
        return out

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    launches, dev_ms, last = 0, 0.0, None
    for _ in range(a.steps):
        last = step()
        launches += last[4]
        dev_ms += last[3]
    e1.record()
    barrier()
    elapsed = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
    t = torch.tensor([elapsed, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the slowest rank defines the step
    npat = torch.tensor([last[0], launches], dtype=torch.int64, device="cuda")
    dist.all_reduce(npat, op=dist.ReduceOp.SUM)
    # ---- end to end: pinned host shard -> H2D -> distributed train -> this rank's share of the model -> D2H (pinned)
    host = torch.empty(corpus.nbytes, dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = corpus.download()
    cap_pat = int(last[0] * 1.3) + 1024
    out_keys = torch.empty(cap_pat * 16, dtype=torch.uint8, pin_memory=True)
    out_len = torch.empty(cap_pat + 1, dtype=torch.int16, pin_memory=True)
    out_cnt = torch.empty(cap_pat, dtype=torch.int32, pin_memory=True)

    def e2e_step():
        c = cb.Corpus.from_host_pointer(host.data_ptr(), host.numel(), device=local)
        eng = CudaShardEngine(c, opts, rank, world, local)
        if peers is not None:
            eng.use_peers(peers)
        model, _, _ = train_distributed(eng, dist, torch, a.mintokens, a.maxlength, bool(a.skipgrams))
        n, kb, _ = model.export_sizes()
        if n > cap_pat or kb > out_keys.numel():
            raise RuntimeError("e2e export buffers too small")
        model.export_compact_into(out_keys.data_ptr(), out_len.data_ptr(), out_cnt.data_ptr())
        model.close()
        eng.close()
        c.close()
        return kb + 2 * n + 4 * n

    e2e_step()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(a.steps):
        d2h = e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    xfer = torch.tensor([corpus.nbytes, d2h], dtype=torch.int64, device="cuda")
    dist.all_reduce(xfer, op=dist.ReduceOp.SUM)
    parity = strong_scaling_parity(a, dist, torch, cb, rank, world, local, opts, peers)
    if rank == 0:
        clocks = sampler.stop()
        elapsed = float(t[0].item())
        tokens = last[1]["tokens"]
        peak, peak_src = measured_peaks()
        line = {
            "metric": metric, "value": tokens * a.steps / elapsed, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * elapsed / a.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
            "config": {"workload": workload + (" cut into %d shards at sentence boundaries" % world if strong else " PER GPU (shards of one global stream)"), "global_tokens": tokens, "patterns": int(npat[0].item()),
                       "parallelism": "corpus sharded by sentence x%d, model hash-partitioned; windows shipped to their owners by %s" % (
                           world, "NVLink peer stores from the split/reply kernels (symmetric memory)" if peers is not None else "NCCL all-to-all"),
                       "l2": "inputs exceed L2", "timing": "max over ranks of max(CUDA events, wall clock) around K steps"},
            "device_ms_per_step_max_rank": 1e3 * float(t[1].item()) / a.steps, "passes": last[2], "rank0_phase_ms_last_step": last[5],
            "roofline": {"bound": "hbm", "kernel": "count_ngrams_kernel", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "peak_source": peak_src, "traffic": None,
                         "note": "per-kernel roofline is reported by the 1-GPU run; the N-GPU line reports whole-job throughput",
                         "hbm_read_roofline_frac": (a.maxlength * corpus.nbytes / (elapsed / a.steps) / 1e9) / peak},
            "clocks": clocks, "gpu_launches": int(npat[1].item()), "parity": parity,
            "e2e": {"value": tokens * a.steps / float(e2e_t.item()), "unit": unit, "h2d_bytes_per_step": int(xfer[0].item()), "d2h_bytes_per_step": int(xfer[1].item()),
                    "ms_per_step": 1e3 * float(e2e_t.item()) / a.steps, "api": "per rank: colibri_b200_corpus_stage(pinned host shard) + shard phases + NCCL + colibri_b200_model_export_compact (pinned)"},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
