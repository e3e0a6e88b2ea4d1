"""Worker for tests/test_multigpu_gpu.py (launched by torch.distributed.run): every rank trains its shard through the
CUDA shard phases + NCCL; rank 0 gathers the shares and compares with the oracle on the concatenated corpus."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import colibri_core_b200 as cb  # noqa: E402
import colibri_core_b200.multigpu as mg  # noqa: E402


def main():
    per, vocab, seed, maxlength, mintokens = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(vocab=vocab, seed=seed, mean_sentence=15, phrase_permille=150, nphrases=500)
    corpus = cb.Corpus.synthetic(per, device=local, first_token=rank * per, **kw)
    if len(sys.argv) > 6 and sys.argv[6] == "constrained":
        return constrained(per, kw, maxlength, mintokens, rank, world, local, corpus)
    skip = len(sys.argv) > 7 and sys.argv[7] == "skipgrams"
    opts = cb.PatternModelOptions(MINTOKENS=mintokens, MAXLENGTH=maxlength, DOSKIPGRAMS_EXHAUSTIVE=int(skip), streamed=0 if skip else 1, QUIET=1, device=local)
    eng = mg.CudaShardEngine(corpus, opts, rank, world, local)
    if len(sys.argv) > 6 and sys.argv[6] == "p2p":  # NVLink peer-store mode (symmetric memory)
        eng.use_peers(mg.PeerBuffers.get(dist, torch, world, per * 2, local))
    model, passes, head = mg.train_distributed(eng, dist, torch, mintokens, maxlength, skip)
    keys, off, counts, _ = model.export()
    share = (keys.tobytes(), off.tolist(), counts.tolist())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((share, passes, head), gathered, dst=0)
    ok = True
    if rank == 0:
        import oracle

        merged = {}
        for (kb, of, cn), p, h in gathered:
            for i in range(len(cn)):
                k = kb[of[i]:of[i + 1]]
                assert k not in merged, "pattern exported twice"
                merged[k] = cn[i]
            assert p == passes and h == head
        body = b"".join(oracle.synth_corpus(per, first_token=r * per, **kw).tobytes() for r in range(world))
        want = oracle.train(body, mintokens=mintokens, maxlength=maxlength, doskipgrams_exhaustive=int(skip), streamed=0 if skip else 1)
        ok = merged == want.as_dict() and [tuple(p) for p in passes] == want.passes and (head["tokens"], head["types"], head["maxn"], head["minn"]) == (
            want.tokens, want.types, want.maxn, want.minn)
        print("DIST_RESULT", "OK" if ok else "MISMATCH", len(merged), len(want), passes, want.passes, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def constrained(per, kw, maxlength, mintokens, rank, world, local, corpus):
    """Sharded constrained training (SURVEY 8f-2 x 8e): the stage-1 model of another corpus, replicated; local counting; one all-reduce."""
    kw1 = dict(kw, seed=kw["seed"] + 1)
    stage1 = cb.train(cb.Corpus.synthetic(per, device=local, **kw1), MINTOKENS=2, MAXLENGTH=maxlength, QUIET=1, device=local)
    # every rank must number the patterns identically: in practice they all read the same model file; here rank 0's bytes are broadcast
    # (the export order of a trained model depends on which thread claimed which table slot, so two ranks' files differ in order)
    box = [stage1.to_bytes() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    blob = box[0]
    constrain = cb.load_model(blob, MINTOKENS=mintokens, MAXLENGTH=maxlength, QUIET=1, device=local)
    opts = cb.PatternModelOptions(MINTOKENS=mintokens, MAXLENGTH=maxlength, streamed=1, QUIET=1, device=local)
    eng = mg.CudaConstrainedEngine(corpus, constrain, opts, local, torch)
    model = mg.train_constrained_distributed(eng, dist, torch, inplace=False)
    keys, off, counts, _ = model.export()
    mine = {keys[int(off[i]):int(off[i + 1])].tobytes(): int(counts[i]) for i in range(len(counts))}
    head = (model.tokens(), model.types(), model.passes())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine, head), gathered, dst=0)
    ok = True
    if rank == 0:
        import oracle

        body = b"".join(oracle.synth_corpus(per, first_token=r * per, **kw).tobytes() for r in range(world))
        want = oracle.train_constrained(body, oracle.load_model(blob, mintokens=mintokens, maxlength=maxlength), inplace=False, mintokens=mintokens, maxlength=maxlength, streamed=1)
        ok = all(g[0] == want.as_dict() and g[1] == (want.tokens, want.types, want.passes) for g in gathered)
        print("DIST_RESULT", "OK" if ok else "MISMATCH", len(mine), len(want), head, (want.tokens, want.types, want.passes), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
