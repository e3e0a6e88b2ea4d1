"""World-1 run of the sharded path at bench size (development aid): every shard kernel on one GPU, e.g. under ncu:
   ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/shard_launches.csv python scripts/shard_profile.py [tokens] [nccl|p2p]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RANK", "0")
os.environ.setdefault("WORLD_SIZE", "1")
os.environ.setdefault("LOCAL_RANK", "0")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import colibri_core_b200 as cb  # noqa: E402
import colibri_core_b200.multigpu as mg  # noqa: E402

ntok = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
mode = sys.argv[2] if len(sys.argv) > 2 else "p2p"
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
opts = cb.PatternModelOptions(MINTOKENS=2, MAXLENGTH=5, QUIET=1, device=0)
corpus = cb.Corpus.synthetic(ntok, vocab=100000, seed=1, device=0)
peers = mg.PeerBuffers.get(dist, torch, 1, int(ntok * 1.06) + 1024, 0) if mode == "p2p" else None
for step in range(3):
    eng = mg.CudaShardEngine(corpus, opts, 0, 1, 0)
    if peers is not None:
        eng.use_peers(peers)
    model, passes, head = mg.train_distributed(eng, dist, torch, 2, 5, False)
    print(step, len(model), {k: round(v, 3) for k, v in eng.phase_ms().items()}, flush=True)
    model.close()
    eng.close()
dist.destroy_process_group()
