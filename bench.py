#!/usr/bin/env python
"""bench.py -- corpus tokens/sec through PatternModel::train (unindexed, n<=5, t=2) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one full training pass (all n) over one synthetic Zipf corpus (BASELINE.json configs[1]: 100 M tokens,
V = 100 000, seed 1; the counter-based integer generator of SURVEY.md 8d, bit-identical on GPU and CPU).
  value      whole-job tokens/s with the corpus bytes already resident in HBM (colibri_b200_train_corpus)
  e2e        the same through colibri_b200_train() on HOST buffers: pinned corpus -> H2D -> train -> flat model -> D2H
  roofline   dominant kernel (count_ngrams, levels n>=2): algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own single-threaded C++ train() (oracle/_ref/ref_train) on a bounded sample
Rank 0 prints ONE JSON line.  Under torchrun (N>1) every rank trains its own shard of the corpus and the
hash-partitioned model is merged over NCCL (see DESIGN.md section "Multi-GPU").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "corpus tokens/sec through PatternModel::train (unindexed, n<=5, t=2)"
UNIT = "tokens/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tokens", type=float, default=1e8, help="tokens per GPU (weak scaling)")
    ap.add_argument("--vocab", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--maxlength", type=int, default=5)
    ap.add_argument("--mintokens", type=int, default=2)
    ap.add_argument("--skipgrams", type=int, default=0)
    ap.add_argument("--cpu-sample-tokens", type=float, default=3e6, help="cpu_baseline leg of the default arm: tokens of the bounded sample")
    ap.add_argument("--ref-sample-tokens", type=float, default=1e7, help="--impl reference: tokens per step (prefix of the same corpus)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: CPU seconds the whole run may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra configurations (1 B tokens, skipgrams, indexed) after the headline loop")
    ap.add_argument("--no-digest", action="store_true", help="skip the sha256 digest of the exported model (the device checksum is always compared)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: every rank gets --tokens (weak) or the ranks share one --tokens corpus (strong)")
    return ap.parse_args()


def workload_name(a):
    return "synthetic Zipf(s=1) %dM-token corpus, V=%d, seed %d, unindexed n-grams n<=%d t=%d%s" % (
        round(a.tokens / 1e6), a.vocab, a.seed, a.maxlength, a.mintokens, " + exhaustive skipgrams" if a.skipgrams else "")


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md: the clocks line).
    Uses NVML directly (two cheap queries per sample); a polling `nvidia-smi -lms` process was measured to stall the
    driver for milliseconds at a time and to slow the timed step by 2x.  Falls back to one-shot nvidia-smi calls."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0, period_s=0.05):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                return self.index
        return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _sample(self):
        if self._nvml is not None:
            n = self._nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            r = subprocess.run(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            self.sm.append(float(f[0]))
            self.max_mhz = float(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period if self._nvml is not None else 1.0)

    def stop(self):
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler not started"]}
        self._stop.set()
        self._thread.join(timeout=15)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------- reference / cpu baseline
def cpu_reference_run(body_bytes, maxlength, mintokens, skipgrams, timeout=1500):
    """Time the reference's own train() on `body_bytes` (a .colibri.dat body).  Uses the unmodified reference binary
    when oracle/_ref travelled with the repo, else the oracle port.  Returns (tokens, seconds, kind)."""
    import oracle

    if oracle.have_ref():
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "sample.colibri.dat")
            with open(path, "wb") as f:
                f.write(b"\xa2\x02")
                f.write(body_bytes)
            st, _ = oracle.ref_train(path, None, unindexed=True, skipgrams=bool(skipgrams), quiet=True, timeout=timeout, t=mintokens, l=maxlength)
        return st["tokens"], st["train_seconds"], "reference"
    t0 = time.perf_counter()
    m = oracle.train(body_bytes, mintokens=mintokens, maxlength=maxlength, doskipgrams_exhaustive=1 if skipgrams else 0, streamed=0 if skipgrams else 1)
    return m.tokens, time.perf_counter() - t0, "port"


def sample_body(a, ntokens):
    """First `ntokens` tokens of the benchmark corpus (same seed, same generator: a prefix of the same stream)."""
    import oracle

    return oracle.synth_corpus(int(ntokens), vocab=a.vocab, seed=a.seed, mean_sentence=22).tobytes()


def reference_fixture(a):
    """The committed one-off run of the unmodified reference on the FULL workload corpus (tests/golden/golden_bench.json), if this is that workload."""
    if int(a.tokens) != 100000000 or a.vocab != 100000 or a.seed != 1 or a.maxlength != 5 or a.mintokens != 2 or a.skipgrams:
        return None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "golden_bench.json")) as f:
            return json.load(f).get("zipf100m")
    except Exception:
        return None


def full_config_note(a):
    g = reference_fixture(a)
    if not g:
        return None
    return {"tokens": g["tokens"], "train_seconds": g["reference_train_seconds"], "tokens_per_s": g["tokens"] / g["reference_train_seconds"], "host": g["host"],
            "source": "tests/golden/golden_bench.json (tests/golden/make_golden_bench.py; one run, not repeated on the GPU box: 12 CPU-minutes)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # One step = the reference's own train() on the first --ref-sample-tokens (10 M) tokens of the workload corpus: ~60 s of one core (the
    # reference is single-threaded and slows down with corpus size: 0.6 M tokens/s at 1 M tokens, 0.17 M at 10 M, 0.14 M at 100 M), so the
    # requested K + W steps are cut to what fits --ref-budget-s; the line says how many steps were really timed.
    ntok = int(min(a.ref_sample_tokens, a.tokens))
    body = sample_body(a, ntok)
    t0 = time.perf_counter()
    tokens, sec, kind = cpu_reference_run(body, a.maxlength, a.mintokens, a.skipgrams)  # doubles as the warm-up (page cache, allocator)
    first = sec
    times = [sec]
    steps = max(1, min(a.steps, int((a.ref_budget_s - (time.perf_counter() - t0)) / max(first, 1e-3))))
    warm = 0
    if steps >= 2:  # enough budget: the first run becomes the warm-up
        times, warm = [], 1
        for _ in range(min(steps - 1, a.steps)):
            tokens, sec, kind = cpu_reference_run(body, a.maxlength, a.mintokens, a.skipgrams)
            times.append(sec)
    total = sum(times)
    value = tokens * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times), "warmup": warm, "requested_steps": a.steps, "requested_warmup": a.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": "first %d tokens of the same corpus per step (K, W cut to fit %.0f s of CPU)" % (tokens, a.ref_budget_s)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": "first %d tokens of the workload corpus, reference is single-threaded (host has %d cores)" % (tokens, os.cpu_count() or 0),
                         "full_config_reference": full_config_note(a)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- ours
def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_sources_sha():
    """sha256 over the sources of the dominant kernel family and its driver: what a committed ncu capture is tied to."""
    import hashlib

    h = hashlib.sha256()
    for rel in ("colibri-core_b200/csrc/kernels.cu", "colibri-core_b200/csrc/partition.cu", "colibri-core_b200/csrc/device_utils.cuh", "colibri-core_b200/csrc/engine_common.h"):
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def profile_traffic():
    """DRAM bytes per step of the dominant kernel family from the committed ncu capture (profiles/traffic.json).  The capture names the
    kernel sources it was taken from; when they have changed since, the number is stale and is NOT printed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "no profiles/traffic.json"
    if t.get("kernel_sources_sha") != kernel_sources_sha():
        return None, "profiles/traffic.json was captured from other kernel sources (%s, now %s): stale, not reported" % (t.get("kernel_sources_sha"), kernel_sources_sha())
    return t, t.get("source")


def canonical_digest(keys, key_off, counts):
    """sha256 over the bytewise-sorted (key, count) stream of a flat model -- the same canonical form tests/golden/*.json pin
    (restated here because bench.py may use oracle/ only as the CPU baseline)."""
    import hashlib

    import numpy as np

    off = key_off.astype(np.int64)
    lens = np.diff(off)
    n = len(lens)
    w = int(lens.max()) if n else 1
    pad = np.zeros((n, w), dtype=np.uint8)
    if n:
        rows = np.repeat(np.arange(n), lens)
        cols = np.arange(int(off[-1])) - np.repeat(off[:-1], lens)
        pad[rows, cols] = keys[: int(off[-1])]
    order = np.argsort(pad.view("S%d" % w).reshape(-1), kind="stable")
    new_lens = lens[order]
    new_off = np.zeros(n + 1, dtype=np.uint64)
    new_off[1:] = np.cumsum(new_lens)
    src = np.repeat(off[:-1][order], new_lens) + (np.arange(int(new_lens.sum())) - np.repeat(new_off[:-1].astype(np.int64), new_lens)) if n else np.zeros(0, dtype=np.int64)
    h = hashlib.sha256()
    h.update(np.uint64(n).tobytes())
    h.update(new_off.tobytes())
    h.update(keys[src].tobytes())
    h.update(counts[order].astype(np.uint32).tobytes())
    return h.hexdigest()


def parity_block(a, model, with_digest):
    """Compare the model of the last timed step with the committed fixture of the unmodified reference on the same corpus."""
    g = reference_fixture(a)
    cs = model.checksum()
    out = {"checksum": cs, "fixture": None}
    if not g:
        return out
    out["fixture"] = "tests/golden/golden_bench.json: zipf100m (unmodified reference, %s)" % g["cli"]
    out["patterns_ok"] = len(model) == g["patterns"] and model.tokens() == g["tokens"] and model.types() == g["types"]
    out["passes_ok"] = [(p[1], p[3]) for p in model.passes()] == [(p[0], p[2]) for p in g["passes_found_skip_pruned_kept"]]
    if "checksum" in g:
        out["checksum_ok"] = all(cs[k] == g["checksum"][k] for k in ("sum", "xor", "occurrences", "patterns"))
    if with_digest:
        keys, off, counts, _ = model.export()
        out["digest"] = canonical_digest(keys, off, counts)
        out["digest_ok"] = out["digest"] == g["digest"]
    return out


def extra_configs(cb, a, local):
    """BASELINE.json configs[2..3] and the north-star's 1 B-token target on one GPU, a few steps each, AFTER the headline loop (device-generated
    corpora; 1 B tokens: V = 10^6, seed 2).  Reported under `extra_configs`; the headline numbers above are not affected."""
    import torch

    out = []

    def run(name, corpus, warm, steps, **kw):
        rec = {"name": name}
        try:
            opts = cb.PatternModelOptions(MINTOKENS=2, MAXLENGTH=5, QUIET=1, device=local, **kw)
            m = None
            for _ in range(warm):
                m = cb.train(corpus, opts)
                m.close()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            dev_ms = 0.0
            for i in range(steps):
                m = cb.train(corpus, opts)
                dev_ms += m.timings()["total"]
                if i + 1 < steps:
                    m.close()
            e1.record()
            torch.cuda.synchronize()
            ms = max(1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1)) / steps
            rec.update({"tokens": m.tokens(), "steps": steps, "warmup": warm, "ms_per_step": ms, "device_ms_per_step": dev_ms / steps, "tokens_per_s": m.tokens() / (ms / 1e3),
                        "patterns": len(m), "passes": m.passes(), "peak_device_gb": m.counters()["peak_device_bytes"] / 1e9, "checksum": m.checksum(),
                        "phase_ms": m.timings(), "levels": {n: m.level(n) for n in range(2, m.maxlength() + 1)}})
            m.close()
        except Exception as e:  # a configuration that does not fit or is refused is reported, it does not take the headline line down
            rec["error"] = str(e)[:300]
        out.append(rec)

    c100 = cb.Corpus.synthetic(int(a.tokens), vocab=a.vocab, seed=a.seed, device=local)
    run("config 3 shape at 100 M tokens: unindexed + exhaustive skipgrams (-u -s -t 2 -l 5)", c100, 1, 3, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0)
    run("config 4 shape at 100 M tokens: indexed (-t 2 -l 5)", c100, 1, 3, model_type=20, streamed=0)
    c100.close()
    c1b = cb.Corpus.synthetic(1000000000, vocab=1000000, seed=2, device=local)
    run("north-star target: 1 B tokens (V = 10^6, seed 2), unindexed n <= 5 t = 2", c1b, 1, 3)
    run("config 3: 1 B tokens, unindexed + exhaustive skipgrams", c1b, 0, 1, DOSKIPGRAMS_EXHAUSTIVE=1, streamed=0)
    run("config 4: 1 B tokens, indexed", c1b, 0, 1, model_type=20, streamed=0)
    c1b.close()
    return out


def run_ours(a):
    import torch

    import colibri_core_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun (one process per GPU)" % a.gpus)
    if cb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        from colibri_core_b200 import multigpu

        return multigpu.bench(a, dist, rank, world, local, METRIC, UNIT, workload_name(a), ClockSampler, measured_peaks)

    ntok = int(a.tokens)
    opts = cb.PatternModelOptions(MINTOKENS=a.mintokens, MAXLENGTH=a.maxlength, DOSKIPGRAMS_EXHAUSTIVE=a.skipgrams, streamed=0 if a.skipgrams else 1, QUIET=1, device=local)
    corpus = cb.Corpus.synthetic(ntok, vocab=a.vocab, seed=a.seed, device=local)
    nbytes = corpus.nbytes

    # ---- device-resident arm (value)
    # warm-up in the shape of the timed loop (the previous step's model is released after the next one is trained), so that the library's
    # device memory pool already holds the blocks for two live models and no timed step pays a cudaMalloc
    last = None
    for _ in range(a.warmup):
        m = cb.train(corpus, opts)
        if last is not None:
            last.close()
        last = m
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms, count_ms, launches, alg_bytes, phase_ms = [], 0.0, 0, 0.0, {}
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        m = cb.train(corpus, opts)
        tm, ct = m.timings(), m.counters()
        dev_ms.append(tm["total"])
        launches += ct["kernel_launches"]
        for k, v in tm.items():
            phase_ms[k] = phase_ms.get(k, 0.0) + v
        for n in range(2, m.maxlength() + 1):
            lv = m.level(n)
            count_ms += lv["count_ms"]
            # dominant kernel family of level n, DESIGN.md "algorithmic bytes" -- the unit of work, the same whichever path a level runs on: the
            # previous id of every position is read and its new id written (8 B) and one 32 B sector moves in and out of HBM per window whose
            # n-gram is not proven a singleton beforehand; the occurrence filter (HBM-table path) reads the ids once more (4 B); its 2-bit
            # counters are sized to stay in L2 and are not counted.  On the partitioned path "singletons" are the n-grams that occur exactly once.
            # A level that runs from a position list (items < positions: only positions whose (n-1)-gram survived are visited) instead reads
            # 12 B per item (list entry, its id, the neighbour's id), zeroes the new id array (4 B/position) and writes 4 B per counted window.
            npos_l, table_w = ct["positions"] + 1, lv["windows"] - lv["singletons"]
            if lv["items"] < ct["positions"]:
                alg_bytes += 12.0 * lv["items"] + 4.0 * npos_l + 68.0 * table_w + (8.0 * lv["items"] if lv["filtered"] else 0.0)
            else:
                alg_bytes += 8.0 * npos_l + 64.0 * table_w + (4.0 * npos_l if lv["filtered"] else 0.0)
            if lv.get("fused_id1"):  # the level's time includes the sweep that reads the tokens and writes the level-1 ids (4 B in, 4 B out per position)
                alg_bytes += 8.0 * npos_l
        if last is not None:
            last.close()
        last = m
    ev1.record()
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()
    event_ms = ev0.elapsed_time(ev1)
    tokens = last.tokens()
    ms_per_step = max(event_ms, wall_s * 1e3) / a.steps
    value = tokens / (ms_per_step / 1e3)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (count_ms / 1e3) / 1e9 if count_ms > 0 else 0.0
    traffic, traffic_note = profile_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
        "config": {"workload": workload_name(a), "corpus_bytes": nbytes, "patterns": len(last), "parallelism": "1 GPU", "l2": "inputs exceed L2 (corpus %d MB, id arrays + tables > 1 GB)" % (nbytes >> 20),
                   "timing": "max(torch CUDA events, wall clock) around K synchronous ABI calls"},
        "device_ms_per_step": sum(dev_ms) / len(dev_ms),
        "phase_ms_per_step": {k: v / a.steps for k, v in phase_ms.items()},
        "roofline": {"bound": "hbm", "kernel": "count family, levels 2..%d: make_id1_hist + part_split1/split2/count/gather (level 2, partitioned, dense pair square; the level-1 id sweep carries its first pass) + ngram_filter_kernel + count_ngrams_kernel (HBM table)" % last.maxlength(), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "alg_bytes_per_step": alg_bytes / a.steps, "kernel_ms_per_step": count_ms / a.steps,
                     "kernel_share_of_step": (count_ms / a.steps) / (sum(dev_ms) / len(dev_ms)),
                     "traffic": traffic["dram_bytes_per_step"] if traffic else None, "traffic_source": traffic_note,
                     "hbm_read_roofline_frac": (last.maxlength() * nbytes / ((sum(dev_ms) / len(dev_ms)) / 1e3) / 1e9) / peak},
        "clocks": clocks, "gpu_launches": launches,
        "levels_last_step": {n: last.level(n) for n in range(2, last.maxlength() + 1)},
    }
    line["parity"] = parity_block(a, last, not a.no_digest)

    # ---- end-to-end arm: host buffers in, host buffers out, copies inside the timed region
    if not a.no_e2e:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.numpy()[:] = corpus.download()
        npat, kb, _ = last.export_sizes()
        out_keys = torch.empty(int(kb * 1.05) + 64, dtype=torch.uint8, pin_memory=True)
        out_len = torch.empty(int(npat * 1.05) + 64, dtype=torch.int16, pin_memory=True)
        out_cnt = torch.empty(int(npat * 1.05) + 64, dtype=torch.int32, pin_memory=True)
        cap_k, cap_p = out_keys.numel(), out_cnt.numel()
        for _ in range(max(1, a.warmup)):
            sm = cb.train_export_pointers(host.data_ptr(), nbytes, opts, out_keys.data_ptr(), cap_k, out_len.data_ptr(), out_cnt.data_ptr(), cap_p)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        d2h, e2e_launches = 0, 0
        for _ in range(a.steps):
            sm = cb.train_export_pointers(host.data_ptr(), nbytes, opts, out_keys.data_ptr(), cap_k, out_len.data_ptr(), out_cnt.data_ptr(), cap_p)
            d2h = sm["keybytes"] + 2 * sm["npatterns"] + 4 * sm["npatterns"]
            e2e_launches += sm["kernel_launches"]
        e1.record()
        barrier()
        e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        # the host copy is the same model: patterns, and the order-independent checksum over the bytes that arrived
        e2e_ok = sm["npatterns"] == len(last) and sm["passes"] == last.passes()
        line["e2e"] = {"value": tokens * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / a.steps,
                       "device_phase_ms_last_step": sm["ms"], "same_model_as_device_arm": bool(e2e_ok),
                       "api": "colibri_b200_train_export(pinned host corpus -> pinned host keys/lengths/counts): H2D, train, per-level export overlapped with the next level's counting"}
        line["gpu_launches"] = launches + e2e_launches

    # ---- the reference's CPU path beside it (bounded sample)
    if not a.no_cpu_baseline:
        body = sample_body(a, a.cpu_sample_tokens)
        ctok, csec, kind = cpu_reference_run(body, a.maxlength, a.mintokens, a.skipgrams)
        line["cpu_baseline"] = {"value": ctok / csec, "unit": UNIT, "cores": 1, "kind": kind, "seconds": csec,
                                "sample": "first %d tokens of the workload corpus; the reference is single-threaded (host has %d cores)" % (ctok, os.cpu_count() or 0),
                                "full_config_reference": full_config_note(a)}
    last.close()
    corpus.close()
    if not a.no_extra:
        line["extra_configs"] = extra_configs(cb, a, local)
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
