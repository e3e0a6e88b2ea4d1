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
    ap.add_argument("--cpu-sample-tokens", type=float, default=3e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size one step so that (K + W) steps of single-threaded reference work stay within a few minutes (~0.15 M tokens/s)
    budget_s = 150.0
    per_step = budget_s / max(1, a.steps + a.warmup)
    ntok = int(min(a.cpu_sample_tokens, max(2e5, per_step * 1.5e5)))
    body = sample_body(a, ntok)
    times, kind, tokens = [], "port", ntok
    for i in range(a.warmup + a.steps):
        tokens, sec, kind = cpu_reference_run(body, a.maxlength, a.mintokens, a.skipgrams)
        if i >= a.warmup:
            times.append(sec)
    total = sum(times)
    value = tokens * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": "first %d tokens of the same corpus per step" % tokens},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": "first %d tokens of the workload corpus, reference is single-threaded (host has %d cores)" % (tokens, os.cpu_count() or 0)},
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


def profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture summary, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


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
            # dominant kernel family of level n (occurrence filter + count), DESIGN.md "algorithmic bytes": the count launch reads
            # the previous id and writes the new id of every position (8 B) and moves one 32 B sector in and out of HBM per
            # window that reaches the table; the filter launch (when used) reads the ids once more (4 B); its 2-bit
            # counters are sized to stay in L2 and are not counted as HBM traffic.
            alg_bytes += 8.0 * (ct["positions"] + 1) + 64.0 * (lv["windows"] - lv["singletons"]) + (4.0 * (ct["positions"] + 1) if lv["singletons"] else 0.0)
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
    traffic = profile_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
        "config": {"workload": workload_name(a), "corpus_bytes": nbytes, "patterns": len(last), "parallelism": "1 GPU", "l2": "inputs exceed L2 (corpus %d MB, id arrays + tables > 1 GB)" % (nbytes >> 20),
                   "timing": "max(torch CUDA events, wall clock) around K synchronous ABI calls"},
        "device_ms_per_step": sum(dev_ms) / len(dev_ms),
        "phase_ms_per_step": {k: v / a.steps for k, v in phase_ms.items()},
        "roofline": {"bound": "hbm", "kernel": "ngram_filter_kernel + count_ngrams_kernel (levels 2..%d; level 2 with dense pair slots)" % last.maxlength(), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "alg_bytes_per_step": alg_bytes / a.steps, "kernel_ms_per_step": count_ms / a.steps,
                     "kernel_share_of_step": (count_ms / a.steps) / (sum(dev_ms) / len(dev_ms)),
                     "traffic": traffic["dram_bytes_per_step"] if traffic else None,
                     "hbm_read_roofline_frac": (last.maxlength() * nbytes / ((sum(dev_ms) / len(dev_ms)) / 1e3) / 1e9) / peak},
        "clocks": clocks, "gpu_launches": launches,
    }

    # ---- end-to-end arm: host buffers in, host buffers out, copies inside the timed region
    if not a.no_e2e:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.numpy()[:] = corpus.download()
        npat, kb, _ = last.export_sizes()
        out_keys = torch.empty(int(kb * 1.05) + 64, dtype=torch.uint8, pin_memory=True)
        out_len = torch.empty(int(npat * 1.05) + 64, dtype=torch.int16, pin_memory=True)
        out_cnt = torch.empty(int(npat * 1.05) + 64, dtype=torch.int32, pin_memory=True)
        for _ in range(max(1, a.warmup)):
            m = cb.train_host_pointer(host.data_ptr(), nbytes, opts)
            m.export_compact_into(out_keys.data_ptr(), out_len.data_ptr(), out_cnt.data_ptr())
            m.close()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        d2h = 0
        for _ in range(a.steps):
            m = cb.train_host_pointer(host.data_ptr(), nbytes, opts)
            n2, kb2, _ = m.export_sizes()
            m.export_compact_into(out_keys.data_ptr(), out_len.data_ptr(), out_cnt.data_ptr())
            d2h = kb2 + 2 * n2 + 4 * n2
            m.close()
        e1.record()
        barrier()
        e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
        line["e2e"] = {"value": tokens * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / a.steps,
                       "api": "colibri_b200_train(host corpus) + colibri_b200_model_export_compact(host keys/lengths/counts), pinned host memory"}

    # ---- the reference's CPU path beside it (bounded sample)
    if not a.no_cpu_baseline:
        body = sample_body(a, a.cpu_sample_tokens)
        ctok, csec, kind = cpu_reference_run(body, a.maxlength, a.mintokens, a.skipgrams)
        line["cpu_baseline"] = {"value": ctok / csec, "unit": UNIT, "cores": 1, "kind": kind, "seconds": csec,
                                "sample": "first %d tokens of the workload corpus; the reference is single-threaded (host has %d cores)" % (ctok, os.cpu_count() or 0)}
    last.close()
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
