"""pytest configuration: markers, repo root on sys.path, shared golden fixtures."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


_corpus_cache = {}


def corpus_body(golden, name) -> bytes:
    """Realise a corpus named in golden.json: the body of a .colibri.dat (bytes after the 0xA2 0x02 header)."""
    if name in _corpus_cache:
        return _corpus_cache[name]
    spec = golden["corpora"][name]
    if spec["kind"] == "file":
        body = open(os.path.join(GOLDEN_DIR, spec["arg"]), "rb").read()[2:]
    elif spec["kind"] == "hex":
        body = bytes.fromhex(spec["arg"])
    else:
        import oracle

        body = oracle.synth_corpus(**spec["arg"]).tobytes()
    _corpus_cache[name] = body
    return body


def case_id(case) -> str:
    return "%s-%s%s-%s" % (case["corpus"], "u" if case["unindexed"] else "i", "s" if case["skipgrams"] else "", "".join("%s%s" % kv for kv in sorted(case["cli"].items())) or "default")


def load_cases():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)["cases"]


# ---- constrained training: golden_constrained.json, written by tests/golden/make_golden_constrained.py
def load_constrained_cases():
    with open(os.path.join(GOLDEN_DIR, "golden_constrained.json")) as f:
        return json.load(f)["cases"]


def constrained_case_id(case) -> str:
    return "%s<-%s-%s%s-%s" % (case["corpus"], case["stage1_corpus"], case["mode"], "u" if case["unindexed"] else "i", "".join("%s%s" % kv for kv in sorted(case["cli"].items())))


def cli_load_and_train_options(case):
    """What the reference CLI does with a constrained case before train() runs (src/patternmodeller.cpp:717-737, :777-852):
    returns (load filter kwargs, train option kwargs without the widened lengths, inplace)."""
    cli = case["cli"]
    t, l, m = cli.get("t", -1), cli.get("l", 100), cli.get("m", 1)
    indexed = 0 if case["unindexed"] else 1
    if case["mode"] == "I":  # in-place rebuild: model loaded AS the output type with DORESET, corpus preloaded
        return dict(mintokens=t, minlength=m, maxlength=l, doreset=1, indexed=indexed), dict(mintokens=t, maxlength=l, minlength=m, indexed=indexed, streamed=0), True
    # -j: PatternSetModel(file, options); an unindexed output model streams the corpus file
    return dict(mintokens=t, minlength=m, maxlength=l, indexed=0), dict(mintokens=t, maxlength=l, minlength=m, indexed=indexed, streamed=1 if case["unindexed"] else 0), False


# ---- the large-corpus machinery of the device path, forced onto small inputs (csrc/engine_common.h: Tuning).
# bench.py runs 100 M tokens per GPU, where the occurrence filter, the per-block hot-key cache, the dense pair slots of level 2 and the
# list mode of the sparse levels are all on; the oracle cannot follow there in seconds, so the parity suite switches each of them on by
# environment knob on corpora it can check, with a filter so small (2^8 .. 2^12 buckets) that buckets collide all the time.
FORCED_PATHS = {
    "default": {},
    "filter": {"COLIBRI_B200_FILTER_MIN": "0", "COLIBRI_B200_FILTER_LOG2_MIN": "8", "COLIBRI_B200_FILTER_LOG2": "12", "COLIBRI_B200_HOT": "0", "COLIBRI_B200_SPARSE_DIV": "0",
               "COLIBRI_B200_DENSE": "0", "COLIBRI_B200_FILTER_1BIT": "0"},  # (the 2-bit counters read directly; the other paths read their packed "hit twice" bits)
    "filter+hot+dense": {"COLIBRI_B200_FILTER_MIN": "0", "COLIBRI_B200_FILTER_LOG2_MIN": "10", "COLIBRI_B200_FILTER_LOG2": "16", "COLIBRI_B200_HOT": "2", "COLIBRI_B200_DENSE_MIN": "0",
                         "COLIBRI_B200_DENSE": "48", "COLIBRI_B200_SPARSE_DIV": "0"},
    "list": {"COLIBRI_B200_SPARSE_DIV": "1", "COLIBRI_B200_HOT": "0"},
    "bench": {"COLIBRI_B200_FILTER_MIN": "0", "COLIBRI_B200_FILTER_LOG2_MIN": "12", "COLIBRI_B200_FILTER_LOG2": "20", "COLIBRI_B200_HOT": "2", "COLIBRI_B200_DENSE_MIN": "0",
              "COLIBRI_B200_DENSE": "3072", "COLIBRI_B200_SPARSE_DIV": "1"},
    # the partitioned counting path (csrc/partition.cu: what bench.py's large levels run), alone and with the dense square + list mode around it
    "part": {"COLIBRI_B200_PART_MIN": "0", "COLIBRI_B200_PART_ALL": "1", "COLIBRI_B200_SPARSE_DIV": "0", "COLIBRI_B200_DENSE": "0"},
    "part+dense+list": {"COLIBRI_B200_PART_MIN": "0", "COLIBRI_B200_PART_ALL": "1", "COLIBRI_B200_DENSE_MIN": "0", "COLIBRI_B200_DENSE": "48", "COLIBRI_B200_SPARSE_DIV": "1"},
    # what bench.py runs at 100 M tokens: level 2 partitioned around its dense square, the later levels on the HBM table with filter, hot keys and list mode
    "part-bench": {"COLIBRI_B200_PART_MIN": "0", "COLIBRI_B200_DENSE_MIN": "0", "COLIBRI_B200_DENSE": "3072", "COLIBRI_B200_SPARSE_DIV": "4", "COLIBRI_B200_FILTER_MIN": "0",
                   "COLIBRI_B200_FILTER_LOG2_MIN": "12", "COLIBRI_B200_FILTER_LOG2": "20", "COLIBRI_B200_HOT": "2"},
}


def force_path(monkeypatch, name):
    for k in ("COLIBRI_B200_FILTER_MIN", "COLIBRI_B200_FILTER_LOG2_MIN", "COLIBRI_B200_FILTER_LOG2", "COLIBRI_B200_HOT", "COLIBRI_B200_HOT_MIN", "COLIBRI_B200_DENSE",
              "COLIBRI_B200_DENSE_MIN", "COLIBRI_B200_SPARSE_DIV", "COLIBRI_B200_NO_FILTER", "COLIBRI_B200_PART_MIN", "COLIBRI_B200_PART_ALL", "COLIBRI_B200_FILTER_1BIT"):
        monkeypatch.delenv(k, raising=False)
    for k, v in FORCED_PATHS[name].items():
        monkeypatch.setenv(k, v)
