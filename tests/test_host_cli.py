"""The drop-in front end: colibri-patternmodeller (B200 build) and the C++ PatternModel API mirror in colibri-core_b200/host/.
CPU: it builds, parses the reference's flags and fails loudly without a GPU.  GPU: its model files equal the reference's."""
import os
import subprocess
import tempfile

import pytest

import oracle
from conftest import GOLDEN_DIR, ROOT, load_cases

CLI = os.path.join(ROOT, "colibri-core_b200", "bin", "colibri-patternmodeller")


@pytest.fixture(scope="module")
def cli():
    subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True)
    assert os.path.exists(CLI)
    return CLI


def test_cli_usage_and_refusals(cli):
    r = subprocess.run([cli, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "-f FILE" in r.stderr
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([cli, "-f", "/nonexistent.colibri.dat", "-u", "-o", "/tmp/x"], capture_output=True, text=True)
    assert r.returncode == 2 and "Can't open corpus data" in r.stderr  # reference src/patternmodeller.cpp:749-751
    r = subprocess.run([cli, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat"), "-u"], capture_output=True, text=True)
    assert r.returncode == 2 and "Ooops" in r.stderr  # reference :296-301
    r = subprocess.run([cli, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat"), "-P"], capture_output=True, text=True)
    assert r.returncode == 2 and "not part of the B200 training front end" in r.stderr


def test_cli_fails_loudly_without_gpu(cli):
    import colibri_core_b200 as cb

    if cb.device_count() > 0:
        pytest.skip("checks the behaviour WITHOUT a GPU")
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([cli, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat"), "-u", "-t", "2", "-l", "3", "-o", os.path.join(td, "m")], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
        assert not os.path.exists(os.path.join(td, "m"))


CLI_CASES = [c for c in load_cases() if c["corpus"] in ("hamlet", "republic") and c["options"].get("maxbackofflength", 100) >= 100]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CLI_CASES, ids=["%s-%s%s%s" % (c["corpus"], "u" if c["unindexed"] else "i", "s" if c["skipgrams"] else "", "".join("%s%s" % kv for kv in sorted(c["cli"].items()))) for c in CLI_CASES])
def test_cli_model_files_equal_reference(cli, case):
    """colibri-patternmodeller -f X -u [-s] ... -o M  ->  M parses to the same patterns/counts/header as the reference's file."""
    corpus = os.path.join(GOLDEN_DIR, case["corpus"] + ".colibri.dat")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "m.colibri.patternmodel")
        cmd = [cli, "-f", corpus, "-o", out] + (["-u"] if case["unindexed"] else []) + (["-s"] if case["skipgrams"] else [])
        for k, v in case["cli"].items():
            cmd += ["-" + k, str(v)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        m = oracle.parse_modelfile(open(out, "rb").read())
        assert (m.tokens, m.types, len(m), m.model_type) == (case["tokens"], case["types"], case["patterns"], 10 if case["unindexed"] else 20)
        assert m.digest() == case["digest"]
        # the progress lines are the reference's: " Found X ngrams...pruned Y...total kept: Z"
        assert [(p[0], p[2]) for p in oracle.parse_ref_passes(r.stderr)] == [(p[0], p[2]) for p in case["passes"]]


def test_host_api_mirror_cpp(cli, tmp_path):
    """Compile and run tests/host/test_host_api.cpp against colibri-core_b200/host/*.h (Pattern semantics, hash KATs, refusals)."""
    exe = str(tmp_path / "test_host_api")
    libdir = os.path.join(ROOT, "colibri-core_b200", "lib")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "colibri-core_b200", "host"), os.path.join(ROOT, "tests", "host", "test_host_api.cpp"),
                    "-L" + libdir, "-lcolibri_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe, os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "host API tests ok" in r.stderr and "FAILED" not in r.stderr


# ---- constrained modes of the CLI: -j, -I, -2 and -i alone (load + filter + write)
from conftest import constrained_case_id, corpus_body, load_constrained_cases  # noqa: E402

CONSTRAINED_CLI_CASES = [c for c in load_constrained_cases() if c["corpus"] in ("hamlet", "republic") and c["stage1_corpus"] in ("hamlet", "republic")]


@pytest.mark.gpu
def test_host_api_mirror_cpp_on_gpu(cli, tmp_path):
    """The same program on a GPU box: it also trains an IndexedPatternModel and checks getreverseindex / getleftcooc / computeflexgrams_fromcooc of the
    C++ mirror against answers of the unmodified reference (oracle/_ref/ref_relations on hamlet, -t 2 -l 3)."""
    test_host_api_mirror_cpp(cli, tmp_path)


def test_cli_constrained_refusals(cli, tmp_path):
    hamlet = os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")
    out = str(tmp_path / "m")
    r = subprocess.run([cli, "-f", hamlet, "-j", "/nonexistent.model", "-u", "-o", out], capture_output=True, text=True)
    assert r.returncode == 2 and "No such file" in r.stderr  # reference assert_file_exists, src/patternmodeller.cpp:233-239
    r = subprocess.run([cli, "-i", "/nonexistent.model", "-I", "-f", hamlet, "-o", out], capture_output=True, text=True)
    assert r.returncode == 2 and "No such file" in r.stderr
    r = subprocess.run([cli, "-f", hamlet, "-2", "-t", "2"], capture_output=True, text=True)
    assert r.returncode == 2 and "mandatory for two-stage" in r.stderr  # reference :640-643
    r = subprocess.run([cli, "-f", hamlet, "-2", "-s", "-o", out], capture_output=True, text=True)
    assert r.returncode != 0  # stage 1 needs a GPU here; with one, stage 2 refuses -s under a constraint


@pytest.mark.gpu
@pytest.mark.parametrize("case", CONSTRAINED_CLI_CASES, ids=[constrained_case_id(c) for c in CONSTRAINED_CLI_CASES])
def test_cli_constrained_model_files_equal_reference(cli, golden, case, tmp_path):
    """colibri-patternmodeller -f X -j S | -i S -I ...: the written file parses to the reference CLI's patterns, counts and header."""
    corpus = os.path.join(GOLDEN_DIR, case["corpus"] + ".colibri.dat")
    stage1 = str(tmp_path / "stage1.colibri.patternmodel")
    with open(stage1, "wb") as f:
        f.write(oracle.train_to_modelfile(corpus_body(golden, case["stage1_corpus"]), **case["stage1_options"]))
    out = str(tmp_path / "m.colibri.patternmodel")
    cmd = [cli, "-f", corpus, "-o", out] + (["-i", stage1, "-I"] if case["mode"] == "I" else ["-j", stage1]) + (["-u"] if case["unindexed"] else [])
    for k, v in case["cli"].items():
        cmd += ["-" + k, str(v)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = oracle.parse_modelfile(open(out, "rb").read())
    assert (m.tokens, m.types, len(m), m.model_type) == (case["tokens"], case["types"], case["patterns"], 10 if case["unindexed"] else 20)
    assert m.digest() == case["digest"]
    assert "constrained by another model" in r.stderr
    assert [(p[0], p[2]) for p in oracle.parse_ref_passes(r.stderr)] == [(p[0], p[2]) for p in case["passes"]]


@pytest.mark.gpu
def test_cli_two_stage_and_convert(cli, tmp_path):
    """-2: stage 1 unindexed (.stage1 file) + stage 2 indexed in-place rebuild = the golden case `hamlet I i t2 l4`;  -i alone re-writes a model."""
    hamlet = os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")
    out = str(tmp_path / "two.colibri.patternmodel")
    r = subprocess.run([cli, "-f", hamlet, "-2", "-t", "2", "-l", "4", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "STARTING STAGE 1/2" in r.stderr and "STARTING STAGE 2/2" in r.stderr and os.path.exists(out + ".stage1")
    want = [c for c in load_constrained_cases() if c["corpus"] == "hamlet" and c["mode"] == "I" and not c["unindexed"] and c["cli"] == {"t": 2, "l": 4}][0]
    m = oracle.parse_modelfile(open(out, "rb").read())
    assert m.digest() == want["digest"] and (m.tokens, m.types, m.model_type) == (want["tokens"], want["types"], 20)
    s1 = oracle.parse_modelfile(open(out + ".stage1", "rb").read())
    assert s1.model_type == 10 and len(s1) == 93
    conv = str(tmp_path / "conv.colibri.patternmodel")
    r = subprocess.run([cli, "-i", out, "-u", "-t", "3", "-l", "2", "-o", conv], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = oracle.parse_modelfile(open(conv, "rb").read())
    assert got.same_patterns(oracle.load_model(open(out, "rb").read(), mintokens=3, maxlength=2, indexed=0).flat()) and got.model_type == 10
