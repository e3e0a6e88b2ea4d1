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
