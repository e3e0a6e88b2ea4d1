"""The binding a Colibri Core maintainer adds (examples/reference_binding/b200_patternmodel.h), run for real: the
UNMODIFIED reference's PatternModel<uint32_t> with ONE override of train() that calls the C ABI.  oracle/_ref/ref_binding_check (built by
`make -C oracle ref` where /root/reference is mounted; the binary travels to the GPU box) trains a corpus through the reference's own CPU
train() and through the override in one process and compares the models with the reference's own accessors."""
import os
import subprocess
import sys
import tempfile

import pytest

import oracle
from conftest import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "oracle", "_ref", "ref_binding_check")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(CHECK), reason="oracle/_ref/ref_binding_check is not built (needs /root/reference)")]


@pytest.mark.parametrize("args", [["-t", "2", "-l", "3"], ["-t", "2", "-l", "8", "-p"], ["-t", "1", "-l", "4"], ["-t", "2", "-l", "6", "-s"]], ids=lambda a: "".join(a))
def test_reference_class_with_the_b200_override_builds_the_reference_model_hamlet(args):
    r = subprocess.run([CHECK, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")] + args, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("IDENTICAL"), r.stdout + r.stderr


@pytest.mark.parametrize("args", [["-i", "-t", "2", "-l", "3"], ["-i", "-t", "2", "-l", "8"], ["-i", "-t", "1", "-l", "4"], ["-i", "-S", "-t", "2", "-l", "6"]], ids=lambda a: "".join(a))
def test_reference_indexed_class_with_the_b200_override_builds_the_reference_model_hamlet(args):
    """IndexedPatternModel<> with the same override: the (sentence, token) list of every pattern, -S with trainskipgrams (the 133-pattern KAT)."""
    r = subprocess.run([CHECK, "-f", os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")] + args, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("IDENTICAL indexed"), r.stdout + r.stderr


def test_reference_indexed_class_with_the_b200_override_synthetic():
    body = oracle.synth_corpus(200000, vocab=3000, seed=4, mean_sentence=15, phrase_permille=150, nphrases=400)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.colibri.dat")
        with open(path, "wb") as f:
            f.write(b"\xa2\x02" + body.tobytes())
        r = subprocess.run([CHECK, "-f", path, "-i", "-t", "2", "-l", "5"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("IDENTICAL indexed"), r.stdout + r.stderr


@pytest.mark.parametrize("seed,skip", [(1, False), (2, False), (3, True)])
def test_reference_class_with_the_b200_override_builds_the_reference_model_synthetic(seed, skip):
    body = oracle.synth_corpus(200000, vocab=3000, seed=seed, mean_sentence=15, phrase_permille=150, nphrases=400)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.colibri.dat")
        with open(path, "wb") as f:
            f.write(b"\xa2\x02" + body.tobytes())
        r = subprocess.run([CHECK, "-f", path, "-t", "2", "-l", "5"] + (["-s"] if skip else []), capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("IDENTICAL"), r.stdout + r.stderr
