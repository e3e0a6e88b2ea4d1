"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/colibri_b200.h
declares, mirrors the reference's option defaults, and refuses to compute without a GPU (no silent fallback)."""
import ctypes
import os
import subprocess

import pytest

import colibri_core_b200 as cb


def test_library_exports_every_declared_symbol():
    lib = cb.library()
    names = cb.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", cb.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(names) <= exported


def test_no_torch_or_cxx_types_in_the_header():
    import re

    text = open(cb.HEADER_PATH).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only, comments (which cite the C++ reference) removed
    for banned in ("torch", "at::", "std::", "template", "class ", "&"):
        assert banned not in code, banned


def test_option_defaults_mirror_patternmodeloptions():
    """include/patternmodel.h:153-180 of the reference."""
    o = cb.PatternModelOptions()
    assert (o.MINTOKENS, o.MINTOKENS_SKIPGRAMS, o.MINTOKENS_UNIGRAMS) == (-1, -1, 1)
    assert (o.MINLENGTH, o.MAXLENGTH, o.MAXBACKOFFLENGTH) == (1, 100, 100)
    assert (o.MINSKIPTYPES, o.MAXSKIPS) == (2, 3)
    assert (o.DOSKIPGRAMS, o.DOSKIPGRAMS_EXHAUSTIVE, o.DOPATTERNPERLINE, o.QUIET, o.DEBUG) == (0, 0, 0, 0, 0)
    with pytest.raises(AttributeError):
        cb.PatternModelOptions(NOSUCHOPTION=1)


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(os.path.dirname(cb.HERE), "colibri-core_b200")
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(root, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "oracle.h" not in text and "from oracle" not in text, os.path.join(root, f)
    out = subprocess.run(["ldd", cb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.mark.skipif(cb.device_count() > 0, reason="checks the behaviour WITHOUT a GPU")
def test_compute_fails_loudly_without_a_gpu():
    with pytest.raises(cb.ColibriError) as ei:
        cb.train(bytes([6, 7, 0]), MINTOKENS=1)
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(cb.ColibriError):
        cb.hash64_batch([b"\x06"])
    with pytest.raises(cb.ColibriError):
        cb.Corpus.synthetic(1000, vocab=100)
    # models that do not come out of train(): upload, load, constrained training -- no host-side stand-in either
    import numpy as np

    with pytest.raises(cb.ColibriError) as ei:
        cb.Model.from_flat(np.array([6, 7], dtype=np.uint8), np.array([0, 1, 2], dtype=np.uint64), np.array([2, 2], dtype=np.uint32))
    assert ei.value.code == 3
    blob = bytes([0, 10, 2]) + (3).to_bytes(8, "little") + (2).to_bytes(8, "little") + (1).to_bytes(8, "little") + bytes([6, 0]) + (3).to_bytes(4, "little")
    with pytest.raises(cb.ColibriError) as ei:
        cb.load_model(blob)
    assert ei.value.code == 3


def test_invalid_arguments_are_reported():
    lib = cb.library()
    assert lib.colibri_b200_train(None, 0, None, None) != 0
    assert b"NULL" in lib.colibri_b200_last_error()
