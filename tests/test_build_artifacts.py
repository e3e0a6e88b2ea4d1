"""What the built library must look like before any GPU time is spent on it (cuobjdump on colibri-core_b200/lib/libcolibri_b200.so, no GPU needed):
sm_100a code only, the register budgets the launch bounds of the hot kernels promise, no big local-memory frames, and the instructions the
design rests on (the 128-bit compare-and-swap that claims a table slot, non-returning REDs for the counters)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "colibri-core_b200", "lib", "libcolibri_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="cuobjdump or the built library is missing")


def run(*args):
    return subprocess.run(["cuobjdump", *args, LIB], capture_output=True, text=True, check=True).stdout


@pytest.fixture(scope="module")
def resources():
    """{demangled-ish kernel name: (registers, stack bytes, shared bytes)}"""
    out = {}
    name = None
    for line in run("--dump-resource-usage").splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            out[name] = tuple(int(x) for x in m.groups())
            name = None
    return out


def test_only_sm_100a_code_is_embedded():
    elfs = [l.split()[-1] for l in run("-lelf").splitlines() if l.startswith("ELF file")]
    assert elfs and all(e.endswith(".sm_100a.cubin") for e in elfs), elfs
    assert "PTX file" not in run("-lptx") or all("sm_100a" in l for l in run("-lptx").splitlines() if l.startswith("PTX file"))


def test_hot_kernels_keep_their_register_budget(resources):
    def regs(fragment):
        found = {k: v for k, v in resources.items() if fragment in k}
        assert found, fragment
        return found

    # 8 blocks of 256 threads per SM (2048 resident threads) need <= 32 registers: the table-path launches are latency-bound and live on occupancy
    for k, (r, stack, _s) in regs("count_ngrams_kernel").items():
        assert r <= 32 and stack <= 32, (k, r, stack)
    for k, (r, stack, _s) in regs("ngram_filter_kernel").items():
        assert r <= 32 and stack == 0, (k, r, stack)
    # the partitioned path: split1 is launched with __launch_bounds__(256, 3) (16 records per thread in registers: a few spilled words are the
    # measured state), split2 with (512, 2); the others may not spill
    for frag, limit, frame in (("part_split1_kernel", 85, 64), ("part_split2_kernel", 64, 0), ("part_count_kernel", 64, 0), ("make_id1_hist_kernel", 40, 0)):
        for k, (r, stack, _s) in regs(frag).items():
            assert r <= limit and stack <= frame, (k, r, stack)
    for frag in ("tokenise_write_kernel", "export_write_kernel", "unigram_hist_kernel"):
        for k, (r, stack, _s) in regs(frag).items():
            assert stack == 0 and r <= 40, (k, r, stack)


def test_no_kernel_has_a_large_local_frame(resources):
    big = {k: v for k, v in resources.items() if v[1] > 256 and "rindex_cooc_kernel" not in k}  # (rindex_cooc keeps up to 32 ids per position in a local array)
    assert not big, big


def test_the_instructions_the_design_rests_on_are_there():
    sass = run("-sass")
    assert "ATOMG.E.CAS.128" in sass  # one 128-bit compare-and-swap claims {key, count = 1, position} of a table slot
    assert "REDG.E.ADD" in sass and "REDG.E.OR" in sass  # counters and filter bits whose old value nobody reads are REDs, not atomics with a return
    assert "ATOMS.CAS.64" in sass  # the per-warp shared-memory tables of the partitioned path
    assert "HMMA" not in sass and "UTCHMMA" not in sass  # integer, HBM-bound work: nothing here belongs on the tensor cores
