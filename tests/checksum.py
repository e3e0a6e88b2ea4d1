"""Host-side (numpy) restatement of colibri_b200_model_checksum (include/colibri_b200.h), used by the tests and by tests/golden/make_golden_bench.py
to put the reference's own model files on the same scale.  Checker code: the product never imports it."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def fmix64(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return x


def key_hashes(keys, key_off):
    """FNV-1a 64 of every key (vectorised over the patterns, one step per byte position)."""
    off = np.asarray(key_off, dtype=np.int64)
    lens = np.diff(off)
    h = np.full(len(lens), 0xcbf29ce484222325, dtype=np.uint64)
    keys = np.asarray(keys, dtype=np.uint8)
    with np.errstate(over="ignore"):
        for j in range(int(lens.max()) if len(lens) else 0):
            sel = np.nonzero(lens > j)[0]
            h[sel] = (h[sel] ^ keys[off[:-1][sel] + j].astype(np.uint64)) * np.uint64(0x100000001b3)
    return h


def model_checksum(keys, key_off, counts, ref_off=None, ref_sentence=None, ref_token=None):
    counts = np.asarray(counts, dtype=np.uint64)
    h = key_hashes(keys, key_off)
    with np.errstate(over="ignore"):
        v = fmix64(h ^ (counts * np.uint64(0x9E3779B97F4A7C15)))
        out = {"sum": int(np.add.reduce(v, dtype=np.uint64)) if len(v) else 0, "xor": int(np.bitwise_xor.reduce(v)) if len(v) else 0,
               "occurrences": int(counts.sum()), "patterns": int(len(counts)), "refsum": 0, "refs": 0}
        if ref_off is not None and len(counts):
            ro = np.asarray(ref_off, dtype=np.int64)
            owner = np.repeat(np.arange(len(counts)), np.diff(ro))
            r = (np.asarray(ref_sentence, dtype=np.uint64) << np.uint64(16)) | np.asarray(ref_token, dtype=np.uint64)
            rv = fmix64(h[owner] ^ (r * np.uint64(0xD6E8FEB86659FD93)))
            out["refsum"] = int(np.add.reduce(rv, dtype=np.uint64)) if len(rv) else 0
            out["refs"] = int(len(rv))
    return out


def flat_checksum(fm):
    """Checksum of an oracle.FlatModel."""
    return model_checksum(fm.keys, fm.key_off, fm.counts, fm.ref_off, fm.ref_sentence, fm.ref_token)
