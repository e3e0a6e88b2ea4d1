"""colibri-core_b200: B200-native PatternModel::train (proycon/colibri-core's training hot path).

This Python module is a thin ctypes binding over the C ABI in ``include/colibri_b200.h`` (implemented by the
sm_100a library ``colibri-core_b200/lib/libcolibri_b200.so``).  It exists for tests, ``bench.py`` and
``torch.distributed`` plumbing; the drop-in front end for reference users is the C++ API mirror and the
``colibri-patternmodeller`` CLI in ``colibri-core_b200/host/``.

There is no CPU fallback: if the shared library is missing this import fails loudly, and every compute call
fails with ``ColibriError`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcolibri_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "colibri_b200.h")

UNINDEXEDPATTERNMODEL = 10
INDEXEDPATTERNMODEL = 20

T_TOTAL, T_TOKENISE, T_UNIGRAMS, T_COUNT, T_SKIPGRAMS, T_PRUNE, T_EXPORT, T_H2D, T_INDEX, T_NPHASES = range(10)
PHASE_NAMES = ["total", "tokenise", "unigrams", "count", "skipgrams", "prune", "export", "h2d", "index"]

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class ColibriError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (the C++ wrapper throws InternalError instead)."""

    def __init__(self, code, msg):
        super().__init__("colibri_b200 error %d: %s" % (code, msg))
        self.code = code


class COptions(C.Structure):
    """struct colibri_b200_options (field names = PatternModelOptions, reference include/patternmodel.h:103-180)."""

    _fields_ = [(n, C.c_int32) for n in (
        "MINTOKENS", "MINTOKENS_SKIPGRAMS", "MINTOKENS_UNIGRAMS", "MINLENGTH", "MAXLENGTH", "MAXBACKOFFLENGTH", "MINSKIPTYPES", "MAXSKIPS",
        "DOSKIPGRAMS", "DOSKIPGRAMS_EXHAUSTIVE", "DOPATTERNPERLINE", "PRUNENONSUBSUMED", "PRUNESUBSUMED", "QUIET", "DEBUG", "model_type", "streamed", "device",
        "DOREMOVEINDEX", "DOREMOVENGRAMS", "DOREMOVESKIPGRAMS", "DOREMOVEFLEXGRAMS", "DORESET")]


class CTrainSummary(C.Structure):
    """struct colibri_b200_train_summary (include/colibri_b200.h)."""

    _fields_ = [("npatterns", C.c_uint64), ("keybytes", C.c_uint64), ("totaltokens", C.c_uint64), ("totaltypes", C.c_uint64), ("maxn", C.c_int32), ("minn", C.c_int32),
                ("hasskipgrams", C.c_int32), ("npasses", C.c_int32), ("passes", (C.c_uint64 * 4) * 32), ("ms", C.c_double * 16), ("counters", C.c_uint64 * 8)]


class CSynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("ntokens", C.c_uint64), ("vocab", C.c_uint32), ("mean_sentence", C.c_uint32), ("phrase_permille", C.c_uint32), ("nphrases", C.c_uint32),
                ("first_token", C.c_uint64)]


_lib = None


def library():
    """Load libcolibri_b200.so (built in-tree by `make lib` / __graft_entry__.build()).  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make lib` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.colibri_b200_last_error.restype = C.c_char_p
    L.colibri_b200_version.restype = C.c_char_p
    L.colibri_b200_device_count.restype = C.c_int
    L.colibri_b200_options_default.argtypes = [C.POINTER(COptions)]
    L.colibri_b200_corpus_stage.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_corpus_from_device.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_corpus_bytes.argtypes = [C.c_void_p]
    L.colibri_b200_corpus_bytes.restype = C.c_size_t
    L.colibri_b200_corpus_free.argtypes = [C.c_void_p]
    L.colibri_b200_corpus_free.restype = None
    L.colibri_b200_corpus_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.colibri_b200_corpus_tokens.argtypes = [C.c_void_p, _u32p, C.c_uint64, _u64p]
    L.colibri_b200_train.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(COptions), C.POINTER(C.c_void_p)]
    L.colibri_b200_train_corpus.argtypes = [C.c_void_p, C.POINTER(COptions), C.POINTER(C.c_void_p)]
    L.colibri_b200_train_export.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(COptions), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(CTrainSummary)]
    L.colibri_b200_model_free.argtypes = [C.c_void_p]
    L.colibri_b200_model_free.restype = None
    for f in ("size", "tokens", "types"):
        fn = getattr(L, "colibri_b200_model_" + f)
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_uint64
    for f in ("maxn", "minn", "hasskipgrams", "type", "passes"):
        fn = getattr(L, "colibri_b200_model_" + f)
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_int
    L.colibri_b200_model_pass_stats.argtypes = [C.c_void_p, C.c_int, _u64p]
    L.colibri_b200_model_export_sizes.argtypes = [C.c_void_p, _u64p, _u64p, _u64p]
    L.colibri_b200_model_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.colibri_b200_model_export_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.colibri_b200_model_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.colibri_b200_model_lookup.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, _u32p]
    L.colibri_b200_model_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.colibri_b200_model_from_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                               C.POINTER(C.c_void_p)]
    L.colibri_b200_model_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(COptions), C.c_void_p, C.POINTER(C.c_void_p)]
    L.colibri_b200_train_constrained.argtypes = [C.c_void_p, C.POINTER(COptions), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_constrained_count.argtypes = [C.c_void_p, C.POINTER(COptions), C.c_void_p, C.c_void_p, _u64p, _u64p]
    L.colibri_b200_constrained_finish.argtypes = [C.POINTER(COptions), C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_model_flexgrams_fromskipgrams.argtypes = [C.c_void_p, _u64p, C.POINTER(C.c_void_p)]
    L.colibri_b200_model_hasflexgrams.argtypes = [C.c_void_p]
    L.colibri_b200_model_hasflexgrams.restype = C.c_int
    L.colibri_b200_model_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.colibri_b200_model_counters.argtypes = [C.c_void_p, _u64p]
    L.colibri_b200_model_level_counters.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.colibri_b200_model_level_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.colibri_b200_model_checksum.argtypes = [C.c_void_p, _u64p]
    L.colibri_b200_shard_begin.argtypes = [C.c_void_p, C.POINTER(COptions), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_shard_info.argtypes = [C.c_void_p, _u64p]
    L.colibri_b200_shard_device_ms.argtypes = [C.c_void_p]
    L.colibri_b200_shard_device_ms.restype = C.c_double
    L.colibri_b200_shard_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.colibri_b200_shard_unigram_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.colibri_b200_shard_unigram_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
    L.colibri_b200_shard_level_split_count.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p]
    L.colibri_b200_shard_level_split_write.argtypes = [C.c_void_p, C.c_void_p]
    L.colibri_b200_shard_level_owner.argtypes = [C.c_void_p, C.c_void_p, _u64p, C.c_void_p, _u64p, _u64p]
    L.colibri_b200_shard_level_owner_survivors.argtypes = [C.c_void_p, C.c_void_p]
    L.colibri_b200_shard_level_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p]
    L.colibri_b200_shard_skip_split_count.argtypes = [C.c_void_p, _u64p, _u64p]
    L.colibri_b200_shard_skip_split_write.argtypes = [C.c_void_p, C.c_void_p]
    L.colibri_b200_shard_skip_owner.argtypes = [C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p]
    L.colibri_b200_shard_skip_owner_survivors.argtypes = [C.c_void_p, C.c_void_p]
    L.colibri_b200_shard_skip_finish.argtypes = [C.c_void_p, C.c_void_p, _u64p]
    L.colibri_b200_shard_set_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.colibri_b200_shard_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.colibri_b200_shard_set_peers.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, C.c_uint64, C.c_uint64]
    L.colibri_b200_shard_p2p_split.argtypes = [C.c_void_p, C.c_int, _u64p]
    L.colibri_b200_shard_p2p_owner.argtypes = [C.c_void_p, _u64p]
    L.colibri_b200_shard_p2p_finish.argtypes = [C.c_void_p, _u64p, _u64p]
    L.colibri_b200_shard_finish.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_shard_free.argtypes = [C.c_void_p]
    L.colibri_b200_shard_free.restype = None
    L.colibri_b200_train_multi.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(COptions), C.c_void_p, C.c_int, C.c_void_p]
    L.colibri_b200_rindex_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.colibri_b200_rindex_free.argtypes = [C.c_void_p]
    L.colibri_b200_rindex_free.restype = None
    L.colibri_b200_rindex_info.argtypes = [C.c_void_p, _u64p]
    L.colibri_b200_rindex_lengths.argtypes = [C.c_void_p, _u32p, C.c_uint32, _u32p]
    L.colibri_b200_rindex_sentence_starts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.colibri_b200_rindex_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.colibri_b200_rindex_cooc.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
    L.colibri_b200_rindex_cooc_of.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
    L.colibri_b200_hash64_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.colibri_b200_synth_corpus.argtypes = [C.POINTER(CSynthParams), C.c_int, C.POINTER(C.c_void_p)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise ColibriError(rc, library().colibri_b200_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(library().colibri_b200_device_count())


def declared_symbols():
    """Every function name declared in include/colibri_b200.h (used by the CPU test that the library exports them all)."""
    import re

    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(colibri_b200_[a-z0-9_]+)\s*\(", text)))


class PatternModelOptions:
    """Same public fields and defaults as the reference's PatternModelOptions (include/patternmodel.h:103-180)."""

    def __init__(self, **kw):
        self._c = COptions()
        library().colibri_b200_options_default(C.byref(self._c))
        for k, v in kw.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if k == "_c":
            object.__setattr__(self, k, v)
        elif k in dict(COptions._fields_):
            setattr(self._c, k, int(v))
        else:
            raise AttributeError("No such option: " + k)

    def __getattr__(self, k):
        if k in dict(COptions._fields_):
            return getattr(self._c, k)
        raise AttributeError(k)


def _as_u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


class Corpus:
    """A .colibri.dat body staged in HBM (colibri_b200_corpus)."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device

    @classmethod
    def from_bytes(cls, body, device=0):
        a = _as_u8(body)
        h = C.c_void_p()
        _check(library().colibri_b200_corpus_stage(a.ctypes.data if a.size else None, a.size, device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_host_pointer(cls, ptr, nbytes, device=0):
        h = C.c_void_p()
        _check(library().colibri_b200_corpus_stage(ptr, nbytes, device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_device_pointer(cls, ptr, nbytes, device=0):
        h = C.c_void_p()
        _check(library().colibri_b200_corpus_from_device(ptr, nbytes, device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_file(cls, path, device=0):
        data = np.fromfile(path, dtype=np.uint8)
        if data.size < 2 or data[0] != 0xA2 or data[1] != 2:
            raise ColibriError(4, "not a .colibri.dat v2 file (expected 0xA2 0x02 header): " + path)
        return cls.from_bytes(data[2:], device)

    @classmethod
    def synthetic(cls, ntokens, vocab=100000, seed=1, mean_sentence=22, phrase_permille=0, nphrases=0, device=0, first_token=0):
        p = CSynthParams(seed, ntokens, vocab, mean_sentence, phrase_permille, nphrases, first_token)
        h = C.c_void_p()
        _check(library().colibri_b200_synth_corpus(C.byref(p), device, C.byref(h)))
        return cls(h, device)

    @property
    def nbytes(self):
        return int(library().colibri_b200_corpus_bytes(self._h))

    def download(self) -> np.ndarray:
        out = np.empty(self.nbytes, dtype=np.uint8)
        _check(library().colibri_b200_corpus_download(self._h, out.ctypes.data, out.size))
        return out

    def tokens(self) -> np.ndarray:
        """Class id per position, 0 = sentence delimiter (device-side bytestoint)."""
        n = C.c_uint64()
        _check(library().colibri_b200_corpus_tokens(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint32)
        _check(library().colibri_b200_corpus_tokens(self._h, out.ctypes.data_as(_u32p), out.size, C.byref(n)))
        return out

    def close(self):
        if self._h:
            library().colibri_b200_corpus_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """A trained pattern model (colibri_b200_model): header numbers on the host, patterns resident on the device."""

    def __init__(self, handle):
        self._h = handle
        self._flat = None

    def __len__(self):
        return int(library().colibri_b200_model_size(self._h))

    def tokens(self):
        return int(library().colibri_b200_model_tokens(self._h))

    def types(self):
        return int(library().colibri_b200_model_types(self._h))

    def maxlength(self):
        return int(library().colibri_b200_model_maxn(self._h))

    def minlength(self):
        return int(library().colibri_b200_model_minn(self._h))

    @property
    def hasskipgrams(self):
        return bool(library().colibri_b200_model_hasskipgrams(self._h))

    @property
    def hasflexgrams(self):
        return bool(library().colibri_b200_model_hasflexgrams(self._h))

    def flexgrams_fromskipgrams(self):
        """IndexedPatternModel::computeflexgrams_fromskipgrams: (number of new flexgrams, a new Model = this one + the flexgrams)."""
        found = C.c_uint64()
        h = C.c_void_p()
        _check(library().colibri_b200_model_flexgrams_fromskipgrams(self._h, C.byref(found), C.byref(h)))
        return int(found.value), Model(h)

    @property
    def model_type(self):
        return int(library().colibri_b200_model_type(self._h))

    def passes(self):
        """[(n, found n-grams, found skipgrams, pruned)] -- the numbers of the reference's progress lines."""
        out = []
        st = (C.c_uint64 * 4)()
        for p in range(library().colibri_b200_model_passes(self._h)):
            _check(library().colibri_b200_model_pass_stats(self._h, p, st))
            out.append(tuple(int(x) for x in st))
        return out

    def export(self, pinned=None):
        """(keys uint8 blob, key_off uint64[n+1], counts uint32[n], refs or None) copied to host memory."""
        if self._flat is not None and pinned is None:
            return self._flat
        n, kb, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(library().colibri_b200_model_export_sizes(self._h, C.byref(n), C.byref(kb), C.byref(nr)))
        if pinned is not None:
            keys, key_off, counts = pinned(kb.value, n.value)
        else:
            keys = np.empty(kb.value, dtype=np.uint8)
            key_off = np.empty(n.value + 1, dtype=np.uint64)
            counts = np.empty(n.value, dtype=np.uint32)
        rs = rt = ro = None
        if self.model_type == INDEXEDPATTERNMODEL:
            rs = np.empty(nr.value, dtype=np.uint32)
            rt = np.empty(nr.value, dtype=np.uint16)
            ro = np.empty(n.value + 1, dtype=np.uint64)
        _check(library().colibri_b200_model_export(self._h, keys.ctypes.data, key_off.ctypes.data, counts.ctypes.data, rs.ctypes.data if rs is not None else None,
                                                    rt.ctypes.data if rt is not None else None, ro.ctypes.data if ro is not None else None))
        flat = (keys, key_off, counts, (rs, rt, ro) if rs is not None else None)
        if pinned is None:
            self._flat = flat
        return flat

    def export_sizes(self):
        n, kb, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(library().colibri_b200_model_export_sizes(self._h, C.byref(n), C.byref(kb), C.byref(nr)))
        return n.value, kb.value, nr.value

    def export_into(self, keys_ptr, off_ptr, counts_ptr):
        """Export into caller-owned (e.g. pinned) host buffers given as raw addresses."""
        _check(library().colibri_b200_model_export(self._h, keys_ptr, off_ptr, counts_ptr, None, None, None))

    def export_compact(self):
        """(keys blob, key_len uint16[n], counts uint32[n]): the flat model without the 8-byte offsets."""
        n, kb, _ = self.export_sizes()
        keys, lens, counts = np.empty(kb, dtype=np.uint8), np.empty(n, dtype=np.uint16), np.empty(n, dtype=np.uint32)
        _check(library().colibri_b200_model_export_compact(self._h, keys.ctypes.data, lens.ctypes.data, counts.ctypes.data))
        return keys, lens, counts

    def export_compact_into(self, keys_ptr, len_ptr, counts_ptr):
        _check(library().colibri_b200_model_export_compact(self._h, keys_ptr, len_ptr, counts_ptr))

    def to_bytes(self) -> bytes:
        """The .colibri.patternmodel byte stream (reference include/patternmodel.h:1609-1624)."""
        need = C.c_size_t()
        _check(library().colibri_b200_model_write(self._h, None, 0, C.byref(need)))
        buf = np.empty(need.value, dtype=np.uint8)
        _check(library().colibri_b200_model_write(self._h, buf.ctypes.data, buf.size, C.byref(need)))
        return buf.tobytes()

    def write(self, filename):
        with open(filename, "wb") as f:
            f.write(self.to_bytes())

    def occurrencecount(self, key: bytes) -> int:
        c = C.c_uint32()
        _check(library().colibri_b200_model_lookup(self._h, key, len(key), C.byref(c)))
        return int(c.value)

    def has(self, key: bytes) -> bool:
        return self.lookup_batch([key])[1][0] >= 0

    def lookup_batch(self, keys):
        """(counts uint32[n], index int64[n]) of n patterns given as byte strings: occurrencecount()/has() on the device.
        index = position in the export order, -1 when the pattern is not in the model."""
        blob = np.frombuffer(b"".join(keys), dtype=np.uint8) if keys else np.zeros(0, dtype=np.uint8)
        off = np.zeros(len(keys) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(k) for k in keys])
        counts = np.zeros(len(keys), dtype=np.uint32)
        index = np.full(len(keys), -1, dtype=np.int64)
        _check(library().colibri_b200_model_lookup_batch(self._h, blob.ctypes.data if blob.size else None, off.ctypes.data, len(keys), counts.ctypes.data, index.ctypes.data))
        return counts, index

    @classmethod
    def from_flat(cls, keys, key_off, counts=None, refs=None, tokens=0, types=0, model_type=UNINDEXEDPATTERNMODEL, device=0) -> "Model":
        """Upload a pattern set given as flat host arrays (the export form)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
        n = len(key_off) - 1
        counts = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
        rs = rt = ro = None
        if refs is not None:
            rs, rt, ro = (np.ascontiguousarray(refs[0], dtype=np.uint32), np.ascontiguousarray(refs[1], dtype=np.uint16), np.ascontiguousarray(refs[2], dtype=np.uint64))
        h = C.c_void_p()
        _check(library().colibri_b200_model_from_flat(keys.ctypes.data if keys.size else None, key_off.ctypes.data, counts.ctypes.data if counts is not None else None, n,
                                                       rs.ctypes.data if rs is not None and rs.size else None, rt.ctypes.data if rt is not None and rt.size else None,
                                                       ro.ctypes.data if ro is not None else None, int(tokens), int(types), int(model_type), int(device), C.byref(h)))
        return cls(h)

    def timings(self):
        ms = (C.c_double * T_NPHASES)()
        _check(library().colibri_b200_model_timings(self._h, ms))
        return {PHASE_NAMES[i]: float(ms[i]) for i in range(T_NPHASES - 1)}

    def counters(self):
        c = (C.c_uint64 * 8)()
        _check(library().colibri_b200_model_counters(self._h, c))
        names = ["positions", "corpus_bytes", "kernel_launches", "ngram_upserts", "skipgram_upserts", "slots_initialised", "unigram_increments", "peak_device_bytes"]
        return {k: int(v) for k, v in zip(names, c)}

    def checksum(self):
        """Order-independent checksum of the model content (see include/colibri_b200.h): dict of sum, xor, occurrences, patterns, refsum, refs."""
        out = (C.c_uint64 * 6)()
        _check(library().colibri_b200_model_checksum(self._h, out))
        return dict(zip(("sum", "xor", "occurrences", "patterns", "refsum", "refs"), (int(x) for x in out)))

    def level(self, n):
        out = (C.c_double * 8)()
        _check(library().colibri_b200_model_level_info(self._h, n, out))
        return {"windows": int(out[0]), "capacity": int(out[1]), "count_ms": float(out[2]), "singletons": int(out[3]), "items": int(out[4]),
                "path": "partitioned" if out[5] else "table", "filtered": bool(out[6]), "fused_id1": bool(out[7])}

    def close(self):
        if self._h:
            library().colibri_b200_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def train_multi(body, devices, options: PatternModelOptions | None = None, **kw):
    """colibri_b200_train_multi: one process, one host thread per GPU.  Returns the list of the devices' shares (Model objects)."""
    o = options if options is not None else PatternModelOptions(**kw)
    a = np.ascontiguousarray(np.frombuffer(bytes(body), dtype=np.uint8) if not isinstance(body, np.ndarray) else body, dtype=np.uint8)
    devs = (C.c_int * len(devices))(*devices)
    outs = (C.c_void_p * len(devices))()
    _check(library().colibri_b200_train_multi(a.ctypes.data if a.size else None, a.size, C.byref(o._c), devs, len(devices), outs))
    return [Model(C.c_void_p(h)) for h in outs]


class ReverseIndex:
    """colibri_b200_rindex: a model's n-grams matched against every position of a corpus once; answers getreverseindex in batches and the
    co-occurrence relations of all patterns (reference include/patternmodel.h:1746-1824, :3460-3531)."""

    def __init__(self, model: Model, corpus: "Corpus", streamed: int = 0):
        self._h = C.c_void_p()
        self.model, self.corpus = model, corpus  # keep them alive
        _check(library().colibri_b200_rindex_build(model._h, corpus._h, int(streamed), C.byref(self._h)))
        out = (C.c_uint64 * 4)()
        _check(library().colibri_b200_rindex_info(self._h, out))
        self.sentences, self.positions = int(out[0]), int(out[1])
        n = C.c_uint32()
        buf = (C.c_uint32 * 256)()
        _check(library().colibri_b200_rindex_lengths(self._h, buf, 256, C.byref(n)))
        self.lengths = [int(buf[i]) for i in range(n.value)]

    def sentence_starts(self) -> np.ndarray:
        out = np.empty(self.sentences + 1, dtype=np.uint32)
        _check(library().colibri_b200_rindex_sentence_starts(self._h, out.ctypes.data, out.size))
        return out

    def query(self, refs) -> np.ndarray:
        """refs: iterable of (sentence, token), sentences from 1.  Returns u32[len(refs), len(self.lengths)]: pattern index + 1 in the model's export order, 0 = none."""
        refs = list(refs)
        s = np.ascontiguousarray([r[0] for r in refs], dtype=np.uint32)
        t = np.ascontiguousarray([r[1] for r in refs], dtype=np.uint16)
        out = np.zeros((len(refs), max(len(self.lengths), 1)), dtype=np.uint32)
        if len(refs) and self.lengths:
            _check(library().colibri_b200_rindex_query(self._h, s.ctypes.data, t.ctypes.data, len(refs), out.ctypes.data))
        return out[:, :len(self.lengths)]

    def cooc(self, left: bool = False):
        """(index of P, index of Q, joint) of every getrightcooc / getleftcooc relation, as three arrays."""
        n = C.c_uint64()
        _check(library().colibri_b200_rindex_cooc(self._h, 1 if left else 0, None, None, None, 0, C.byref(n)))
        p, q, j = np.empty(n.value, dtype=np.uint32), np.empty(n.value, dtype=np.uint32), np.empty(n.value, dtype=np.uint64)
        if n.value:
            _check(library().colibri_b200_rindex_cooc(self._h, 1 if left else 0, p.ctypes.data, q.ctypes.data, j.ctypes.data, n.value, C.byref(n)))
        return p[:n.value], q[:n.value], j[:n.value]

    def cooc_of(self, pattern: int):
        """getcooc of the pattern with export index `pattern` (both directions, no overlap): (index of Q, count) as two arrays."""
        n = C.c_uint64()
        _check(library().colibri_b200_rindex_cooc_of(self._h, pattern, None, None, 0, C.byref(n)))
        q, c = np.empty(n.value, dtype=np.uint32), np.empty(n.value, dtype=np.uint64)
        if n.value:
            _check(library().colibri_b200_rindex_cooc_of(self._h, pattern, q.ctypes.data, c.ctypes.data, n.value, C.byref(n)))
        return q[:n.value], c[:n.value]

    def close(self):
        if self._h:
            library().colibri_b200_rindex_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def train(corpus, options: PatternModelOptions | None = None, **kw) -> Model:
    """PatternModel::train on the device.  `corpus` is a staged Corpus, or host bytes / uint8 array of a .colibri.dat body
    (then the host->device copy is part of the call, like colibri_b200_train)."""
    if options is None:
        options = PatternModelOptions(**kw)
    elif kw:
        raise TypeError("pass either options or keyword fields")
    h = C.c_void_p()
    if isinstance(corpus, Corpus):
        _check(library().colibri_b200_train_corpus(corpus._h, C.byref(options._c), C.byref(h)))
    else:
        a = _as_u8(corpus)
        _check(library().colibri_b200_train(a.ctypes.data if a.size else None, a.size, C.byref(options._c), C.byref(h)))
    return Model(h)


def load_model(blob, options: PatternModelOptions | None = None, constrain: Model | None = None, **kw) -> Model:
    """PatternModel::load: the bytes of a .colibri.patternmodel file read as options.model_type, the options acting as filters
    (MINTOKENS, MINLENGTH, MAXLENGTH, DOREMOVE*, DORESET) and `constrain` as the membership constraint."""
    if options is None:
        options = PatternModelOptions(**kw)
    elif kw:
        raise TypeError("pass either options or keyword fields")
    a = _as_u8(blob)
    h = C.c_void_p()
    _check(library().colibri_b200_model_load(a.ctypes.data if a.size else None, a.size, C.byref(options._c), constrain._h if constrain is not None else None, C.byref(h)))
    return Model(h)


def train_constrained(corpus, constrain: Model, inplace: bool = False, options: PatternModelOptions | None = None, **kw) -> Model:
    """PatternModel::train(corpus, options, constrainbymodel): only patterns of `constrain` are counted.  inplace=True is the
    reference's constrainbymodel == this (CLI -I, stage 2 of -2): `constrain` is the model being rebuilt, loaded with DORESET."""
    if options is None:
        options = PatternModelOptions(**kw)
    elif kw:
        raise TypeError("pass either options or keyword fields")
    own = None
    if not isinstance(corpus, Corpus):
        own = corpus = Corpus.from_bytes(corpus, device=options.device)
    h = C.c_void_p()
    try:
        _check(library().colibri_b200_train_constrained(corpus._h, C.byref(options._c), constrain._h, 1 if inplace else 0, C.byref(h)))
    finally:
        if own is not None:
            own.close()
    return Model(h)


def constrained_count(shard: Corpus, constrain: Model, options: PatternModelOptions, counts_ptr):
    """One rank's share of a sharded constrained run: add this shard's occurrences of every pattern of `constrain` to the device array
    at counts_ptr (uint32[len(constrain)], zeroed by the caller).  Returns (tokens of the shard, kernel launches)."""
    tokens, launches = C.c_uint64(), C.c_uint64()
    _check(library().colibri_b200_constrained_count(shard._h, C.byref(options._c), constrain._h, counts_ptr, C.byref(tokens), C.byref(launches)))
    return int(tokens.value), int(launches.value)


def constrained_finish(constrain: Model, options: PatternModelOptions, counts_ptr, corpus_tokens: int, inplace: bool = False) -> Model:
    """Threshold + compaction of the counters summed over all shards (see constrained_count)."""
    h = C.c_void_p()
    _check(library().colibri_b200_constrained_finish(C.byref(options._c), constrain._h, counts_ptr, int(corpus_tokens), 1 if inplace else 0, C.byref(h)))
    return Model(h)


def train_host_pointer(ptr, nbytes, options: PatternModelOptions) -> Model:
    """colibri_b200_train on a raw host address (e.g. a pinned torch tensor's data_ptr())."""
    h = C.c_void_p()
    _check(library().colibri_b200_train(ptr, nbytes, C.byref(options._c), C.byref(h)))
    return Model(h)


def train_export_pointers(ptr, nbytes, options: PatternModelOptions, keys_ptr, keys_cap, len_ptr, counts_ptr, patterns_cap) -> dict:
    """colibri_b200_train_export on raw host addresses: corpus in, compact flat model out (keys blob, u16 key lengths, u32 counts), with the
    copies of finished levels overlapped with the counting of the next.  Returns the summary as a dict."""
    sm = CTrainSummary()
    _check(library().colibri_b200_train_export(ptr, nbytes, C.byref(options._c), keys_ptr, keys_cap, len_ptr, counts_ptr, patterns_cap, C.byref(sm)))
    return {"npatterns": int(sm.npatterns), "keybytes": int(sm.keybytes), "tokens": int(sm.totaltokens), "types": int(sm.totaltypes), "maxn": sm.maxn, "minn": sm.minn,
            "hasskipgrams": bool(sm.hasskipgrams), "passes": [tuple(int(x) for x in sm.passes[i]) for i in range(min(sm.npasses, 32))],
            "ms": {PHASE_NAMES[i]: float(sm.ms[i]) for i in range(T_NPHASES)}, "kernel_launches": int(sm.counters[2])}


def train_export(body, options: PatternModelOptions | None = None, **kw):
    """train + compact export in one call (numpy in / numpy out).  Returns (keys, key_len, counts, summary)."""
    o = options if options is not None else PatternModelOptions(**kw)
    a = np.ascontiguousarray(np.frombuffer(bytes(body), dtype=np.uint8) if not isinstance(body, np.ndarray) else body, dtype=np.uint8)
    cap_p, cap_k = 1 << 16, 1 << 20
    while True:
        keys, lens, counts = np.empty(cap_k, dtype=np.uint8), np.empty(cap_p, dtype=np.uint16), np.empty(cap_p, dtype=np.uint32)
        sm = CTrainSummary()
        rc = library().colibri_b200_train_export(a.ctypes.data if a.size else None, a.size, C.byref(o._c), keys.ctypes.data, cap_k, lens.ctypes.data, counts.ctypes.data, cap_p, C.byref(sm))
        if rc == 5 and (sm.npatterns > cap_p or sm.keybytes > cap_k):  # COLIBRI_E_CAPACITY: the summary says what is needed
            cap_p, cap_k = max(cap_p, int(sm.npatterns) + 16), max(cap_k, int(sm.keybytes) + 16)
            continue
        _check(rc)
        n, kb = int(sm.npatterns), int(sm.keybytes)
        summary = {"tokens": int(sm.totaltokens), "types": int(sm.totaltypes), "maxn": sm.maxn, "minn": sm.minn, "hasskipgrams": bool(sm.hasskipgrams),
                   "passes": [tuple(int(x) for x in sm.passes[i]) for i in range(min(sm.npasses, 32))]}
        return keys[:kb], lens[:n], counts[:n], summary


def hash64_batch(keys, device=0) -> np.ndarray:
    """Pattern::hash (SpookyHash::Hash64, seed 0) of each byte string, computed on the device."""
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8) if keys else np.zeros(0, dtype=np.uint8)
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(k) for k in keys])
    out = np.zeros(len(keys), dtype=np.uint64)
    _check(library().colibri_b200_hash64_batch(blob.ctypes.data if blob.size else None, off.ctypes.data, len(keys), out.ctypes.data, device))
    return out
