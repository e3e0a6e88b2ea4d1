"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of ``oracle/liboracle.so`` (our plain-C restatement of the
reference's PatternModel::train path, see oracle.h) plus a runner for the
unmodified reference compiled into ``oracle/_ref/`` (``make -C oracle ref``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only
as the checker.  Parity status of the restatement: PINNED (see oracle.h).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_TRAIN = os.path.join(REF_DIR, "ref_train")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class Options(C.Structure):
    """POD mirror of oracle_options (include/patternmodel.h:103-180 of the reference)."""

    _fields_ = [
        ("mintokens", C.c_int32),
        ("mintokens_skipgrams", C.c_int32),
        ("mintokens_unigrams", C.c_int32),
        ("minlength", C.c_int32),
        ("maxlength", C.c_int32),
        ("maxbackofflength", C.c_int32),
        ("minskiptypes", C.c_int32),
        ("maxskips", C.c_int32),
        ("doskipgrams", C.c_int32),
        ("doskipgrams_exhaustive", C.c_int32),
        ("indexed", C.c_int32),
        ("streamed", C.c_int32),
    ]


class LoadOptions(C.Structure):
    """POD mirror of oracle_load_options: the PatternModelOptions fields load() uses as filters."""

    _fields_ = [
        ("mintokens", C.c_int32),
        ("minlength", C.c_int32),
        ("maxlength", C.c_int32),
        ("dongrams", C.c_int32),
        ("doskipgrams", C.c_int32),
        ("doflexgrams", C.c_int32),
        ("doreset", C.c_int32),
        ("load_indexed", C.c_int32),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("ntokens", C.c_uint64),
        ("vocab", C.c_uint32),
        ("mean_sentence", C.c_uint32),
        ("phrase_permille", C.c_uint32),
        ("nphrases", C.c_uint32),
        ("first_token", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc, a second or two).  Building the checker is not using it."""
    src = [os.path.join(HERE, "oracle.c"), os.path.join(HERE, "oracle.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if stale:
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return LIB_PATH


def build_ref() -> bool:
    """Compile the unmodified reference into oracle/_ref/ when /root/reference is mounted (this container only)."""
    if not os.path.isdir("/root/reference/src"):
        return os.path.exists(REF_TRAIN)
    subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)
    return os.path.exists(REF_TRAIN)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    L.oracle_options_default.argtypes = [C.POINTER(Options)]
    L.oracle_train.argtypes = [_u8p, C.c_size_t, C.POINTER(Options), C.POINTER(C.c_void_p)]
    L.oracle_train.restype = C.c_int
    L.oracle_model_free.argtypes = [C.c_void_p]
    L.oracle_last_error.restype = C.c_char_p
    for f in ("size", "tokens", "types"):
        fn = getattr(L, "oracle_model_" + f)
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_uint64
    for f in ("maxn", "minn", "hasskipgrams", "passes"):
        fn = getattr(L, "oracle_model_" + f)
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_int
    L.oracle_model_pass_stats.argtypes = [C.c_void_p, C.c_int, _u64p]
    L.oracle_model_count.argtypes = [C.c_void_p, _u8p, C.c_uint32]
    L.oracle_model_count.restype = C.c_uint32
    L.oracle_model_export_sizes.argtypes = [C.c_void_p, _u64p, _u64p, _u64p]
    L.oracle_model_export.argtypes = [C.c_void_p, _u8p, _u64p, _u32p, _u32p, _u16p, _u64p]
    L.oracle_model_write.argtypes = [C.c_void_p, _u8p, C.c_size_t]
    L.oracle_model_write.restype = C.c_size_t
    L.oracle_modelfile_scan.argtypes = [_u8p, C.c_size_t, _u64p]
    L.oracle_modelfile_parse.argtypes = [_u8p, C.c_size_t, _u8p, _u64p, _u32p, _u32p, _u16p, _u64p]
    L.oracle_model_load.argtypes = [_u8p, C.c_size_t, C.POINTER(LoadOptions), C.c_void_p, C.POINTER(C.c_void_p)]
    L.oracle_model_load.restype = C.c_int
    L.oracle_model_from_keys.argtypes = [_u8p, _u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
    L.oracle_model_from_keys.restype = C.c_int
    L.oracle_train_constrained.argtypes = [_u8p, C.c_size_t, C.POINTER(Options), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.oracle_train_constrained.restype = C.c_int
    L.oracle_computeflexgrams_fromskipgrams.argtypes = [C.c_void_p]
    L.oracle_computeflexgrams_fromskipgrams.restype = C.c_int64
    L.oracle_inttobytes.argtypes = [_u8p, C.c_uint32]
    L.oracle_inttobytes.restype = C.c_uint
    L.oracle_bytestoint.argtypes = [_u8p, C.POINTER(C.c_uint)]
    L.oracle_bytestoint.restype = C.c_uint32
    L.oracle_spooky_hash64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
    L.oracle_spooky_hash64.restype = C.c_uint64
    L.oracle_pattern_hash.argtypes = [_u8p, C.c_size_t]
    L.oracle_pattern_hash.restype = C.c_uint64
    L.oracle_skip_configurations.argtypes = [C.c_int, C.c_int, _u32p, C.c_int]
    L.oracle_skip_configurations.restype = C.c_int
    L.oracle_skipgram_collapse.argtypes = [_u8p, C.c_size_t, C.c_uint32, _u8p]
    L.oracle_skipgram_collapse.restype = C.c_size_t
    L.oracle_synth_corpus.argtypes = [C.POINTER(SynthParams), _u8p, C.c_size_t]
    L.oracle_synth_corpus.restype = C.c_size_t
    _lib = L
    return L


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


# --------------------------------------------------------------------------- canonical model form
@dataclass
class FlatModel:
    """Order-insensitive comparison form of a pattern model: parallel arrays, canonical = sorted by key bytes."""

    keys: np.ndarray  # uint8 blob
    key_off: np.ndarray  # uint64, npatterns+1
    counts: np.ndarray  # uint32
    tokens: int = 0
    types: int = 0
    maxn: int = 0
    minn: int = 999
    hasskipgrams: bool = False
    model_type: int = 10
    ref_sentence: np.ndarray | None = None
    ref_token: np.ndarray | None = None
    ref_off: np.ndarray | None = None
    passes: list = field(default_factory=list)  # [(n, found_ngrams, found_skipgrams, pruned)]
    flexfound: int = 0  # new flexgrams of computeflexgrams_fromskipgrams (train(..., flexfromskip=1))

    def __len__(self):
        return len(self.counts)

    def key(self, i: int) -> bytes:
        return self.keys[int(self.key_off[i]) : int(self.key_off[i + 1])].tobytes()

    def as_dict(self) -> dict:
        return {self.key(i): int(self.counts[i]) for i in range(len(self))}

    def refs(self, i: int):
        a, b = int(self.ref_off[i]), int(self.ref_off[i + 1])
        return list(zip(self.ref_sentence[a:b].tolist(), self.ref_token[a:b].tolist()))

    def padded_keys(self, width: int | None = None) -> np.ndarray:
        """Keys as a fixed-width 'S' array (keys never contain 0x00, so zero padding preserves memcmp order)."""
        off = self.key_off.astype(np.int64)
        lens = np.diff(off)
        w = int(lens.max()) if len(lens) else 1
        if width is not None:
            w = max(w, width)
        out = np.zeros((len(lens), w), dtype=np.uint8)
        if len(lens):
            rows = np.repeat(np.arange(len(lens)), lens)
            cols = np.arange(int(off[-1])) - np.repeat(off[:-1], lens)
            out[rows, cols] = self.keys[: int(off[-1])]
        return out.view("S%d" % w).reshape(-1)

    def canonical(self) -> "FlatModel":
        """Return a copy whose patterns are sorted bytewise (and whose per-pattern refs stay in stored order)."""
        pk = self.padded_keys()
        order = np.argsort(pk, kind="stable")
        off = self.key_off.astype(np.int64)
        lens = np.diff(off)
        new_lens = lens[order]
        new_off = np.zeros(len(order) + 1, dtype=np.uint64)
        new_off[1:] = np.cumsum(new_lens)
        if len(order):
            src = np.repeat(off[:-1][order], new_lens) + (np.arange(int(new_lens.sum())) - np.repeat(new_off[:-1].astype(np.int64), new_lens))
            new_keys = self.keys[src]
        else:
            new_keys = self.keys[:0]
        rs = rt = ro = None
        if self.ref_off is not None:
            roff = self.ref_off.astype(np.int64)
            rl = np.diff(roff)[order]
            ro = np.zeros(len(order) + 1, dtype=np.uint64)
            ro[1:] = np.cumsum(rl)
            if len(order) and rl.sum():
                rsrc = np.repeat(roff[:-1][order], rl) + (np.arange(int(rl.sum())) - np.repeat(ro[:-1].astype(np.int64), rl))
                rs, rt = self.ref_sentence[rsrc], self.ref_token[rsrc]
            else:
                rs, rt = self.ref_sentence[:0], self.ref_token[:0]
        return FlatModel(new_keys, new_off, self.counts[order], self.tokens, self.types, self.maxn, self.minn, self.hasskipgrams, self.model_type, rs, rt, ro,
                         list(self.passes))

    def sorted_refs(self) -> "FlatModel":
        """Canonical copy whose occurrence lists are ascending inside every pattern (flexgram lists of the reference are in insertion order)."""
        c = self.canonical()
        if c.ref_off is None or len(c) == 0:
            return c
        ro = c.ref_off.astype(np.int64)
        owner = np.repeat(np.arange(len(c)), np.diff(ro))
        order = np.lexsort((c.ref_token, c.ref_sentence, owner))
        return FlatModel(c.keys, c.key_off, c.counts, c.tokens, c.types, c.maxn, c.minn, c.hasskipgrams, c.model_type, c.ref_sentence[order], c.ref_token[order], c.ref_off,
                         list(c.passes), c.flexfound)

    def same_patterns(self, other: "FlatModel") -> bool:
        a, b = self.canonical(), other.canonical()
        ok = len(a) == len(b) and np.array_equal(a.key_off, b.key_off) and np.array_equal(a.keys[: int(a.key_off[-1])], b.keys[: int(b.key_off[-1])]) and np.array_equal(a.counts, b.counts)
        if ok and (a.ref_off is not None or b.ref_off is not None):
            ok = a.ref_off is not None and b.ref_off is not None and np.array_equal(a.ref_off, b.ref_off) and np.array_equal(a.ref_sentence, b.ref_sentence) and np.array_equal(a.ref_token, b.ref_token)
        return bool(ok)

    def digest(self) -> str:
        """sha256 over the canonical (key, count[, refs]) stream -- what tests/golden/*.json pin."""
        import hashlib

        c = self.canonical()
        h = hashlib.sha256()
        h.update(np.uint64(len(c)).tobytes())
        h.update(c.key_off.astype(np.uint64).tobytes())
        h.update(c.keys[: int(c.key_off[-1])].tobytes())
        h.update(c.counts.astype(np.uint32).tobytes())
        if c.ref_off is not None:
            h.update(c.ref_off.astype(np.uint64).tobytes())
            h.update(c.ref_sentence.astype(np.uint32).tobytes())
            h.update(c.ref_token.astype(np.uint16).tobytes())
        return h.hexdigest()


def default_options(**kw) -> Options:
    o = Options()
    lib().oracle_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, int(v))
    return o


def _flat_from_handle(h, indexed: bool) -> FlatModel:
    L = lib()
    np_, kb, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.oracle_model_export_sizes(h, C.byref(np_), C.byref(kb), C.byref(nr))
    keys = np.zeros(kb.value + 1, dtype=np.uint8)
    key_off = np.zeros(np_.value + 1, dtype=np.uint64)
    counts = np.zeros(np_.value, dtype=np.uint32)
    rs = rt = ro = None
    if indexed:
        rs = np.zeros(nr.value, dtype=np.uint32)
        rt = np.zeros(nr.value, dtype=np.uint16)
        ro = np.zeros(np_.value + 1, dtype=np.uint64)
    L.oracle_model_export(h, _ptr(keys, _u8p), _ptr(key_off, _u64p), _ptr(counts, _u32p), _ptr(rs, _u32p) if indexed else None, _ptr(rt, _u16p) if indexed else None,
                          _ptr(ro, _u64p) if indexed else None)
    passes = []
    st = (C.c_uint64 * 4)()
    for p in range(L.oracle_model_passes(h)):
        L.oracle_model_pass_stats(h, p, st)
        passes.append(tuple(int(x) for x in st))
    return FlatModel(keys[: kb.value], key_off, counts, int(L.oracle_model_tokens(h)), int(L.oracle_model_types(h)), L.oracle_model_maxn(h), L.oracle_model_minn(h),
                     bool(L.oracle_model_hasskipgrams(h)), 20 if indexed else 10, rs, rt, ro, passes)


def train(corpus, flexfromskip=0, **kw) -> FlatModel:
    """Run the C restatement on the body of a .colibri.dat (bytes after the 2-byte header).  flexfromskip: follow up with
    computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744), as the CLI's `-F S` does."""
    L = lib()
    data = _as_u8(corpus)
    o = default_options(**kw)
    h = C.c_void_p()
    rc = L.oracle_train(_ptr(data, _u8p), data.size, C.byref(o), C.byref(h))
    if rc != 0:
        raise RuntimeError("oracle_train: " + L.oracle_last_error().decode())
    try:
        found = 0
        if flexfromskip:
            found = int(L.oracle_computeflexgrams_fromskipgrams(h))
            if found < 0:
                raise RuntimeError("oracle_computeflexgrams_fromskipgrams: " + L.oracle_last_error().decode())
        fm = _flat_from_handle(h, bool(o.indexed))
        fm.flexfound = found
        return fm
    finally:
        L.oracle_model_free(h)


class LoadedModel:
    """A model held inside the oracle library (result of load_model / model_from_flat): the constraint side of train_constrained."""

    def __init__(self, handle, indexed: bool):
        self.h = handle
        self.indexed = indexed

    def flat(self) -> FlatModel:
        return _flat_from_handle(self.h, self.indexed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_model_free(self.h)
            self.h = None


def load_model(blob, mintokens=-1, minlength=1, maxlength=100, dongrams=1, doskipgrams=1, doflexgrams=1, doreset=0, indexed=0, constrain: "LoadedModel | None" = None) -> LoadedModel:
    """PatternModel::load with the options as filters (include/patternmodel.h:781-861, include/patternstore.h:555-619)."""
    L = lib()
    data = _as_u8(blob)
    lo = LoadOptions(int(mintokens), int(minlength), int(maxlength), int(dongrams), int(doskipgrams), int(doflexgrams), int(doreset), int(indexed))
    h = C.c_void_p()
    if L.oracle_model_load(_ptr(data, _u8p), data.size, C.byref(lo), constrain.h if constrain is not None else None, C.byref(h)) != 0:
        raise RuntimeError("oracle_model_load: " + L.oracle_last_error().decode())
    return LoadedModel(h, bool(indexed))


def model_from_flat(fm: FlatModel, indexed=0) -> LoadedModel:
    """A zero-count model holding fm's patterns and totals (constraint side when no file is at hand)."""
    L = lib()
    keys = np.ascontiguousarray(fm.keys, dtype=np.uint8)
    if keys.size == 0:
        keys = np.zeros(1, dtype=np.uint8)
    off = np.ascontiguousarray(fm.key_off, dtype=np.uint64)
    h = C.c_void_p()
    if L.oracle_model_from_keys(_ptr(keys, _u8p), _ptr(off, _u64p), len(fm), int(fm.tokens), int(fm.types), int(indexed), C.byref(h)) != 0:
        raise RuntimeError("oracle_model_from_keys: " + L.oracle_last_error().decode())
    return LoadedModel(h, bool(indexed))


def train_constrained(corpus, constrain: LoadedModel, inplace=False, **kw) -> FlatModel:
    """train() under a constraint model (include/patternmodel.h:880-1345 with constrainbymodel != NULL)."""
    L = lib()
    data = _as_u8(corpus)
    o = default_options(**kw)
    h = C.c_void_p()
    if L.oracle_train_constrained(_ptr(data, _u8p), data.size, C.byref(o), constrain.h, int(bool(inplace)), C.byref(h)) != 0:
        raise RuntimeError("oracle_train_constrained: " + L.oracle_last_error().decode())
    try:
        return _flat_from_handle(h, bool(o.indexed))
    finally:
        L.oracle_model_free(h)


def train_to_modelfile(corpus, **kw) -> bytes:
    """oracle_train + oracle_model_write: the .colibri.patternmodel bytes (canonical order)."""
    L = lib()
    data = _as_u8(corpus)
    o = default_options(**kw)
    h = C.c_void_p()
    if L.oracle_train(_ptr(data, _u8p), data.size, C.byref(o), C.byref(h)) != 0:
        raise RuntimeError("oracle_train: " + L.oracle_last_error().decode())
    try:
        need = L.oracle_model_write(h, None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        L.oracle_model_write(h, _ptr(buf, _u8p), need)
        return buf.tobytes()
    finally:
        L.oracle_model_free(h)


def parse_modelfile(blob) -> FlatModel:
    """Parse a .colibri.patternmodel (type 10 or 20) written by the reference, by us, or by the oracle."""
    L = lib()
    data = _as_u8(blob)
    hdr = (C.c_uint64 * 7)()
    if L.oracle_modelfile_scan(_ptr(data, _u8p), data.size, hdr) != 0:
        raise RuntimeError("parse_modelfile: " + L.oracle_last_error().decode())
    typ, _ver, tokens, types, np_, kb, nr = (int(x) for x in hdr)
    keys = np.zeros(kb + 1, dtype=np.uint8)
    key_off = np.zeros(np_ + 1, dtype=np.uint64)
    counts = np.zeros(np_, dtype=np.uint32)
    rs = rt = ro = None
    if typ == 20:
        rs = np.zeros(nr, dtype=np.uint32)
        rt = np.zeros(nr, dtype=np.uint16)
        ro = np.zeros(np_ + 1, dtype=np.uint64)
    rc = L.oracle_modelfile_parse(_ptr(data, _u8p), data.size, _ptr(keys, _u8p), _ptr(key_off, _u64p), _ptr(counts, _u32p), _ptr(rs, _u32p) if typ == 20 else None,
                                  _ptr(rt, _u16p) if typ == 20 else None, _ptr(ro, _u64p) if typ == 20 else None)
    if rc != 0:
        raise RuntimeError("parse_modelfile: " + L.oracle_last_error().decode())
    return FlatModel(keys[:kb], key_off, counts, tokens, types, model_type=typ, ref_sentence=rs, ref_token=rt, ref_off=ro)


# --------------------------------------------------------------------------- L1 helpers
def inttobytes(cls: int) -> bytes:
    buf = (C.c_uint8 * 8)()
    n = lib().oracle_inttobytes(buf, cls)
    return bytes(buf[:n])


def bytestoint(b: bytes):
    a = (C.c_uint8 * (len(b) + 1))(*b)
    ln = C.c_uint()
    v = lib().oracle_bytestoint(a, C.byref(ln))
    return int(v), int(ln.value)


def spooky_hash64(msg: bytes, seed: int = 0) -> int:
    return int(lib().oracle_spooky_hash64(C.c_char_p(msg), len(msg), seed))


def pattern_hash(key: bytes) -> int:
    a = (C.c_uint8 * (len(key) + 1))(*key)
    return int(lib().oracle_pattern_hash(a, len(key)))


def skip_configurations(n: int, maxskips: int = 3):
    out = (C.c_uint32 * 65536)()
    k = lib().oracle_skip_configurations(n, maxskips, out, 65536)
    if k < 0:
        raise ValueError("n too large")
    return [int(out[i]) for i in range(k)]


def skipgram_collapse(ngram: bytes, mask: int) -> bytes:
    a = (C.c_uint8 * (len(ngram) + 1))(*ngram)
    out = (C.c_uint8 * (len(ngram) + 1))()
    n = lib().oracle_skipgram_collapse(a, len(ngram), mask, out)
    return bytes(out[:n])


def encode_corpus(sentences) -> bytes:
    """[[class ids]] -> .colibri.dat v2 body (src/classencoder.cpp:550-600: varint tokens, 0x00 after every sentence)."""
    out = bytearray()
    for s in sentences:
        for c in s:
            out += inttobytes(int(c))
        out.append(0)
    return bytes(out)


def synth_corpus(ntokens: int, vocab: int = 100000, seed: int = 1, mean_sentence: int = 22, phrase_permille: int = 0, nphrases: int = 0, first_token: int = 0) -> np.ndarray:
    """The counter-based synthetic corpus (body only, no 0xA2 0x02 header), CPU realisation."""
    p = SynthParams(seed, ntokens, vocab, mean_sentence, phrase_permille, nphrases, first_token)
    need = lib().oracle_synth_corpus(C.byref(p), None, 0)
    buf = np.zeros(need, dtype=np.uint8)
    lib().oracle_synth_corpus(C.byref(p), _ptr(buf, _u8p), need)
    return buf


# --------------------------------------------------------------------------- the real reference
def have_ref() -> bool:
    return os.path.exists(REF_TRAIN) and os.access(REF_TRAIN, os.X_OK)


def ref_train(corpus_path: str, model_path: str | None = None, unindexed=True, skipgrams=False, quiet=False, timeout=None, **kw):
    """Run the unmodified reference (oracle/_ref/ref_train).  Returns (stats dict, stderr text).

    kw: t (MINTOKENS), l (MAXLENGTH), m (MINLENGTH), b (MAXBACKOFFLENGTH), y (MINTOKENS_SKIPGRAMS),
        T (MINSKIPTYPES), W (MINTOKENS_UNIGRAMS) -- the CLI letters of src/patternmodeller.cpp:504-618."""
    cmd = [REF_TRAIN, "-f", corpus_path]
    if model_path:
        cmd += ["-o", model_path]
    if unindexed:
        cmd.append("-u")
    if skipgrams:
        cmd.append("-s")
    if quiet:
        cmd.append("-q")
    for k, v in kw.items():
        cmd += ["-" + k, str(v)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("ref_train failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


REF_CLI = os.path.join(REF_DIR, "colibri-patternmodeller")


def ref_cli(args, timeout=None):
    """Run the unmodified reference CLI (oracle/_ref/colibri-patternmodeller).  Returns (exit code, stderr text)."""
    r = subprocess.run([REF_CLI] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stderr


def parse_ref_passes(stderr: str):
    """Per-pass (found_ngrams, found_skipgram_occurrences|None, pruned, kept) from the reference's progress lines
    (include/patternmodel.h:1195-1245: ' Found X ngrams...[S skipgram occurrences...]pruned Y[ plus E extra skipgrams..]...total kept: Z')."""
    import re

    out = []
    for line in stderr.splitlines():
        m = re.search(r"Found (\d+) ngrams\.\.\.(?:(\d+) skipgram occurrences\.\.\.)?.*?pruned (\d+)(?: plus (\d+) extra skipgrams\.\.)?\.\.\.total kept: (-?\d+)", line)
        if m:
            found, sk, pr, extra, kept = m.groups()
            out.append((int(found), None if sk is None else int(sk), int(pr) + (int(extra) if extra else 0), int(kept)))
            continue
        # IndexedPatternModel::trainskipgrams (:2995-3006): " Found X skipgrams...pruned Y[ plus E extra skipgrams..]...total kept: Z"
        m = re.search(r"Found (\d+) skipgrams\.\.\.pruned (\d+)(?: plus (\d+) extra skipgrams\.\.)?\.\.\.total kept: (-?\d+)", line)
        if m:
            sk, pr, extra, kept = m.groups()
            out.append((0, int(sk), int(pr) + (int(extra) if extra else 0), int(kept)))
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# Relations on an indexed model: restatement of the reference's getreverseindex, getrightcooc, getleftcooc,
# npmi, computenpmi and computeflexgrams_fromcooc.  Pure Python (small cases only); pinned to the unmodified reference through
# oracle/ref_relations.cpp (tests/golden/make_golden_relations.py, tests/test_oracle_relations.py).
def corpus_sentences(body) -> list:
    """Sentences of a class-encoded corpus as lists of token byte strings, the way IndexedCorpus sees them (preloaded source:
    a missing final delimiter still closes the last sentence; empty sentences are kept and numbered)."""
    b = bytes(_as_u8(body).tobytes()) if not isinstance(body, (bytes, bytearray)) else bytes(body)
    sentences, cur, tok = [], [], bytearray()
    for x in b:
        if x >= 128:
            tok.append(x)
        elif x == 0 and not tok:
            sentences.append(cur)
            cur = []
        else:
            tok.append(x)
            cur.append(bytes(tok))
            tok = bytearray()
    if cur or tok:
        sentences.append(cur)
    return sentences


def reverse_index(body, patterns: dict, minn: int | None = None, maxn: int | None = None) -> dict:
    """getreverseindex for every position (include/patternmodel.h:1746-1824), n-grams: {(sentence, token): [pattern bytes, by length]}.
    patterns: {pattern bytes: count} (n-grams); sentences count from 1."""
    ntok = {k: len(_split_tokens(k)) for k in patterns}
    if minn is None:
        minn = min(ntok.values()) if ntok else 1
    if maxn is None:
        maxn = max(ntok.values()) if ntok else 0
    out = {}
    for s, sent in enumerate(corpus_sentences(body), start=1):
        for t in range(len(sent)):
            found = []
            n = minn
            while t + n <= len(sent) and n <= maxn:
                key = b"".join(sent[t:t + n])
                if key in patterns:
                    found.append(key)
                n += 1
            out[(s, t)] = found
    return out


def _split_tokens(key: bytes) -> list:
    toks, cur = [], bytearray()
    for x in key:
        cur.append(x)
        if x < 128:
            toks.append(bytes(cur))
            cur = bytearray()
    return toks


def cooc(body, patterns: dict, left: bool = False) -> dict:
    """getrightcooc / getleftcooc of every pattern as the reference computes them (:3460-3493, :3502-3531): getreverseindex_right / _left
    (:1867-1878, :1885-1892) call getreverseindex(ref, ...) with the ORIGINAL reference for every neighbouring position ref2, so the
    neighbours reported at ref2 are the patterns that start at ref itself.  Returns {(P, Q): joint}."""
    rindex = reverse_index(body, patterns)
    sents = corpus_sentences(body)
    out = {}
    for (s, t), here in rindex.items():
        sl = len(sents[s - 1])
        for P in here:  # every occurrence (s, t) of P: the model's occurrence list is the set of positions where the window is P
            nP = len(_split_tokens(P))
            for Q in here:
                nQ = len(_split_tokens(Q))
                if left:
                    w = sum(1 for i in range(0, t) if i + nQ < t)  # ref2.token + n(neighbour) < ref.token
                else:
                    w = sum(1 for i in range(t + 1, sl) if i > t + nP)  # ref2.token > ref.token + n(pattern)
                if w:
                    out[(P, Q)] = out.get((P, Q), 0) + w
    return out


def cooc_both(body, patterns: dict, occurrencethreshold: int = 0, size: int = 0, ordersignificant: bool = False) -> dict:
    """getcooc of every pattern (:3543-3576): for every occurrence (s, t) of P and every model pattern Q that starts at a position t2 of the
    same sentence (getreverseindex_bysentence :1850-1862), one count if the two do not overlap and are not adjacent either --
    t2 + |Q| < t or t2 > t + |P|.  Neighbours occurring less than occurrencethreshold times or of another size are skipped,
    ordersignificant skips neighbours that sort before P (Pattern::operator<, src/pattern.cpp:1114-1125: bytewise), and relations counted
    less than occurrencethreshold times are pruned (:3066-3078).  Returns {(P, Q): count}."""
    rindex = reverse_index(body, patterns)
    ntok = {k: len(_split_tokens(k)) for k in patterns}
    by_sentence = {}
    for (s, t), here in rindex.items():
        by_sentence.setdefault(s, []).extend((t, q) for q in here)
    out = {}
    for s, items in by_sentence.items():
        for t, P in items:
            for t2, Q in items:
                if ordersignificant and Q < P:
                    continue
                if not (t2 + ntok[Q] < t or t2 > t + ntok[P]):
                    continue
                if occurrencethreshold and patterns[Q] < occurrencethreshold:
                    continue
                if size and ntok[Q] != size:
                    continue
                out[(P, Q)] = out.get((P, Q), 0) + 1
    if occurrencethreshold:
        out = {k: v for k, v in out.items() if v >= occurrencethreshold}
    return out


def group_stats(patterns: dict) -> dict:
    """The model's group statistics {(category, n): (occurrences, patterns, word types)} for category 0..3 (0 = all, 1 n-gram, 2 skipgram,
    3 flexgram) and n = 0 (all lengths) .. longest pattern: computestats (:1903-1933; flexgrams have no per-length entry) and
    computecoveragestats (:1946-1984; the word types of a group are the distinct tokens of its patterns, a gap counts as a token; asked for
    length 1 only the unigram patterns themselves count).  patterns: {pattern bytes: occurrence count}."""
    shape = {}
    maxn = 0
    for k in patterns:
        toks = _split_tokens(k)
        cat = 3 if b"\x04" in toks else 2 if b"\x03" in toks else 1
        shape[k] = (cat, len(toks), toks)
        maxn = max(maxn, len(toks))
    out = {}
    for c in range(4):
        for n in range(maxn + 1):
            occ = npat = 0
            types = set()
            for k, (cat, pn, toks) in shape.items():
                if c and cat != c:
                    continue
                if n == 0 or (pn == n and cat != 3):
                    occ += patterns[k]
                    npat += 1
                if pn == 1 and n <= 1:
                    types.add(k)
                elif n == 0 or pn == n:
                    types.update(toks)
            out[(c, n)] = (occ, npat, len(types))
    return out


def npmi(count1: int, count2: int, joint: int, total: int) -> float:
    """PatternModel::npmi (:3582-3585), the same expression in the same order (the product is an unsigned 32-bit product in the reference)."""
    import math

    prod = (count1 * count2) & 0xFFFFFFFF
    return math.log(joint / prod) / -math.log(joint / total)


def flexgrams_fromcooc(body, patterns: dict, threshold: float) -> tuple:
    """computeflexgrams_fromcooc (:3751-3774) with a clean iteration over the patterns the model held before the call (the reference inserts
    into the map it iterates over).  Every match (ref, ref2) of getrightcooc(P) adds ref to EVERY flexgram P {*} Q whose npmi passes, so a
    flexgram's occurrence count is the number of matches of P.  Returns (found, {flexgram bytes: occurrence count})."""
    right = cooc(body, patterns, left=False)
    total = sum(patterns.values())
    matches = {}
    for (P, _Q), j in right.items():
        matches[P] = matches.get(P, 0) + j
    flex = {}
    for (P, Q), j in right.items():
        if npmi(patterns[P], patterns[Q], j, total) >= threshold:
            flex[P + b"\x04" + Q] = matches[P]
    return len(flex), flex
