"""The reference's Python binding (`import colibricore`, /root/reference/colibricore_wrapper.in.pyx + colibricore_patternmodel.pxi) restated on the
B200 library for the pattern-model path: the same class and method names and argument meaning, so that the model part of the reference's
test.py (:230-311) reads the same against this module:

    import colibricore_b200 as colibricore
    options = colibricore.PatternModelOptions(mintokens=2, maxlength=5)
    model = colibricore.IndexedPatternModel(reverseindex=colibricore.IndexedCorpus("corpus.colibri.dat"))
    model.train("corpus.colibri.dat", options)

Training, loading, has/occurrencecount, the reverse index and the co-occurrence relations run on the GPU through include/colibri_b200.h (ctypes);
what stays here is string work: class files, Pattern <-> text, reports.  Nothing falls back to a CPU implementation of the hot path: without the
library or a GPU the calls raise.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import (INDEXEDPATTERNMODEL, UNINDEXEDPATTERNMODEL, ColibriError, Corpus, Model, ReverseIndex, load_model, train, train_constrained)
from . import PatternModelOptions as _Options

NGRAM, SKIPGRAM, FLEXGRAM = 1, 2, 3  # PatternCategory, include/pattern.h


# --------------------------------------------------------------------------------------------- patterns and class files
def _split(data: bytes):
    toks, cur = [], bytearray()
    for x in data:
        cur.append(x)
        if x < 128:
            toks.append(bytes(cur))
            cur = bytearray()
    return toks


def _class_of(tok: bytes) -> int:
    v = 0
    for i, x in enumerate(tok):  # little-endian base 128, continuation bytes carry bit 7 (src/common.cpp bytestoint)
        v |= (x & 0x7F) << (7 * i)
    return v


def _bytes_of(cls: int) -> bytes:
    out = bytearray()
    while True:
        if cls < 128:
            out.append(cls)
            return bytes(out)
        out.append((cls & 0x7F) | 0x80)
        cls >>= 7


class Pattern:
    """A class-encoded pattern (include/pattern.h): its bytes are the model's key."""

    __slots__ = ("data",)

    def __init__(self, data: bytes = b""):
        self.data = bytes(data)

    def __bytes__(self):
        return self.data

    def __len__(self):
        return len(_split(self.data))

    def __hash__(self):
        return hash(self.data)

    def __eq__(self, other):
        return isinstance(other, Pattern) and self.data == other.data

    def __lt__(self, other):
        return self.data < other.data

    def __add__(self, other):
        return Pattern(self.data + bytes(other))

    def __iter__(self):
        return (Pattern(t) for t in _split(self.data))

    def __getitem__(self, item):
        toks = _split(self.data)
        if isinstance(item, slice):
            return Pattern(b"".join(toks[item]))
        return Pattern(toks[item])

    def bytesize(self):
        return len(self.data)

    def category(self):
        classes = [_class_of(t) for t in _split(self.data)]
        return FLEXGRAM if 4 in classes else SKIPGRAM if 3 in classes else NGRAM

    def skipcount(self):
        n, prev = 0, False
        for c in (_class_of(t) for t in _split(self.data)):
            gap = c in (3, 4)
            n += gap and not prev
            prev = gap
        return n

    def tostring(self, decoder: "ClassDecoder") -> str:
        return decoder.decode(self.data)

    def __repr__(self):
        return "Pattern(%s)" % self.data.hex()


class ClassDecoder:
    """class -> word, from a .colibri.cls file (one `class<TAB>word` line per class; src/classdecoder.cpp)."""

    def __init__(self, filename: str | None = None):
        self.words = {0: "\n", 1: "{|}", 2: "{?}", 3: "{*}", 4: "{**}"}  # boundary, unknown, skip, flex (include/classdecoder.h:48-52, src/classdecoder.cpp:86-89)
        if filename:
            with open(filename, encoding="utf-8") as f:
                for line in f:
                    line = line.rstrip("\n")
                    if "\t" in line:
                        c, w = line.split("\t", 1)
                        self.words[int(c)] = w

    def __len__(self):
        return len(self.words)

    def decode(self, data: bytes) -> str:
        return " ".join(self.words.get(_class_of(t), "{?}") for t in _split(data))


class ClassEncoder:
    """word -> class, from a .colibri.cls file (src/classencoder.cpp); buildpattern() encodes a space-separated string."""

    def __init__(self, filename: str | None = None):
        self.classes = {"{|}": 1, "{?}": 2, "{*}": 3, "{**}": 4}  # src/classencoder.cpp:128-131
        if filename:
            with open(filename, encoding="utf-8") as f:
                for line in f:
                    line = line.rstrip("\n")
                    if "\t" in line:
                        c, w = line.split("\t", 1)
                        self.classes[w] = int(c)

    def __len__(self):
        return len(self.classes)

    def buildpattern(self, text: str, allowunknown: bool = True, autoaddunknown: bool = False) -> Pattern:
        """(src/classencoder.cpp:364-433) `{*}` a gap, `{**}` a flexible gap, `{?}` the unknown word, `{*N*}` N gaps in a row; a word that is
        not in the class file becomes the unknown class (2), gets a new class (autoaddunknown) or raises KeyError (allowunknown=False)."""
        out = bytearray()
        for w in text.split():
            if len(w) > 4 and w.startswith("{*") and w.endswith("*}") and w[2:-2].isdigit():
                out += _bytes_of(3) * int(w[2:-2])
            elif w in self.classes:
                out += _bytes_of(self.classes[w])
            elif autoaddunknown:
                self.classes[w] = max(max(self.classes.values()), 5) + 1
                out += _bytes_of(self.classes[w])
            elif not allowunknown:
                raise KeyError(w)
            else:
                out += _bytes_of(2)
        return Pattern(bytes(out))


class PatternModelOptions:
    """The reference binding's option object (colibricore_wrapper.in.pyx:870-990): attributes MINTOKENS, MAXLENGTH, ... (any case here), the same
    names as constructor keywords."""

    _MAP = {"mintokens": "MINTOKENS", "maxlength": "MAXLENGTH", "minlength": "MINLENGTH", "maxbackofflength": "MAXBACKOFFLENGTH", "mintokens_unigrams": "MINTOKENS_UNIGRAMS",
            "mintokens_skipgrams": "MINTOKENS_SKIPGRAMS", "minskiptypes": "MINSKIPTYPES", "maxskips": "MAXSKIPS", "doskipgrams": "DOSKIPGRAMS",
            "doskipgrams_exhaustive": "DOSKIPGRAMS_EXHAUSTIVE", "doreverseindex": "DOREVERSEINDEX", "dopatternperline": "DOPATTERNPERLINE", "doreset": "DORESET",
            "prunenonsubsumed": "PRUNENONSUBSUMED", "quiet": "QUIET", "debug": "DEBUG", "device": "device"}

    # what an option reads as before it is set: the reference's defaults (include/patternmodel.h:105-180); QUIET is on here because the
    # progress lines are the C++ front end's business, not the binding's
    _DEFAULTS = {"mintokens": -1, "maxlength": 100, "minlength": 1, "maxbackofflength": 100, "mintokens_unigrams": 1, "mintokens_skipgrams": -1, "minskiptypes": 2,
                 "maxskips": 3, "doskipgrams": False, "doskipgrams_exhaustive": False, "doreverseindex": True, "dopatternperline": False, "doreset": False,
                 "prunenonsubsumed": 0, "quiet": True, "debug": False, "device": 0}

    def __init__(self, **kw):
        object.__setattr__(self, "_kw", {"quiet": True})
        for k, v in kw.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        k = k.lower()  # the reference's attributes are upper case, its constructor keywords any case (colibricore_wrapper.in.pyx:898-900)
        if k not in self._MAP:
            raise KeyError("No such option: " + k)
        self._kw[k] = int(v) if isinstance(v, bool) else v

    def __getattr__(self, k):
        if k.lower() in self._MAP:
            return self._kw.get(k.lower(), self._DEFAULTS[k.lower()])
        raise AttributeError(k)

    def to_native(self, model_type: int, streamed: int) -> _Options:
        kw = {self._MAP[k]: v for k, v in self._kw.items() if self._MAP[k] not in ("DOREVERSEINDEX", "DORESET", "DEBUG")}
        return _Options(model_type=model_type, streamed=streamed, **kw)


class IndexedCorpus:
    """The corpus as a reverse index (include/classdecoder.h / IndexedCorpus in include/patternmodel.h:296-470): sentences and positions.
    The bytes live on the device (staged once); sentence boundaries come from the device tokeniser."""

    def __init__(self, filename: str | None = None, body: bytes | None = None, device: int = 0):
        if filename is not None:
            with open(filename, "rb") as f:
                blob = f.read()
            body = blob[2:] if blob[:1] == b"\xa2" else blob
        self.body = bytes(body or b"")
        self.device = device
        self.corpus = Corpus.from_bytes(self.body, device=device) if self.body else None
        self._tok = None

    def _tokens(self):
        if self._tok is None:
            self._tok = self.corpus.tokens() if self.corpus is not None else np.zeros(0, dtype=np.uint32)
            ends = np.flatnonzero(self._tok == 0)
            starts = np.concatenate([[0], ends + 1])
            if len(self._tok) and self._tok[-1] != 0:  # a missing final delimiter still closes the last sentence
                ends = np.concatenate([ends, [len(self._tok)]])
            else:
                starts = starts[:-1]
            self._bounds = list(zip(starts.tolist(), ends.tolist()))
        return self._tok

    def sentencecount(self) -> int:
        self._tokens()
        return len(self._bounds)

    def sentencelength(self, sentence: int) -> int:
        self._tokens()
        a, b = self._bounds[sentence - 1]
        return b - a

    def getsentence(self, sentence: int) -> Pattern:
        tok = self._tokens()
        a, b = self._bounds[sentence - 1]
        return Pattern(b"".join(_bytes_of(int(c)) for c in tok[a:b]))

    def sentences(self):
        for i in range(1, self.sentencecount() + 1):
            yield self.getsentence(i)

    def __len__(self):
        tok = self._tokens()
        return int((tok != 0).sum())

    def __iter__(self):
        tok = self._tokens()
        for s, (a, b) in enumerate(self._bounds, start=1):
            for t in range(a, b):
                yield (s, t - a), Pattern(_bytes_of(int(tok[t])))


# --------------------------------------------------------------------------------------------- pattern models
class _PatternModel:
    MODELTYPE = UNINDEXEDPATTERNMODEL

    def __init__(self, filename: str = "", options: PatternModelOptions | None = None, constrainmodel=None, reverseindex: IndexedCorpus | None = None, device: int = 0):
        self._m: Model | None = None
        self._export = None
        self._index = None
        self._rindex: ReverseIndex | None = None
        self._right = self._left = None  # all getrightcooc / getleftcooc relations, fetched once
        self.corpus = reverseindex
        self.device = device
        if filename:
            self.load(filename, options, constrainmodel)

    # ---- building
    def _opts(self, options, streamed):
        return (options or PatternModelOptions()).to_native(self.MODELTYPE, streamed)

    def _set(self, model: Model):
        self._m, self._export, self._index = model, None, None
        self._right = self._left = None
        if self._rindex is not None:
            self._rindex.close()
            self._rindex = None

    def train(self, filename: str, options: PatternModelOptions | None = None, constrainmodel=None):
        """PatternModel::train (include/patternmodel.h:880-1345): from a .colibri.dat file, or from the reverse index when no file name is given."""
        if filename:
            with open(filename, "rb") as f:
                blob = f.read()
            body = blob[2:] if blob[:1] == b"\xa2" else blob
        elif self.corpus is not None:
            body = self.corpus.body
        else:
            raise ValueError("No filename and no reverse index")
        # the reference streams the file unless the model was given a reverse index (the CLI preloads the corpus for indexed models)
        streamed = 0 if self.corpus is not None else 1
        o = self._opts(options, streamed)
        if constrainmodel is not None:
            self._set(train_constrained(Corpus.from_bytes(body, device=self.device), constrainmodel._m, options=o))
        else:
            self._set(train(body, o))

    def load(self, filename: str, options: PatternModelOptions | None = None, constrainmodel=None):
        with open(filename, "rb") as f:
            blob = f.read()
        o = self._opts(options, 1)
        if options is None or options.mintokens is None:
            o.MINTOKENS = -1
        self._set(load_model(blob, o, constrainmodel._m if constrainmodel is not None else None))

    def write(self, filename: str):
        self._m.write(filename)

    # ---- content
    def _flat(self):
        if self._export is None:
            keys, off, counts, refs = self._m.export()
            kb = keys.tobytes()
            self._export = ([kb[int(off[i]):int(off[i + 1])] for i in range(len(counts))], counts, refs)
            self._index = {k: i for i, k in enumerate(self._export[0])}
        return self._export

    def __len__(self):
        return len(self._m) if self._m is not None else 0

    def types(self):
        return self._m.types()

    def tokens(self):
        return self._m.tokens()

    def minlength(self):
        return self._m.minlength()

    def maxlength(self):
        return self._m.maxlength()

    def type(self):
        return self.MODELTYPE

    def __contains__(self, pattern):
        return self._m.has(bytes(pattern))

    def has(self, pattern):
        return pattern in self

    def occurrencecount(self, pattern) -> int:
        return self._m.occurrencecount(bytes(pattern))

    def frequency(self, pattern) -> float:
        """occurrences / total occurrences of the patterns of the same category and size (PatternModel::frequency)."""
        keys, counts, _ = self._flat()
        p = Pattern(bytes(pattern))
        n, cat = len(p), p.category()
        total = sum(int(c) for k, c in zip(keys, counts) if len(_split(k)) == n and Pattern(k).category() == cat)
        return self.occurrencecount(pattern) / total if total else 0.0

    def __iter__(self):
        return (Pattern(k) for k in self._flat()[0])

    def totaloccurrencesingroup(self, category: int = 0, n: int = 0) -> int:
        """(computestats :1903-1933; flexgrams have no per-length entry)"""
        keys, counts, _ = self._flat()
        return sum(int(c) for k, c in zip(keys, counts) if self._in_group(k, category, n))

    def totalpatternsingroup(self, category: int = 0, n: int = 0) -> int:
        return sum(1 for k in self._flat()[0] if self._in_group(k, category, n))

    def totalwordtypesingroup(self, category: int = 0, n: int = 0) -> int:
        """Distinct tokens of the group's patterns, a gap counting as a token; asked for length 1, the unigram patterns themselves
        (computecoveragestats :1946-1984, without its cache: see tests/golden/make_golden_stats.py)."""
        types = set()
        for k in self._flat()[0]:
            if category and Pattern(k).category() != category:
                continue
            toks = _split(k)
            if len(toks) == 1 and n <= 1:
                types.add(k)
            elif n == 0 or len(toks) == n:
                types.update(toks)
        return len(types)

    @staticmethod
    def _in_group(key: bytes, category: int, n: int) -> bool:
        cat = Pattern(key).category()
        return (not category or cat == category) and (not n or (cat != FLEXGRAM and len(_split(key)) == n))

    # ---- views (text only; the numbers come from the model)
    def printmodel(self, decoder: ClassDecoder):
        for pattern, value in self.items():
            print("%s\t%d" % (pattern.tostring(decoder), value if isinstance(value, int) else len(value)))

    def report(self):
        keys, counts, _ = self._flat()
        print("REPORT\n  patterns: %d\n  tokens: %d\n  types: %d" % (len(self), self.tokens(), self.types()))
        by = {}
        for k, c in zip(keys, counts):
            cat, n = Pattern(k).category(), len(_split(k))
            a = by.setdefault((cat, n), [0, 0])
            a[0] += 1
            a[1] += int(c)
        for (cat, n), (p, o) in sorted(by.items()):
            print("  %s n=%d: %d patterns, %d occurrences" % ({1: "n-gram", 2: "skipgram", 3: "flexgram"}[cat], n, p, o))

    def histogram(self):
        h = {}
        for c in self._flat()[1]:
            h[int(c)] = h.get(int(c), 0) + 1
        print("HISTOGRAM")
        for c in sorted(h):
            print("%d\t%d" % (c, h[c]))

    # ---- flexgrams
    def computeflexgrams_fromskipgrams(self) -> int:
        found, m = self._m.flexgrams_fromskipgrams()
        self._set(m)
        return found


class UnindexedPatternModel(_PatternModel):
    MODELTYPE = UNINDEXEDPATTERNMODEL

    def __getitem__(self, pattern) -> int:
        if pattern not in self:
            raise KeyError(pattern)
        return self.occurrencecount(pattern)

    def items(self):
        keys, counts, _ = self._flat()
        return ((Pattern(k), int(c)) for k, c in zip(keys, counts))


class IndexedPatternModel(_PatternModel):
    MODELTYPE = INDEXEDPATTERNMODEL

    def _refs(self, i):
        _, _, refs = self._flat()
        rs, rt, ro = refs
        return [(int(rs[j]), int(rt[j])) for j in range(int(ro[i]), int(ro[i + 1]))]

    def getdata(self, pattern):
        self._flat()
        i = self._index.get(bytes(pattern))
        if i is None:
            raise KeyError(pattern)
        return self._refs(i)

    __getitem__ = getdata

    def items(self):
        keys, _, _ = self._flat()
        return ((Pattern(k), self._refs(i)) for i, k in enumerate(keys))

    # ---- reverse index and relations (device: colibri_b200_rindex_*)
    def _ri(self) -> ReverseIndex:
        if self.corpus is None or self.corpus.corpus is None:
            raise ValueError("No reverse index loaded")
        if self._rindex is None:
            self._rindex = ReverseIndex(self._m, self.corpus.corpus, streamed=0)
        return self._rindex

    def reverseindex(self):
        return self.corpus

    def getreverseindex(self, indexreference, occurrencecount: int = 0, category: int = 0, size: int = 0):
        """Generator over the patterns of the model that begin at (sentence, token) (include/patternmodel.h:1746-1824; n-grams)."""
        if not isinstance(indexreference, tuple) or len(indexreference) != 2:
            raise ValueError("Expected tuple")
        yield from self.getreverseindex_batch([indexreference], occurrencecount, category, size)[0]

    def getreverseindex_batch(self, refs, occurrencecount: int = 0, category: int = 0, size: int = 0):
        ri = self._ri()
        keys, counts, _ = self._flat()
        out = []
        for row in ri.query(refs):
            found = []
            for k, idx1 in enumerate(row):
                if idx1 and (not size or ri.lengths[k] == size) and (not category or category == NGRAM) and (not occurrencecount or counts[idx1 - 1] >= occurrencecount):
                    found.append(Pattern(keys[idx1 - 1]))
            out.append(found)
        return out

    def getreverseindex_bysentence(self, sentence: int):
        n = self.corpus.sentencelength(sentence)
        refs = [(sentence, t) for t in range(n)]
        for ref, pats in zip(refs, self.getreverseindex_batch(refs)):
            for p in pats:
                yield ref, p

    def _cooc_all(self, left: bool):
        keys, _, _ = self._flat()
        p, q, j = self._ri().cooc(left=left)
        rel = {}
        for a, b, c in zip(p.tolist(), q.tolist(), j.tolist()):
            rel.setdefault(keys[a], {})[keys[b]] = c
        return rel

    def _cooc_of(self, pattern, left, occurrencethreshold, category, size):
        if pattern not in self:
            raise KeyError(pattern)
        attr = "_left" if left else "_right"
        if getattr(self, attr) is None:
            setattr(self, attr, self._cooc_all(left))
        out = []
        for k, c in getattr(self, attr).get(bytes(pattern), {}).items():
            q = Pattern(k)
            if (occurrencethreshold and self.occurrencecount(q) < occurrencethreshold) or (category and q.category() != category) or (size and len(q) != size):
                continue
            if occurrencethreshold and c < occurrencethreshold:  # prunerelations
                continue
            out.append((q, c))
        return out

    def getrightcooc(self, pattern, occurrencethreshold: int = 0, category: int = 0, size: int = 0):
        """(:3460-3493, with the reference's arithmetic: see include/colibri_b200.h, colibri_b200_rindex_cooc)"""
        return iter(self._cooc_of(pattern, False, occurrencethreshold, category, size))

    def getleftcooc(self, pattern, occurrencethreshold: int = 0, category: int = 0, size: int = 0):
        return iter(self._cooc_of(pattern, True, occurrencethreshold, category, size))

    def getcooc(self, pattern, occurrencethreshold: int = 0, category: int = 0, size: int = 0, ordersignificant: bool = False):
        """Co-occurrence in both directions without overlap (:3543-3576): (neighbour, count) for every model n-gram that occurs in a sentence
        of `pattern` and neither overlaps nor touches it; counted on the device (colibri_b200_rindex_cooc_of), filtered here as the reference
        filters -- neighbours below occurrencethreshold, of another category or size, sorting before `pattern` (ordersignificant), and
        relations counted less than occurrencethreshold times (prunerelations :3066-3078)."""
        if pattern not in self:
            raise KeyError(pattern)
        keys, _, _ = self._flat()
        q, c = self._ri().cooc_of(self._index[bytes(pattern)])
        me = bytes(pattern)
        out = []
        for i, n in zip(q.tolist(), c.tolist()):
            nb = Pattern(keys[i])
            if ordersignificant and keys[i] < me:  # Pattern::operator< compares the bytes (src/pattern.cpp:1114-1125)
                continue
            if (occurrencethreshold and self.occurrencecount(nb) < occurrencethreshold) or (category and nb.category() != category) or (size and len(nb) != size):
                continue
            if occurrencethreshold and n < occurrencethreshold:
                continue
            out.append((nb, n))
        return iter(out)

    def npmi(self, pattern1, pattern2, jointcount: int) -> float:
        """PatternModel::npmi (:3582-3585): the same expression in double precision (the product of the two counts is a 32-bit product there)."""
        prod = (self.occurrencecount(pattern1) * self.occurrencecount(pattern2)) & 0xFFFFFFFF
        return math.log(jointcount / prod) / -math.log(jointcount / self.totaloccurrencesingroup(0, 0))

    def computenpmi(self, threshold: float, right: bool = True, left: bool = False):
        """{pattern: {pattern2: npmi}} of the relations that pass the threshold (:3671-3691; right or left co-occurrence)."""
        if not right and not left:
            return {}
        keys, counts, _ = self._flat()
        total = self.totaloccurrencesingroup(0, 0)
        cnt = {k: int(c) for k, c in zip(keys, counts)}
        out = {}
        if right and left:  # getcooc of every pattern (:3681-3682): one device pass per pattern
            rels = {}
            for i, k in enumerate(keys):
                q, c = self._ri().cooc_of(i)
                if len(q):
                    rels[k] = {keys[b]: n for b, n in zip(q.tolist(), c.tolist())}
        else:
            rels = self._cooc_all(left)
        for p, rel in rels.items():
            for q, j in rel.items():
                v = math.log(j / ((cnt[p] * cnt[q]) & 0xFFFFFFFF)) / -math.log(j / total)
                if v >= threshold:
                    out.setdefault(Pattern(p), {})[Pattern(q)] = v
        return out

    def computeflexgrams_fromcooc(self, threshold: float) -> int:
        """computeflexgrams_fromcooc (:3751-3774), iterating over the patterns the model held before the call: every pair P, Q whose right
        co-occurrence passes the npmi threshold yields the flexgram P {**} Q, and every match of getrightcooc(P) adds its reference to it."""
        keys, counts, refs = self._flat()
        total = self.totaloccurrencesingroup(0, 0)
        cnt = {k: int(c) for k, c in zip(keys, counts)}
        right = self._cooc_all(False)
        ri = self._ri()
        newkeys, newrefs = [], []
        for p, rel in right.items():
            passing = [q for q, j in rel.items() if math.log(j / ((cnt[p] * cnt[q]) & 0xFFFFFFFF)) / -math.log(j / total) >= threshold]
            if not passing:
                continue
            # the matches of P: per occurrence (s, t), one per (position right of the pattern, pattern that starts at (s, t))
            occ = self._refs(self._index[p])
            rows = ri.query(occ)
            n = len(_split(p))
            plist = []
            for (s, t), row in zip(occ, rows):
                w = max(0, self.corpus.sentencelength(s) - 1 - (t + n))
                plist += [(s, t)] * (w * int((row != 0).sum()))
            for q in passing:
                newkeys.append(p + b"\x04" + q)
                newrefs.append(plist)
        found = len(newkeys)
        if not found:
            return 0
        allkeys = list(keys) + newkeys
        allrefs = [self._refs(i) for i in range(len(keys))] + newrefs
        off = np.zeros(len(allkeys) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(k) for k in allkeys])
        ro = np.zeros(len(allkeys) + 1, dtype=np.uint64)
        ro[1:] = np.cumsum([len(r) for r in allrefs])
        rs = np.array([s for r in allrefs for s, _ in r], dtype=np.uint32)
        rt = np.array([t for r in allrefs for _, t in r], dtype=np.uint16)
        cts = np.array([len(r) for r in allrefs], dtype=np.uint32)
        m = Model.from_flat(np.frombuffer(b"".join(allkeys), dtype=np.uint8), off, cts, (rs, rt, ro), tokens=self.tokens(), types=self.types(), model_type=INDEXEDPATTERNMODEL,
                            device=self.device)
        self._set(m)
        return found
