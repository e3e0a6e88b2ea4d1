"""The oracle (oracle/oracle.c, our CPU restatement) against golden vectors produced by the UNMODIFIED reference
(tests/golden/golden.json, written by tests/golden/make_golden.py from oracle/_ref) and against the known-answer
numbers in the reference's own tests (src/test.cpp, test.py).  CPU only."""
import pytest

import oracle
from conftest import case_id, corpus_body, load_cases

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[case_id(c) for c in CASES])
def test_oracle_matches_reference_model(golden, case):
    body = corpus_body(golden, case["corpus"])
    m = oracle.train(body, **case["options"])
    assert m.tokens == case["tokens"]
    assert m.types == case["types"]
    assert len(m) == case["patterns"]
    assert (m.maxn, m.minn) == (case["maxn"], case["minn"])
    assert int(m.hasskipgrams) == case["hasskipgrams"]
    assert int(m.counts.sum()) == case["occurrences"]
    # per-pass "Found X ngrams ... pruned Y" numbers of the reference's progress lines (patternmodel.h:1195-1245)
    assert [(p[1], p[3]) for p in m.passes] == [(p[0], p[2]) for p in case["passes"]]
    assert m.digest() == case["digest"]
    if "model" in case:
        c = m.canonical()
        got = [[c.key(i).hex(), int(c.counts[i])] + ([[list(r) for r in c.refs(i)]] if c.ref_off is not None else []) for i in range(len(c))]
        assert got == case["model"]


def test_reference_known_answers_hamlet(golden):
    """src/test.cpp:1211-1246: 111 patterns / 186 types / 354 tokens, 'or not to' occurs 6 times; :1261-1283: 385 with skipgrams."""
    body = corpus_body(golden, "hamlet")
    m = oracle.train(body)
    assert (len(m), m.types, m.tokens) == (111, 186, 354)
    assert sum(1 for i in range(len(m)) if sum(b < 128 for b in m.key(i)) == 1) == 45  # covered unigram types, :1221
    cls = {}
    for line in open(__import__("os").path.join(__import__("conftest").GOLDEN_DIR, "hamlet.colibri.cls"), encoding="utf-8"):
        k, w = line.rstrip("\n").split("\t")
        cls[w] = int(k)
    key = b"".join(oracle.inttobytes(cls[w]) for w in "or not to".split())
    assert m.as_dict()[key] == 6
    assert m.as_dict()[oracle.inttobytes(cls["not"])] == 7  # src/test.cpp:1729-1733
    s = oracle.train(body, doskipgrams_exhaustive=1, streamed=0)
    assert (len(s), s.types, s.tokens) == (385, 186, 354)  # test.py:236-238
    # probe numbers recorded in SURVEY.md section 4 / BASELINE.md
    assert [sum(1 for i in range(len(m5)) if sum(b < 128 for b in m5.key(i)) == n) for m5 in [oracle.train(body, maxlength=5)] for n in range(1, 6)] == [45, 22, 14, 12, 9]


def test_spooky_known_answers(golden):
    """SpookyHash::Hash64 (include/SpookyV2.h:59-66 -> Short, src/SpookyV2.cpp:21-113) for every length 1..191."""
    assert len(golden["spooky"]) > 190
    for hexmsg, h in golden["spooky"]:
        assert oracle.spooky_hash64(bytes.fromhex(hexmsg)) == h, hexmsg
    assert oracle.pattern_hash(b"") == 0  # src/pattern.cpp:235-236


def test_skip_configurations(golden):
    """compute_skip_configurations (src/algorithms.cpp:79-94); 1/3/15 masks for n=3/4/6 (src/test.cpp:1119-1159)."""
    for k, masks in golden["masks"].items():
        n, ms = (int(x) for x in k.split(","))
        assert oracle.skip_configurations(n, ms) == masks
    assert [len(oracle.skip_configurations(n, 3)) for n in (3, 4, 6)] == [1, 3, 15]


def test_codec_roundtrip():
    """inttobytes / bytestoint (src/classencoder.cpp:22-42, src/classdecoder.cpp:20-43)."""
    assert oracle.inttobytes(6) == b"\x06" and oracle.inttobytes(127) == b"\x7f"
    assert oracle.inttobytes(128) == b"\x80\x01" and oracle.inttobytes(300) == b"\xac\x02"
    assert oracle.inttobytes(16384) == b"\x80\x80\x01"
    for c in list(range(0, 400)) + [16383, 16384, 2097151, 2097152, 2**28 - 1, 2**28, 2**32 - 1]:
        b = oracle.inttobytes(c)
        assert oracle.bytestoint(b) == (c, len(b))
        assert all(x >= 128 for x in b[:-1]) and b[-1] < 128


def test_skipgram_collapse():
    """Pattern(const PatternPointer&) for skipgrams: src/pattern.cpp:886-908 (SURVEY 8a5 probe: 10,300,12 mask 0b010 -> 0A 03 0C)."""
    ng = oracle.inttobytes(10) + oracle.inttobytes(300) + oracle.inttobytes(12)
    assert oracle.skipgram_collapse(ng, 0b010) == bytes([0x0A, 0x03, 0x0C])
    assert oracle.skipgram_collapse(bytes([10, 11, 12, 13, 14]), 0b00110) == bytes([0x0A, 0x03, 0x03, 0x0D, 0x0E])


def test_modelfile_roundtrip(golden):
    """oracle_model_write emits the reference layout (patternmodel.h:1609-1624, patternstore.h:534-542); header bytes of
    the hamlet n<=3 model as probed from the reference: 00 0A 02, 354, 186, 81; 563 bytes."""
    body = corpus_body(golden, "hamlet")
    blob = oracle.train_to_modelfile(body, mintokens=2, maxlength=3)
    assert len(blob) == 563
    assert blob[:3] == bytes([0, 10, 2])
    assert int.from_bytes(blob[3:11], "little") == 354 and int.from_bytes(blob[11:19], "little") == 186 and int.from_bytes(blob[19:27], "little") == 81
    back = oracle.parse_modelfile(blob)
    assert back.same_patterns(oracle.train(body, mintokens=2, maxlength=3))
    iblob = oracle.train_to_modelfile(body, mintokens=2, maxlength=5, indexed=1, streamed=0)
    assert len(iblob) == 2795 and iblob[1] == 20  # SURVEY section 4 probe: indexed -l 5 model file is 2795 bytes
    assert oracle.parse_modelfile(iblob).same_patterns(oracle.train(body, mintokens=2, maxlength=5, indexed=1, streamed=0))


def test_synth_corpus_is_wellformed():
    body = oracle.synth_corpus(50000, vocab=3000, seed=3, mean_sentence=10, phrase_permille=200, nphrases=100).tobytes()
    assert body.endswith(b"\x00")
    m = oracle.train(body, mintokens=1, maxlength=1)
    assert m.tokens == 50000
    # classes are 6 .. V+5 only (never the reserved 0-5, SURVEY 8a parity trap 7)
    for i in range(len(m)):
        c, _ = oracle.bytestoint(m.key(i))
        assert 6 <= c < 3006
    again = oracle.synth_corpus(50000, vocab=3000, seed=3, mean_sentence=10, phrase_permille=200, nphrases=100).tobytes()
    assert again == body
