"""The value types of the reference's Python binding names (colibri-core_b200/pybinding.py: Pattern, ClassEncoder, ClassDecoder,
PatternModelOptions) need no GPU: checked here against the class file and the corpus the reference's own tools wrote
(tests/golden/hamlet.colibri.cls / .dat) and against the reference's conventions (reserved classes include/classdecoder.h:48-52,
buildpattern src/classencoder.cpp:364-433, Pattern::category src/pattern.cpp:23-47)."""
import os

import pytest

import oracle
from conftest import GOLDEN_DIR

CLS = os.path.join(GOLDEN_DIR, "hamlet.colibri.cls")
DAT = os.path.join(GOLDEN_DIR, "hamlet.colibri.dat")


@pytest.fixture(scope="module")
def cc():
    import colibricore_b200

    return colibricore_b200


def test_reserved_classes_follow_the_reference(cc):
    enc, dec = cc.ClassEncoder(), cc.ClassDecoder()
    assert (enc.classes["{|}"], enc.classes["{?}"], enc.classes["{*}"], enc.classes["{**}"]) == (1, 2, 3, 4)
    assert (dec.words[1], dec.words[2], dec.words[3], dec.words[4]) == ("{|}", "{?}", "{*}", "{**}")


def test_buildpattern_and_decode_round_trip_the_corpus(cc):
    enc, dec = cc.ClassEncoder(CLS), cc.ClassDecoder(CLS)
    body = open(DAT, "rb").read()[2:]
    sentences = oracle.corpus_sentences(body)
    assert len(sentences) == 40  # reference src/test.cpp:1549
    for sent in sentences:
        data = b"".join(sent)
        text = dec.decode(data)
        assert len(text.split()) == len(sent)
        assert bytes(enc.buildpattern(text)) == data  # the reference's encoder wrote these bytes for these words
    assert dec.decode(b"".join(sentences[0])).split()[0] == "To"  # reference src/test.cpp:1196-1198


def test_buildpattern_known_values(cc):
    enc = cc.ClassEncoder(CLS)
    assert bytes(enc.buildpattern("to be")) == bytes([7, 21])  # classes of the class file
    assert bytes(enc.buildpattern("or not to")) == bytes([14, 12, 7])
    assert bytes(enc.buildpattern("to zyzzyva be")) == bytes([7, 2, 21])  # an unknown word is class 2 (src/classencoder.cpp:420-423)
    assert bytes(enc.buildpattern("to {?} be")) == bytes([7, 2, 21])
    assert bytes(enc.buildpattern("to {*} be")) == bytes([7, 3, 21])
    assert bytes(enc.buildpattern("to {*3*} be")) == bytes([7, 3, 3, 3, 21])  # :404-411
    assert bytes(enc.buildpattern("to {**} be")) == bytes([7, 4, 21])
    with pytest.raises(KeyError):
        enc.buildpattern("to zyzzyva be", allowunknown=False)
    before = len(enc)
    added = enc.buildpattern("zyzzyva zyzzyva", autoaddunknown=True)
    assert len(enc) == before + 1 and len(added) == 2 and added[0] == added[1]
    assert bytes(added[0]) == bytes(enc.buildpattern("zyzzyva"))  # the new class is kept (it is the highest class + 1, :414-417)


def test_pattern_value_semantics(cc):
    P = cc.Pattern
    big = bytes([0x80, 0x01])  # class 128: two bytes
    p = P(bytes([7]) + big + bytes([21]))
    assert (len(p), p.bytesize(), p.category()) == (3, 4, cc.NGRAM)
    assert [bytes(t) for t in p] == [bytes([7]), big, bytes([21])]
    assert bytes(p[1]) == big and bytes(p[0:2]) == bytes([7]) + big and bytes(p[-1]) == bytes([21])
    assert bytes(p + P(bytes([6]))) == bytes(p) + bytes([6])
    assert P(bytes([7, 3, 21])).category() == cc.SKIPGRAM and P(bytes([7, 4, 21])).category() == cc.FLEXGRAM
    assert P(bytes([7, 3, 3, 21, 3, 6])).skipcount() == 2  # runs of gaps count once
    assert P(bytes([7, 21])) == P(bytes([7, 21])) and hash(P(bytes([7, 21]))) == hash(P(bytes([7, 21]))) and P(bytes([7, 21])) != P(bytes([7, 22]))
    assert P(bytes([7])) < P(bytes([7, 21])) < P(bytes([8]))  # bytewise, the shorter first (src/pattern.cpp:1114-1125)
    dec = cc.ClassDecoder(CLS)
    assert P(bytes([7, 3, 21])).tostring(dec) == "to {*} be" and P(bytes([7, 2])).tostring(dec) == "to {?}"
    assert P(big).tostring(cc.ClassDecoder()) == "{?}"  # a class the decoder does not know (src/pattern.cpp:316-320)


def test_options_have_the_reference_defaults(cc):
    o = cc.PatternModelOptions()
    assert (o.MINTOKENS, o.MAXLENGTH, o.MINLENGTH, o.MINSKIPTYPES, o.MAXSKIPS, o.DOSKIPGRAMS, o.DOSKIPGRAMS_EXHAUSTIVE) == (-1, 100, 1, 2, 3, False, False)  # include/patternmodel.h:105-180
    o = cc.PatternModelOptions(mintokens=3, maxlength=7, doskipgrams_exhaustive=True)  # the binding's lower-case keywords
    assert (o.MINTOKENS, o.MAXLENGTH, o.DOSKIPGRAMS_EXHAUSTIVE) == (3, 7, True)
    with pytest.raises((AttributeError, TypeError, KeyError)):
        cc.PatternModelOptions(no_such_option=1)
