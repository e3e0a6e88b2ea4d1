// pattern.h -- host-side value types of the drop-in C++ front end (written from scratch for the B200 build).
//
// Mirrors the part of the reference's Pattern API that callers of PatternModel::train touch (reference
// include/pattern.h:73-354, include/common.h:41-54, include/datatypes.h:33-180): same class and method names, same
// meaning.  A Pattern is the class-encoded byte string of an n-gram or skipgram: one little-endian base-128 varint per
// token (bit 7 set on every byte but the last, src/classencoder.cpp:22-42); in a skipgram every gap token is the single
// byte 0x03 (src/pattern.cpp:886-908).  The bytes never contain a terminator here; write() appends the 0x00 the model
// file format wants (src/pattern.cpp:268-277).
#pragma once
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <ostream>
#include <string>
#include <vector>

#include "../csrc/spooky.h"

class InternalError : public std::exception {  // reference include/common.h:41-44
  public:
    const char* what() const throw() override { return "Colibri internal error"; }
};
class KeyError : public std::exception {  // reference include/common.h:46-49
  public:
    const char* what() const throw() override { return "Colibri KeyError"; }
};

enum PatternCategory { UNKNOWNPATTERN = 0, NGRAM = 1, SKIPGRAM = 2, FLEXGRAM = 3, SKIPGRAMORFLEXGRAM = 4 };

/// thrown by queries about a pattern the model does not hold (reference include/common.h)
class NoSuchPattern : public std::exception {
  public:
    const char* what() const noexcept override { return "Colibri FATAL ERROR: No such pattern"; }
};

class Pattern {
    std::string bytes_;

  public:
    static const unsigned char delimiterclass = 0, boundaryclass = 1, unknownclass = 2, skipclass = 3, flexclass = 4;  // reference include/classdecoder.h:48-52

    Pattern() {}
    Pattern(const unsigned char* dataref, const int size) : bytes_(reinterpret_cast<const char*>(dataref), (size_t)size) {}
    explicit Pattern(const std::string& raw) : bytes_(raw) {}
    /// build an n-gram from class ids
    static Pattern fromclasses(const std::vector<uint32_t>& classes) {
        std::string s;
        unsigned char buf[8];
        for (uint32_t c : classes) s.append(reinterpret_cast<char*>(buf), colibri::varint_put(buf, c));
        return Pattern(s);
    }

    const unsigned char* data() const { return reinterpret_cast<const unsigned char*>(bytes_.data()); }
    const std::string&   raw() const { return bytes_; }
    size_t               bytesize() const { return bytes_.size(); }
    /// length in tokens: every byte < 128 ends a token (reference src/pattern.cpp:74-101)
    size_t n() const {
        size_t k = 0;
        for (unsigned char c : bytes_) k += c < 128;
        return k;
    }
    size_t size() const { return n(); }
    bool   empty() const { return bytes_.empty(); }
    /// is token `index` a gap?  (a 0x03 / 0x04 byte standing alone as a token, reference src/pattern.cpp:107-137)
    bool isgap(int index) const {
        int  k        = 0;
        bool prevhigh = false;
        for (unsigned char c : bytes_) {
            if (c < 128) {
                if (k == index) return !prevhigh && (c == skipclass || c == flexclass);
                ++k;
                prevhigh = false;
            } else {
                prevhigh = true;
            }
        }
        return false;
    }
    PatternCategory category() const {  // reference src/pattern.cpp:23-47
        PatternCategory cat      = NGRAM;
        bool            prevhigh = false;
        for (unsigned char c : bytes_) {
            if (!prevhigh && c == flexclass) return FLEXGRAM;
            if (!prevhigh && c == skipclass) cat = SKIPGRAM;
            prevhigh = c >= 128;
        }
        return cat;
    }
    /// SpookyHash::Hash64 of the bytes, seed 0; the empty pattern hashes to 0 (reference src/pattern.cpp:234-238)
    size_t hash() const {
        if (bytes_.empty()) return 0;
        return (size_t)colibri::spooky_hash64(data(), (uint32_t)bytes_.size(), 0);
    }
    /// class ids of the tokens (gaps come out as 3)
    std::vector<uint32_t> tovector() const {
        std::vector<uint32_t> v;
        uint32_t              val = 0;
        int                   sh  = 0;
        for (unsigned char c : bytes_) {
            val |= (uint32_t)(c & 0x7F) << sh;
            sh += 7;
            if (c < 128) {
                v.push_back(val);
                val = 0;
                sh  = 0;
            }
        }
        return v;
    }
    void write(std::ostream& out) const {  // key bytes + end marker
        out.write(bytes_.data(), (std::streamsize)bytes_.size());
        out.put('\0');
    }
    bool operator==(const Pattern& o) const { return bytes_ == o.bytes_; }
    bool operator!=(const Pattern& o) const { return bytes_ != o.bytes_; }
    bool operator<(const Pattern& o) const { return bytes_ < o.bytes_; }
};

namespace std {
template <>
struct hash<Pattern> {
    size_t operator()(const Pattern& p) const noexcept { return p.hash(); }
};
}  // namespace std

/// position in the corpus: sentence is 1-based, token 0-based (reference include/datatypes.h:33-89)
class IndexReference {
  public:
    uint32_t sentence;
    uint16_t token;
    IndexReference() : sentence(0), token(0) {}
    explicit IndexReference(uint32_t s, uint16_t t) : sentence(s), token(t) {}
    bool operator<(const IndexReference& o) const { return sentence < o.sentence || (sentence == o.sentence && token < o.token); }
    bool operator==(const IndexReference& o) const { return sentence == o.sentence && token == o.token; }
    void write(std::ostream& out) const {
        out.write(reinterpret_cast<const char*>(&sentence), 4);
        out.write(reinterpret_cast<const char*>(&token), 2);
    }
};

/// the value of an indexed model: all positions of one pattern (reference include/datatypes.h:95-180)
class IndexedData {
  public:
    std::vector<IndexReference> data;
    unsigned int                count() const { return (unsigned int)data.size(); }
    size_t                      size() const { return data.size(); }
    typedef std::vector<IndexReference>::const_iterator const_iterator;
    const_iterator begin() const { return data.begin(); }
    const_iterator end() const { return data.end(); }
};
