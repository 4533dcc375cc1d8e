// ORACLE — TEST INFRASTRUCTURE ONLY. Restatement of halo2_proofs::transcript::{Blake2bWrite,
// Blake2bRead, Challenge255} over a from-scratch BLAKE2b
// (RFC 7693) with personalisation. KAT: Python hashlib.blake2b(digest_size=64,
// person=b"Halo2-Transcript") in tests/test_oracle_kat.py.
#pragma once
#include <stdexcept>

#include "curve.hpp"

namespace oracle {

struct Blake2b {
    u64 h[8];
    u64 t = 0;  // bytes compressed so far (messages < 2^64 bytes)
    uint8_t buf[128];
    size_t buflen = 0;
    size_t outlen;

    static const u64* iv() {
        static const u64 v[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return v;
    }
    Blake2b(size_t outlen_, const char* personal16) : outlen(outlen_) {
        for (int i = 0; i < 8; ++i) h[i] = iv()[i];
        h[0] ^= 0x01010000ull ^ (u64)outlen;  // digest length, no key, fanout = depth = 1
        if (personal16) {
            u64 p[2];
            memcpy(p, personal16, 16);
            h[6] ^= p[0];
            h[7] ^= p[1];
        }
    }
    static u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(const uint8_t block[128], bool last) {
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        u64 m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = iv()[i];
        }
        v[12] ^= t;
        if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                  \
    v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32); \
    v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = sigma[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]])
            G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
            G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]])
            G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef G
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* in, size_t len) {
        while (len > 0) {
            if (buflen == 128) {  // keep the last block buffered for finalisation
                t += 128;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = std::min(len, 128 - buflen);
            memcpy(buf + buflen, in, take);
            buflen += take;
            in += take;
            len -= take;
        }
    }
    // finalises a COPY of the state (the transcript keeps absorbing afterwards)
    void finalize(uint8_t* out) const {
        Blake2b c = *this;
        c.t += c.buflen;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, c.outlen);
    }
};

// Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>
struct TranscriptWrite {
    Blake2b state{64, "Halo2-Transcript"};
    std::vector<uint8_t> proof;

    void common_point(const G1Affine& p) {
        if (p.is_identity()) throw std::runtime_error("cannot write points at infinity to the transcript");
        uint8_t b[65];
        b[0] = 1;  // BLAKE2B_PREFIX_POINT
        p.x.to_bytes(b + 1);
        p.y.to_bytes(b + 33);
        state.update(b, 65);
    }
    void common_scalar(const Fr& s) {
        uint8_t b[33];
        b[0] = 2;  // BLAKE2B_PREFIX_SCALAR
        s.to_bytes(b + 1);
        state.update(b, 33);
    }
    Fr squeeze_challenge() {
        uint8_t pre = 0;  // BLAKE2B_PREFIX_CHALLENGE
        state.update(&pre, 1);
        uint8_t out[64];
        state.finalize(out);
        u64 w[8];
        memcpy(w, out, 64);
        return Fr::from_u512(w);
    }
    void write_point(const G1Affine& p) {
        common_point(p);
        uint8_t b[32];
        p.to_bytes(b);
        proof.insert(proof.end(), b, b + 32);
    }
    void write_scalar(const Fr& s) {
        common_scalar(s);
        uint8_t b[32];
        s.to_bytes(b);
        proof.insert(proof.end(), b, b + 32);
    }
};

// Blake2bRead
struct TranscriptRead {
    Blake2b state{64, "Halo2-Transcript"};
    const uint8_t* data;
    size_t len, pos = 0;
    TranscriptRead(const uint8_t* d, size_t l) : data(d), len(l) {}
    void common_point(const G1Affine& p) {
        uint8_t b[65];
        b[0] = 1;
        p.x.to_bytes(b + 1);
        p.y.to_bytes(b + 33);
        state.update(b, 65);
    }
    void common_scalar(const Fr& s) {
        uint8_t b[33];
        b[0] = 2;
        s.to_bytes(b + 1);
        state.update(b, 33);
    }
    Fr squeeze_challenge() {
        uint8_t pre = 0;
        state.update(&pre, 1);
        uint8_t out[64];
        state.finalize(out);
        u64 w[8];
        memcpy(w, out, 64);
        return Fr::from_u512(w);
    }
    G1Affine read_point() {
        if (pos + 32 > len) throw std::runtime_error("proof too short");
        G1Affine p;
        if (!G1Affine::from_bytes(data + pos, p)) throw std::runtime_error("invalid point encoding");
        pos += 32;
        common_point(p);
        return p;
    }
    Fr read_scalar() {
        if (pos + 32 > len) throw std::runtime_error("proof too short");
        Fr s;
        if (!Fr::from_bytes(data + pos, s)) throw std::runtime_error("invalid scalar encoding");
        pos += 32;
        common_scalar(s);
        return s;
    }
};

}  // namespace oracle
