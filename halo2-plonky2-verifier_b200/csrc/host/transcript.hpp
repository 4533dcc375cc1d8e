// Host side of the prover that stays on the CPU in the reference as well:
// Blake2b transcript (halo2_proofs::transcript::Blake2bWrite + Challenge255) and the rand_chacha stream behind
// Fr::random. Serial and byte-oriented — a patched halo2_proofs keeps its own Rust versions; these C++ mirrors exist
// so that the C ABI can run create_proof end to end. They share no code with oracle/.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../curve.cuh"

namespace b200zk {
namespace host {

class Blake2b512 {
   public:
    explicit Blake2b512(const char personal[16]) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                       0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        memcpy(iv_, IV, sizeof(IV));
        memcpy(h_, IV, sizeof(IV));
        h_[0] ^= 0x01010040ull;  // digest 64 bytes, fanout 1, depth 1
        uint64_t p[2];
        memcpy(p, personal, 16);
        h_[6] ^= p[0];
        h_[7] ^= p[1];
    }
    void absorb(const void* data, size_t len) {
        const uint8_t* in = (const uint8_t*)data;
        while (len) {
            if (fill_ == 128) {
                count_ += 128;
                block(false);
                fill_ = 0;
            }
            size_t take = 128 - fill_ < len ? 128 - fill_ : len;
            memcpy(buf_ + fill_, in, take);
            fill_ += take;
            in += take;
            len -= take;
        }
    }
    // digest of everything absorbed so far; the hasher itself keeps going
    void peek(uint8_t out[64]) const {
        Blake2b512 c = *this;
        c.count_ += c.fill_;
        memset(c.buf_ + c.fill_, 0, 128 - c.fill_);
        c.block(true);
        memcpy(out, c.h_, 64);
    }

   private:
    static uint64_t ror(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void block(bool last) {
        static const uint8_t S[10][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                          {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                          {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                          {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                          {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint64_t m[16], v[16];
        memcpy(m, buf_, 128);
        for (int i = 0; i < 8; ++i) {
            v[i] = h_[i];
            v[8 + i] = iv_[i];
        }
        v[12] ^= count_;
        if (last) v[14] = ~v[14];
        auto mix = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] += v[b] + x;
            v[d] = ror(v[d] ^ v[a], 32);
            v[c] += v[d];
            v[b] = ror(v[b] ^ v[c], 24);
            v[a] += v[b] + y;
            v[d] = ror(v[d] ^ v[a], 16);
            v[c] += v[d];
            v[b] = ror(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = S[r % 10];
            mix(0, 4, 8, 12, m[s[0]], m[s[1]]);
            mix(1, 5, 9, 13, m[s[2]], m[s[3]]);
            mix(2, 6, 10, 14, m[s[4]], m[s[5]]);
            mix(3, 7, 11, 15, m[s[6]], m[s[7]]);
            mix(0, 5, 10, 15, m[s[8]], m[s[9]]);
            mix(1, 6, 11, 12, m[s[10]], m[s[11]]);
            mix(2, 7, 8, 13, m[s[12]], m[s[13]]);
            mix(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h_[i] ^= v[i] ^ v[8 + i];
    }
    uint64_t h_[8], iv_[8], count_ = 0;
    uint8_t buf_[128];
    size_t fill_ = 0;
};

// canonical little-endian bytes of a field element (to_repr)
template <class C>
inline void field_to_bytes(const Field<C>& a, uint8_t out[32]) {
    Field<C> c = f_from_mont(a);
    memcpy(out, c.l, 32);
}
// halo2curves G1Affine::to_bytes: x LE, bit `sign_bit` (6 by default, [UNVERIFIED-4]) of byte 31 = y odd, the other of
// bits 6/7 = identity
inline void g1_to_bytes(const G1Affine& p, uint8_t out[32], int sign_bit = 6) {
    if (g1_is_identity(p)) {
        memset(out, 0, 32);
        out[31] |= (uint8_t)(1u << (sign_bit == 6 ? 7 : 6));
        return;
    }
    field_to_bytes(p.x, out);
    uint8_t yb[32];
    field_to_bytes(p.y, yb);
    out[31] |= (uint8_t)((yb[0] & 1) << sign_bit);
}

// Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>
class Transcript {
   public:
    explicit Transcript(int point_sign_bit = 6) : hash_("Halo2-Transcript"), sign_bit_(point_sign_bit) {}
    void common_point(const G1Affine& p) {
        if (g1_is_identity(p)) throw std::runtime_error("transcript: cannot absorb the point at infinity");
        uint8_t b[65];
        b[0] = 1;
        field_to_bytes(p.x, b + 1);
        field_to_bytes(p.y, b + 33);
        hash_.absorb(b, 65);
    }
    void common_scalar(const Fr& s) {
        uint8_t b[33];
        b[0] = 2;
        field_to_bytes(s, b + 1);
        hash_.absorb(b, 33);
    }
    Fr squeeze_challenge() {
        const uint8_t tag = 0;
        hash_.absorb(&tag, 1);
        uint8_t d[64];
        hash_.peek(d);
        uint32_t w[16];
        memcpy(w, d, 64);
        return f_from_u512<FrCfg>(w);
    }
    void write_point(const G1Affine& p) {
        common_point(p);
        uint8_t b[32];
        g1_to_bytes(p, b, sign_bit_);
        proof.insert(proof.end(), b, b + 32);
    }
    void write_scalar(const Fr& s) {
        common_scalar(s);
        uint8_t b[32];
        field_to_bytes(s, b);
        proof.insert(proof.end(), b, b + 32);
    }
    std::vector<uint8_t> proof;

   private:
    Blake2b512 hash_;
    int sign_bit_;
};

// rand_chacha BlockRng keystream: block j (16 words) = ChaCha(key, counter j). Fr::random takes 16 consecutive words
// (from_u512). `pos` is the position in 32-bit words; as long as only whole field elements are drawn it stays block
// aligned and draw j = from_u512(block j) — which is what the device-side generator (poly.cu fr_random_stream) relies on.
// fill_bytes (used only by the chunk-seeded random polynomial variant, [UNVERIFIED-3]) may leave it half a block off.
class FrRandomStream {
   public:
    uint32_t key[8];
    int rounds;
    uint64_t pos = 0;
    // external source (b200zk_create_proof_rng): the host's own RngCore::fill_bytes; the ChaCha fields are unused then and
    // every draw — including the n draws of the random polynomial — is pulled through the callback
    int (*fill_fn)(void* user, uint8_t* out, size_t nbytes) = nullptr;
    void* fill_user = nullptr;
    bool external() const { return fill_fn != nullptr; }
    static FrRandomStream from_callback(int (*fn)(void*, uint8_t*, size_t), void* user) {
        FrRandomStream r;
        memset(r.key, 0, sizeof(r.key));
        r.rounds = 0;
        r.fill_fn = fn;
        r.fill_user = user;
        return r;
    }
    void pull(uint8_t* out, size_t nbytes) {
        if (fill_fn(fill_user, out, nbytes) != 0) throw std::runtime_error("create_proof: the random source callback failed");
    }
    bool aligned() const { return pos % 16 == 0; }
    uint64_t block_index() const { return pos / 16; }

    static FrRandomStream std_rng_seed_from_u64(uint64_t state) {  // StdRng = ChaCha12, PCG32-expanded seed
        FrRandomStream r;
        r.rounds = 12;
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            const uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27), rot = (uint32_t)(state >> 59);
            r.key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
        return r;
    }
    static FrRandomStream chacha20_from_seed(const uint8_t seed[32]) {
        FrRandomStream r;
        r.rounds = 20;
        memcpy(r.key, seed, 32);
        return r;
    }
    void block(uint64_t counter, uint32_t out[16]) const {
        uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                           key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t x[16];
        memcpy(x, in, 64);
        auto rl = [](uint32_t v, int n) { return (v << n) | (v >> (32 - n)); };
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 16);
            x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 8);
            x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 7);
        };
        for (int r = 0; r < rounds; r += 2) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
    }
    // the next `count` keystream words
    void words(uint32_t* out, size_t count) {
        if (fill_fn) {
            pull((uint8_t*)out, 4 * count);
            pos += count;
            return;
        }
        uint32_t blk[16];
        uint64_t have = ~0ull;
        for (size_t i = 0; i < count; ++i, ++pos) {
            if (pos / 16 != have) {
                have = pos / 16;
                block(have, blk);
            }
            out[i] = blk[pos % 16];
        }
    }
    Fr next() {
        uint32_t w[16];
        words(w, 16);
        return f_from_u512<FrCfg>(w);
    }
    void fill_bytes32(uint8_t out[32]) {  // RngCore::fill_bytes of 32 bytes: eight whole words off the stream
        uint32_t w[8];
        words(w, 8);
        memcpy(out, w, 32);
    }
    void skip(uint64_t n) {
        if (fill_fn) {  // the draws have to be made to advance the host's generator
            uint8_t buf[4096];
            for (uint64_t left = 64 * n; left > 0;) {
                const size_t take = (size_t)std::min<uint64_t>(left, sizeof(buf));
                pull(buf, take);
                left -= take;
            }
        }
        pos += 16 * n;
    }
};

}  // namespace host
}  // namespace b200zk
