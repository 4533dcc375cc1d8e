// ORACLE — TEST INFRASTRUCTURE ONLY. Restatement of rand_chacha::{ChaCha20Rng, ChaCha12Rng},
// rand_core::SeedableRng::seed_from_u64 (PCG32 expansion) and halo2curves `Fr::random`
// (SURVEY.md Appendix A.1). Callers upstream: halo2-base `gen_srs` (ChaCha20, seed [0;32]) and
// `gen_proof` (StdRng::seed_from_u64(0) = ChaCha12), reached from verifier/src/stark/mod.rs:543,593.
// KAT: ChaCha20 zero-key block 0 (RFC 7539 §A.1 vector 1) in tests/test_oracle_kat.py.
#pragma once
#include "field.hpp"

namespace oracle {

inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

inline void chacha_block(const uint32_t in[16], uint32_t out[16], int rounds) {
    uint32_t x[16];
    memcpy(x, in, 64);
#define QR(a, b, c, d)                       \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int i = 0; i < rounds; i += 2) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

// BlockRng<ChaChaXCore>: 64-word buffer = 4 consecutive blocks, 64-bit block counter in words 12-13,
// stream id (words 14-15) = 0.
struct ChaChaRng {
    uint32_t key[8];
    u64 counter = 0;
    int rounds;
    uint32_t buf[64];
    int index = 64;

    ChaChaRng(const uint8_t seed[32], int rounds_) : rounds(rounds_) { memcpy(key, seed, 32); }
    static ChaChaRng chacha20_from_seed(const uint8_t seed[32]) { return ChaChaRng(seed, 20); }
    // StdRng (rand 0.8) = ChaCha12; seed_from_u64 expands with PCG32
    static ChaChaRng std_rng_seed_from_u64(u64 state) {
        uint8_t seed[32];
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            memcpy(seed + 4 * i, &x, 4);
        }
        return ChaChaRng(seed, 12);
    }
    void refill() {
        static const uint32_t sigma[4] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        for (int b = 0; b < 4; ++b) {
            uint32_t st[16];
            memcpy(st, sigma, 16);
            memcpy(st + 4, key, 32);
            st[12] = (uint32_t)counter;
            st[13] = (uint32_t)(counter >> 32);
            st[14] = st[15] = 0;
            chacha_block(st, buf + 16 * b, rounds);
            ++counter;
        }
        index = 0;
    }
    uint32_t next_u32() {
        if (index >= 64) refill();
        return buf[index++];
    }
    u64 next_u64() {
        if (index < 63) {
            u64 lo = buf[index], hi = buf[index + 1];
            index += 2;
            return (hi << 32) | lo;
        } else if (index >= 64) {
            refill();
            u64 lo = buf[0], hi = buf[1];
            index = 2;
            return (hi << 32) | lo;
        } else {
            u64 lo = buf[63];
            refill();
            u64 hi = buf[0];
            index = 1;
            return (hi << 32) | lo;
        }
    }
    void fill_bytes(uint8_t* out, size_t nbytes) {  // BlockRng::fill_bytes: whole words off the stream
        for (size_t i = 0; i < nbytes; i += 4) {
            const uint32_t w = next_u32();
            memcpy(out + i, &w, nbytes - i < 4 ? nbytes - i : 4);
        }
    }
    Fr random_fr() {
        u64 w[8];
        for (int i = 0; i < 8; ++i) w[i] = next_u64();
        return Fr::from_u512(w);
    }
};

}  // namespace oracle
