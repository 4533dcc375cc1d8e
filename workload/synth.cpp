// Synthetic circuits of the halo2-base shape: the WORKLOAD GENERATOR of bench.py and the tests (host only, its own small
// shared library libfriworkload.so — neither part of the product library nor of the oracle, so both bench arms can use it;
// the counterpart of halo2-base's `utils::testing` circuit builders). The reference's real witness — the FRI-verifier cell stream of verifier/src/stark/mod.rs:594-616 — needs the
// Rust gadgets plus plonky2/plonky2x, none of which exist here, so bench and tests prove shape- and
// distribution-faithful stand-ins: the cell stream replays the reference's measured mix
// (verifier/profile/bn254_rev.svg: 68 % Goldilocks range-check cells, 15 % 64-bit arithmetic, 17 % full-width Poseidon
// rows) with satisfied gates, lookups and copy constraints:
//   * range check of a 64-bit value (what GoldilocksChip::reduce emits through check_less_than_safe,
//     verifier/src/field/goldilocks/base.rs:448-454): 10 cells [l0 l1 B1 acc1 l2 B2 acc2 l3 B3 acc3] with gates at 0,3,6,
//     7 "shift" cells, then a second 10-cell range check — 27 advice cells, 8 limbs copied into lookup columns;
//   * Goldilocks mul_add rows: 64-bit a, b, c with d = a + b·c;
//   * full-width rows (Poseidon-BN254 S-box / mix): uniformly random Fr operands, chained d -> a.
#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <vector>

#include "../halo2-plonky2-verifier_b200/csrc/field.cuh"  // host path of the field arithmetic (no CUDA needed)

using namespace b200zk;

namespace {

// column counts of halo2-base's BaseConfig — the same layout b200zk_keygen documents
struct Shape {
    uint32_t k, A, L, F;
    uint32_t num_advice() const { return A + L; }
    uint32_t num_fixed() const { return F + 1 + A; }  // constants, lookup table, gate selectors
    uint32_t table_col() const { return F; }
    uint32_t selector_col(uint32_t c) const { return F + 1 + c; }
};

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    uint64_t below(uint64_t m) { return next() % m; }
};
Fr fr_small(uint64_t lo, uint64_t hi = 0) {  // value lo + hi·2^64 (< 2^128)
    Fr c = f_zero<FrCfg>();
    c.l[0] = (uint32_t)lo;
    c.l[1] = (uint32_t)(lo >> 32);
    c.l[2] = (uint32_t)hi;
    c.l[3] = (uint32_t)(hi >> 32);
    return f_to_mont(c);
}
Fr fr_pow2(uint32_t bits) {
    Fr c = f_zero<FrCfg>();
    c.l[bits / 32] = 1u << (bits % 32);
    return f_to_mont(c);
}
Fr fr_uniform(SplitMix& g) {
    uint32_t w[16];
    for (int i = 0; i < 8; ++i) {
        uint64_t v = g.next();
        w[2 * i] = (uint32_t)v;
        w[2 * i + 1] = (uint32_t)(v >> 32);
    }
    return f_from_u512<FrCfg>(w);
}

struct Builder {
    Shape sh;
    size_t n, usable;
    uint32_t lookup_bits, value_bits;
    uint64_t shift;  // 2^64 - p for Goldilocks p = 2^64 - 2^32 + 1 (1 when the limbs cannot hold it)
    Fr* fixed;
    Fr* advice;
    uint32_t* copies;
    size_t ncopies = 0, max_copies;
    SplitMix rng;
    // constants column bookkeeping
    size_t const_rows = 0;
    uint32_t row_zero, row_one, row_b[3], row_shift;
    std::vector<uint32_t> const_pool_rows;
    std::vector<Fr> const_pool;
    // lookup cursor
    uint32_t lk_col = 0;
    size_t lk_row = 0;

    Fr* adv(uint32_t c) { return advice + (size_t)c * n; }
    Fr* fix(uint32_t c) { return fixed + (size_t)c * n; }
    void copy(uint32_t ca, uint32_t ra, uint32_t cb, uint32_t rb) {
        if (ncopies >= max_copies) throw std::runtime_error("synth: copy buffer too small");
        uint32_t* p = copies + 4 * ncopies++;
        p[0] = ca; p[1] = ra; p[2] = cb; p[3] = rb;
    }
    uint32_t perm_fixed(uint32_t f) const { return f; }
    uint32_t perm_gate(uint32_t c) const { return sh.F + c; }
    uint32_t perm_lookup(uint32_t l) const { return sh.F + sh.A + l; }
    uint32_t add_constant(const Fr& v) {
        // constants fill column 0 first, then the next constant columns
        const uint32_t col = (uint32_t)(const_rows / usable), row = (uint32_t)(const_rows % usable);
        if (col >= sh.F) throw std::runtime_error("synth: constant columns full");
        fix(col)[row] = v;
        ++const_rows;
        return (uint32_t)(col * n + row);  // encoded (col,row)
    }
    void copy_const(uint32_t enc, uint32_t gate_col, uint32_t row) { copy(perm_fixed((uint32_t)(enc / n)), (uint32_t)(enc % n), perm_gate(gate_col), row); }
    bool lookups_left(size_t cnt) const { return sh.L && (size_t)(sh.L - lk_col) * usable - lk_row >= cnt; }
    void push_lookup(const Fr& limb, uint32_t gate_col, uint32_t row) {
        adv(sh.A + lk_col)[lk_row] = limb;
        copy(perm_lookup(lk_col), (uint32_t)lk_row, perm_gate(gate_col), row);
        if (++lk_row == usable) {
            lk_row = 0;
            ++lk_col;
        }
    }
    // 10 cells: range check of v (< 2^64, or < 2^(4·lookup_bits)) into 4 limbs of lookup_bits
    void range_check(uint32_t c, size_t r, uint64_t v_lo, uint64_t v_hi, uint32_t* out_row_acc) {
        Fr* a = adv(c);
        Fr* q = fix(sh.selector_col(c));
        unsigned __int128 v = ((unsigned __int128)v_hi << 64) | v_lo;
        const uint64_t mask = ((uint64_t)1 << lookup_bits) - 1;
        uint64_t limb[4];
        for (int i = 0; i < 4; ++i) limb[i] = (uint64_t)(v >> (i * lookup_bits)) & mask;
        unsigned __int128 acc = limb[0];
        a[r] = fr_small(limb[0]);
        push_lookup(a[r], c, (uint32_t)r);
        size_t p = r;
        for (int i = 1; i < 4; ++i) {
            // gate at p: a[p] + a[p+1]*a[p+2] = a[p+3]
            a[p + 1] = fr_small(limb[i]);
            a[p + 2] = fr_pow2(i * lookup_bits);
            acc += (unsigned __int128)limb[i] << (i * lookup_bits);
            a[p + 3] = fr_small((uint64_t)acc, (uint64_t)(acc >> 64));
            q[p] = f_one<FrCfg>();
            push_lookup(a[p + 1], c, (uint32_t)(p + 1));
            copy_const(row_b[i - 1], c, (uint32_t)(p + 2));
            p += 3;
        }
        *out_row_acc = (uint32_t)p;  // row of acc3 == v
    }
    // 27 cells: check_less_than_safe-like block on a random 64-bit value
    void range_block(uint32_t c, size_t r) {
        Fr* a = adv(c);
        Fr* q = fix(sh.selector_col(c));
        const uint64_t v = rng.next() >> (64 - value_bits);  // v + shift stays below 2^(4·lookup_bits) and 2^64
        uint32_t acc_row;
        range_check(c, r, v, 0, &acc_row);
        // shift cells: x3 = v + 1·shift ; x6 = x3 + 0·0
        size_t p = r + 10;
        a[p] = fr_small(v);
        a[p + 1] = f_one<FrCfg>();
        a[p + 2] = fr_small(shift);
        a[p + 3] = fr_small(v + shift);
        q[p] = f_one<FrCfg>();
        a[p + 4] = f_zero<FrCfg>();
        a[p + 5] = f_zero<FrCfg>();
        a[p + 6] = a[p + 3];
        q[p + 3] = f_one<FrCfg>();
        copy(perm_gate(c), acc_row, perm_gate(c), (uint32_t)p);
        copy_const(row_one, c, (uint32_t)(p + 1));
        copy_const(row_shift, c, (uint32_t)(p + 2));
        copy_const(row_zero, c, (uint32_t)(p + 4));
        uint32_t acc_row2;
        range_check(c, r + 17, v + shift, 0, &acc_row2);
        copy(perm_gate(c), (uint32_t)(p + 6), perm_gate(c), acc_row2);
    }
    // 4 cells (or 3 when chained onto the previous d): 64-bit mul_add
    size_t arith64(uint32_t c, size_t r, bool chain) {
        Fr* a = adv(c);
        Fr* q = fix(sh.selector_col(c));
        size_t p = r;
        if (chain) p = r - 1;  // a := previous d
        else a[p] = fr_small(rng.next());
        const uint64_t b = rng.next(), cc = rng.next() >> (rng.below(4) == 0 ? 63 : 0);  // some c in {0,1}
        a[p + 1] = fr_small(b);
        a[p + 2] = fr_small(cc);
        a[p + 3] = f_add(a[p], f_mul(a[p + 1], a[p + 2]));
        q[p] = f_one<FrCfg>();
        if (cc <= 1) copy_const(cc ? row_one : row_zero, c, (uint32_t)(p + 2));
        return p + 4;
    }
    size_t full_width(uint32_t c, size_t r, bool chain) {
        Fr* a = adv(c);
        Fr* q = fix(sh.selector_col(c));
        size_t p = r;
        if (chain) p = r - 1;
        else a[p] = fr_uniform(rng);
        a[p + 1] = fr_uniform(rng);
        if (rng.below(8) == 0 && !const_pool.empty()) {  // round-constant style operand
            const size_t t = rng.below(const_pool.size());
            a[p + 2] = const_pool[t];
            copy_const(const_pool_rows[t], c, (uint32_t)(p + 2));
        } else {
            a[p + 2] = fr_uniform(rng);
        }
        a[p + 3] = f_add(a[p], f_mul(a[p + 1], a[p + 2]));
        q[p] = f_one<FrCfg>();
        return p + 4;
    }
    void build() {
        // table column: i for i < 2^lookup_bits, 0 afterwards
        Fr* tab = fix(sh.table_col());
        for (size_t i = 0; i < ((size_t)1 << lookup_bits); ++i) tab[i] = fr_small(i);
        row_zero = add_constant(f_zero<FrCfg>());
        row_one = add_constant(f_one<FrCfg>());
        for (int i = 0; i < 3; ++i) row_b[i] = add_constant(fr_pow2((i + 1) * lookup_bits));
        row_shift = add_constant(fr_small(shift));
        const size_t pool = std::min<size_t>(2000, usable > 64 ? usable / 4 : 4);
        for (size_t i = 0; i < pool; ++i) {
            const_pool.push_back(fr_uniform(rng));
            const_pool_rows.push_back(add_constant(const_pool.back()));
        }
        const size_t limit = n - 9;  // halo2-base leaves 9 unusable rows
        for (uint32_t c = 0; c < sh.A; ++c) {
            size_t r = 0;
            bool can_chain = false;
            while (r + 4 <= limit) {
                const uint64_t pick = rng.below(100);
                if (pick < 68 && r + 27 <= limit && lookups_left(8)) {
                    range_block(c, r);
                    r += 27;
                    can_chain = false;
                } else if (pick < 83) {
                    r = arith64(c, r, can_chain && rng.below(2) == 0);
                    can_chain = true;
                } else {
                    r = full_width(c, r, can_chain && rng.below(2) == 0);
                    can_chain = true;
                }
            }
        }
    }
};

}  // namespace

extern "C" {

__attribute__((visibility("default"))) size_t friworkload_max_copies(uint32_t k, uint32_t A, uint32_t L, uint32_t F) {
    (void)L; (void)F;
    return (size_t)A * ((size_t)1 << k) + 64;
}

// fixed: num_fixed × 2^k, advice: (A+L) × 2^k field elements (uint64_t[4] Montgomery limbs, column-major); copies: up to
// friworkload_max_copies × {col_a,row_a,col_b,row_b}. Returns 0, -2 for bad arguments, -4 when the shape cannot be filled.
__attribute__((visibility("default"))) int friworkload_synth_circuit(uint32_t k, uint32_t A, uint32_t L, uint32_t F, uint64_t seed, uint64_t* fixed,
                                                                      uint64_t* advice, uint32_t* copies, size_t* ncopies) {
    if (!fixed || !advice || !copies || !ncopies || k < 5 || k > 26 || A == 0 || F == 0) return -2;
    try {
        Builder b;
        b.sh = Shape{k, A, L, F};
        b.n = (size_t)1 << k;
        b.usable = b.n - 9;
        b.lookup_bits = k - 1;
        b.value_bits = 4 * b.lookup_bits - 1 < 63 ? 4 * b.lookup_bits - 1 : 63;  // small k: values must fit 4 limbs
        b.shift = b.value_bits > 33 ? 0xffffffffull : 1;
        b.fixed = (Fr*)fixed;
        b.advice = (Fr*)advice;
        b.copies = copies;
        b.max_copies = friworkload_max_copies(k, A, L, F);
        b.rng.s = seed * 0x2545f4914f6cdd1dull + 0x1234567ull;
        memset(fixed, 0, sizeof(Fr) * b.n * b.sh.num_fixed());
        memset(advice, 0, sizeof(Fr) * b.n * b.sh.num_advice());
        b.build();
        *ncopies = b.ncopies;
        return 0;
    } catch (const std::exception&) {
        return -4;
    }
}

}  // extern "C"
