// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of halo2_proofs::arithmetic::{parallelize,
// best_fft, eval_polynomial, kate_division} and halo2curves::msm::best_multiexp (un-vendored crates
// halo2-axiom / halo2curves-axiom; algorithms written out in SURVEY.md Appendix D.1, D.2).
// PARITY UNPINNED vs upstream; pinned by KATs (naive DFT, naive double-and-add MSM).
// The std::thread chunking mirrors upstream's rayon chunking so this doubles as the CPU baseline.
#pragma once
#include <cmath>
#include <functional>
#include <thread>

#include "curve.hpp"

namespace oracle {

inline int& num_threads_ref() {
    static int n = (int)std::max(1u, std::thread::hardware_concurrency());
    return n;
}
inline int num_threads() { return num_threads_ref(); }

// `parallelize`: contiguous chunks, one per thread. f(begin, end, thread_index)
inline void parallel_chunks(size_t n, const std::function<void(size_t, size_t, int)>& f) {
    int t = num_threads();
    if (t <= 1 || n < 1024) {
        f(0, n, 0);
        return;
    }
    size_t chunk = (n + t - 1) / t;
    std::vector<std::thread> th;
    for (int i = 0; i < t; ++i) {
        size_t b = std::min(n, (size_t)i * chunk), e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=, &f] { f(b, e, i); });
    }
    for (auto& x : th) x.join();
}

// ---- best_multiexp ------------------------------------------------------------
inline void multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t len, G1& acc) {
    std::vector<U256> reprs(len);
    for (size_t i = 0; i < len; ++i) reprs[i] = coeffs[i].to_canonical();
    size_t c = len < 4 ? 1 : len < 32 ? 3 : (size_t)std::ceil(std::log((double)len));
    size_t segments = 256 / c + 1;
    auto get_at = [&](size_t seg, const U256& r) -> size_t {
        size_t skip_bits = seg * c;
        if (skip_bits >= 256) return 0;
        size_t limb = skip_bits / 64, off = skip_bits % 64;
        u64 v = r.l[limb] >> off;
        if (off + c > 64 && limb + 1 < 4) v |= r.l[limb + 1] << (64 - off);
        return (size_t)(v & ((1ull << c) - 1));
    };
    struct Bucket {
        int kind = 0;  // 0 none, 1 affine, 2 projective
        G1Affine a;
        G1 p;
        void add_assign(const G1Affine& o) {
            if (kind == 0) {
                kind = 1;
                a = o;
            } else if (kind == 1) {
                p = G1::from_affine(a).add_affine(o);
                kind = 2;
            } else {
                p = p.add_affine(o);
            }
        }
        G1 add_to(const G1& other) const {
            if (kind == 0) return other;
            if (kind == 1) return other.add_affine(a);
            return other.add(p);
        }
    };
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t i = 0; i < c; ++i) acc = acc.dbl();
        std::vector<Bucket> buckets((1u << c) - 1);
        for (size_t i = 0; i < len; ++i) {
            size_t d = get_at(seg, reprs[i]);
            if (d != 0) buckets[d - 1].add_assign(bases[i]);
        }
        G1 running = G1::identity();
        for (size_t b = buckets.size(); b-- > 0;) {
            running = buckets[b].add_to(running);
            acc = acc.add(running);
        }
    }
}

inline G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t len) {
    size_t t = num_threads();
    if (len > t && t > 1) {
        size_t chunk = len / t;
        size_t nchunks = (len + chunk - 1) / chunk;
        std::vector<G1> results(nchunks, G1::identity());
        std::vector<std::thread> th;
        for (size_t i = 0; i < nchunks; ++i) {
            size_t b = i * chunk, e = std::min(len, b + chunk);
            th.emplace_back([=, &results] { multiexp_serial(coeffs + b, bases + b, e - b, results[i]); });
        }
        for (auto& x : th) x.join();
        G1 acc = G1::identity();
        for (auto& r : results) acc = acc.add(r);
        return acc;
    }
    G1 acc = G1::identity();
    multiexp_serial(coeffs, bases, len, acc);
    return acc;
}

inline G1 naive_msm(const Fr* coeffs, const G1Affine* bases, size_t len) {
    G1 acc = G1::identity();
    for (size_t i = 0; i < len; ++i) acc = acc.add(G1::from_affine(bases[i]).mul(coeffs[i]));
    return acc;
}

// ---- best_fft -----------------------------------------------------------------
inline uint32_t bitreverse(uint32_t n, uint32_t l) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; ++i) {
        r = (r << 1) | (n & 1);
        n >>= 1;
    }
    return r;
}

// natural order in, natural order out: out[j] = sum_i a[i] * omega^(i*j). Works for any "group"
// element type G with `G + G`, `G - G` and `G * Fr` (Fr, or G1 via the wrapper in kzg.hpp).
template <class G, class MulFn>
inline void best_fft_generic(G* a, size_t n, const Fr& omega, uint32_t log_n, MulFn mul) {
    for (size_t k = 0; k < n; ++k) {
        size_t rk = bitreverse((uint32_t)k, log_n);
        if (k < rk) std::swap(a[k], a[rk]);
    }
    std::vector<Fr> tw(std::max<size_t>(1, n / 2));
    tw[0] = Fr::one();
    for (size_t i = 1; i < n / 2; ++i) tw[i] = tw[i - 1] * omega;
    for (uint32_t s = 1; s <= log_n; ++s) {
        size_t half = (size_t)1 << (s - 1), step = n >> s;
        parallel_chunks(n / 2, [&](size_t b, size_t e, int) {
            for (size_t x = b; x < e; ++x) {
                size_t j = x & (half - 1), blk = x >> (s - 1);
                size_t lo = (blk << s) + j, hi = lo + half;
                G t = j == 0 ? a[hi] : mul(a[hi], tw[j * step]);
                a[hi] = a[lo] - t;
                a[lo] = a[lo] + t;
            }
        });
    }
}
inline void best_fft(Fr* a, const Fr& omega, uint32_t log_n) {
    best_fft_generic(a, (size_t)1 << log_n, omega, log_n, [](const Fr& x, const Fr& w) { return x * w; });
}
inline void naive_dft(const Fr* a, Fr* out, size_t n, const Fr& omega) {
    for (size_t j = 0; j < n; ++j) {
        Fr wj = omega.pow_u64(j), acc = Fr::zero(), w = Fr::one();
        for (size_t i = 0; i < n; ++i) {
            acc += a[i] * w;
            w *= wj;
        }
        out[j] = acc;
    }
}

// ---- polynomial helpers -----------------------------------------------------------------------
// `eval_polynomial`: chunked Horner, chunks recombined with powers of the point
inline Fr eval_polynomial(const Fr* poly, size_t n, const Fr& point) {
    int t = num_threads();
    if (n < 4096 || t <= 1) {
        Fr acc = Fr::zero();
        for (size_t i = n; i-- > 0;) acc = acc * point + poly[i];
        return acc;
    }
    size_t chunk = (n + t - 1) / t;
    std::vector<Fr> parts((n + chunk - 1) / chunk, Fr::zero());
    parallel_chunks(n, [&](size_t b, size_t e, int ti) {
        Fr acc = Fr::zero();
        for (size_t i = e; i-- > b;) acc = acc * point + poly[i];
        parts[ti] = acc * point.pow_u64(b);
    });
    Fr acc = Fr::zero();
    for (auto& p : parts) acc += p;
    return acc;
}

// `kate_division`: quotient of a(X) by (X - b), remainder dropped; output has n-1 coefficients
inline std::vector<Fr> kate_division(const Fr* a, size_t n, const Fr& b) {
    std::vector<Fr> q(n - 1, Fr::zero());
    Fr tmp = Fr::zero();
    for (size_t i = n - 1; i-- > 0;) {
        Fr lead = a[i + 1] + tmp;  // upstream: q[i] = a[i+1] - (-b)*tmp with b negated
        q[i] = lead;
        tmp = lead * b;
    }
    return q;
}

}  // namespace oracle
