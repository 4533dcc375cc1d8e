// Shared declarations of libb200zk: context, device buffers, kernel launch entry points.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "curve.cuh"

namespace b200zk {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
#define CUDA_CHECK(expr)                                                                                         \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            throw b200zk::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + \
                                    ":" + std::to_string(__LINE__));                                             \
    } while (0)

// 128-bit vectorised element access (Field is 2×uint4)
template <class C>
DEV Field<C> f_load(const Field<C>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Field<C> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class C>
DEV Field<C> f_load_ro(const Field<C>* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Field<C> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class C>
DEV void f_store(Field<C>* p, const Field<C>& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ---- transient device memory: a per-stream stack arena -------------------------------------------------------------------
// create_proof allocates and frees multi-GB transient columns in a fixed pattern. cudaMallocAsync's pool re-maps physical
// memory when that pattern fragments it (observed: random 0.5–1.2 s stalls inside a 0.2 s proof), so transient buffers
// come from one region with stack discipline instead: allocation = pointer bump, free = mark + pop while the top is
// free. Reuse is safe because every consumer runs on the owning stream. The first calls overflow into cudaMallocAsync
// while the high-water mark is learnt; the arena is (re)sized when it is empty.
struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, top = 0, high_water = 0;
    struct Block {
        size_t off, size;
        void* overflow;  // non-null: served by cudaMallocAsync because the arena was too small
        bool freed;
    };
    std::vector<Block> blocks;
    cudaStream_t stream = nullptr;

    void* alloc(size_t bytes) {
        const size_t sz = (bytes + 511) & ~(size_t)511;
        Block b{top, sz, nullptr, false};
        void* p;
        if (top + sz <= cap) {
            p = base + top;
        } else {
            cudaError_t e = cudaMallocAsync(&b.overflow, sz, stream);
            if (e != cudaSuccess) throw CudaError(std::string("arena overflow allocation failed: ") + cudaGetErrorString(e));
            p = b.overflow;
        }
        top += sz;
        if (top > high_water) high_water = top;
        blocks.push_back(b);
        return p;
    }
    void free(void* p) {
        for (size_t i = blocks.size(); i-- > 0;) {
            Block& b = blocks[i];
            if (b.freed) continue;
            void* q = b.overflow ? b.overflow : (void*)(base + b.off);
            if (q != p) continue;
            b.freed = true;
            if (b.overflow) cudaFreeAsync(b.overflow, stream);
            break;
        }
        while (!blocks.empty() && blocks.back().freed) {
            top = blocks.back().off;
            blocks.pop_back();
        }
    }
    // call between API calls: when empty and too small, grow to the learnt high-water mark (+ headroom)
    void maybe_grow() {
        if (!blocks.empty() || high_water <= cap) return;
        cudaStreamSynchronize(stream);
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        const size_t want = high_water + high_water / 16 + ((size_t)64 << 20);
        void* p = nullptr;
        if (cudaMalloc(&p, want) == cudaSuccess) {
            base = (uint8_t*)p;
            cap = want;
        } else {
            cudaGetLastError();  // stay on the pool path
            high_water = 0;
        }
    }
    void destroy() {
        if (base) cudaFree(base);
        base = nullptr;
        cap = top = 0;
        blocks.clear();
    }
};
Arena* arena_for(cudaStream_t s);
void arena_register(cudaStream_t s, Arena* a);

// ---- device buffer ------------------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    bool persistent = false;
    DevBuf() {}
    DevBuf(size_t n_, cudaStream_t s) { alloc(n_, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), stream(o.stream), persistent(o.persistent) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; stream = o.stream; persistent = o.persistent;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    // transient: from the stream's arena (stack discipline)
    void alloc(size_t n_, cudaStream_t s) {
        release();
        n = n_;
        stream = s;
        persistent = false;
        if (!n) return;
        Arena* a = arena_for(s);
        if (a) p = (T*)a->alloc(n * sizeof(T));
        else CUDA_CHECK(cudaMallocAsync((void**)&p, n * sizeof(T), s));
    }
    // long-lived (SRS, keys, tables): plain cudaMalloc
    void alloc_persistent(size_t n_, cudaStream_t s) {
        release();
        n = n_;
        stream = s;
        persistent = true;
        if (n) CUDA_CHECK(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void release() {
        if (p) {
            if (persistent) {
                cudaStreamSynchronize(stream);
                cudaFree(p);
            } else {
                Arena* a = arena_for(stream);
                if (a) a->free(p);
                else cudaFreeAsync(p, stream);
            }
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    T* get() const { return p; }
    size_t size() const { return n; }
};

// ---- NTT ---------------------------------------------------------------------------------------------
// Twiddle table: T[i] = w^i for i < 2^(log_n-1), w a primitive 2^log_n-th root of unity.
struct TwiddleTable {
    DevBuf<Fr> t;
    uint32_t log_n = 0;
    Fr omega;
};
struct NttPlan {
    const Fr* table;       // twiddle table of a 2^table_log root
    uint32_t table_log;    // table covers exponents < 2^(table_log-1)
    uint32_t log_n;        // transform size
    bool inverse;          // use w^-1 (read the table mirrored and negated)
    // fused pre/post processing (all optional)
    const Fr* pre_scale3 = nullptr;   // device ptr to 3 factors applied to input element i by (i mod 3)
    size_t in_len = 0;                // input elements at index >= in_len read as zero (0 = full)
    const Fr* post_scale3 = nullptr;  // 3 factors applied to output element i by (i mod 3) (divisor folded in)
    size_t out_len = 0;               // outputs at index >= out_len are not stored (0 = full)
    size_t in_stride = 1, in_offset = 0;  // input element i is read at in[i·in_stride + in_offset] (a strided subsequence)
};
void build_twiddle_table(Fr* table, const Fr& omega, uint32_t log_n, cudaStream_t stream);
int ntt_num_passes(uint32_t log_n);
extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by this library (bench "gpu_launches")

// Optional per-kernel-family timing with CUDA events on the launching stream (bench.py roofline section).
// Disabled by default: when off, prof_begin/prof_end are a branch on a global flag.
enum ProfId { PROF_MSM_ACCUMULATE = 0, PROF_MSM_OTHER, PROF_NTT_PASS, PROF_QUOTIENT, PROF_COUNT };
struct ProfSpan {
    int id;
    cudaEvent_t a, b;
    double work;  // units of algorithmic work of the launch: mixed additions (MSM), butterflies (NTT), extended rows (h)
};
extern bool g_prof_enabled;
extern std::vector<ProfSpan> g_prof_spans;
extern std::mutex g_prof_mu;  // contexts on several devices may record spans from different threads
void ntt_init_device();
void msm_init_device();
// returns a handle for prof_end (-1 when profiling is off)
inline int prof_begin(int id, cudaStream_t s, double work = 0) {
    if (!g_prof_enabled) return -1;
    ProfSpan sp;
    sp.id = id;
    sp.work = work;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_spans.push_back(sp);
    return (int)g_prof_spans.size() - 1;
}
inline void prof_end(int handle, cudaStream_t s) {
    if (handle < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof_spans[handle].b, s);
}

}  // namespace b200zk
