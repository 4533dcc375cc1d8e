// extern "C" surface of libb200zk (include/b200zk.h): argument checking, error mapping, host<->device staging.
#include "../../include/b200zk.h"

#include "prover.cuh"
#include "collectives.cuh"
#include "upload.cuh"

using namespace b200zk;

namespace b200zk {
G1Affine msm_run(Context& ctx, const G1Affine* bases, const Fr* scalars, size_t n);
G1Affine msm_run_srs(Context& ctx, int basis, const Fr* scalars, size_t n);
void msm_batch_srs(Context& ctx, int basis, const Fr* const* cols, size_t ncols, size_t n, G1Affine* out);
void srs_setup(Context& ctx, uint32_t k, const Fr& s_trapdoor);
void srs_build_tables(Context& ctx);
size_t srs_file_size(uint32_t k, int format);
void srs_read(Context& ctx, const uint8_t* data, size_t len, int format);
size_t srs_write(Context& ctx, int format, uint8_t* out, size_t cap);
}

#include <thread>

struct b200zk_ctx {
    Context c;
    // device group (b200zk_create_multi): this is rank 0 and owns the contexts of ranks 1..N-1
    std::vector<b200zk_ctx*> peers;
    std::vector<Fr*> group_advice;  // per peer: a device buffer for the witness of b200zk_create_proof_dev (grown on demand)
    std::vector<size_t> group_advice_cap;
    b200zk_ctx* rank_ctx(int r) { return r == 0 ? this : peers[r - 1]; }
    int group_size() const { return 1 + (int)peers.size(); }
};

// One process, several GPUs: a call on the group context runs on every device in lockstep, one worker thread per device
// (the sharded prover is SPMD: every rank executes the same create_proof and meets its peers in the collectives).
// Workers re-enter the same exported function with their rank's context; tls_group_worker stops the recursion there.
static thread_local bool tls_group_worker = false;
template <class F>
static int group_call(b200zk_ctx* ctx, F&& fn) {
    const int N = ctx->group_size();
    std::vector<int> rc(N, 0);
    std::vector<std::thread> th;
    for (int r = 1; r < N; ++r)
        th.emplace_back([&, r]() {
            tls_group_worker = true;
            rc[r] = fn(ctx->rank_ctx(r), r);
        });
    tls_group_worker = true;
    rc[0] = fn(ctx, 0);
    tls_group_worker = false;
    for (auto& t : th) t.join();
    for (int r = 0; r < N; ++r)
        if (rc[r] != 0) {
            if (r != 0) ctx->c.last_error = "device group rank " + std::to_string(r) + ": " + ctx->rank_ctx(r)->c.last_error;
            return rc[r];
        }
    return 0;
}
#define GROUP_DISPATCH(ctx, expr)                                            \
    if ((ctx) && !(ctx)->peers.empty() && !tls_group_worker)                 \
        return group_call((ctx), [&](b200zk_ctx* rctx, int rank) -> int {    \
            (void)rank;                                                      \
            return (expr);                                                   \
        });
// entry points that take device pointers of ONE device (or are per-device by nature) run on rank 0 alone in a group
struct SoloGuard {
    Context* c = nullptr;
    explicit SoloGuard(b200zk_ctx* ctx) {
        if (ctx && !ctx->peers.empty() && !tls_group_worker) {
            c = &ctx->c;
            c->solo = true;
        }
    }
    ~SoloGuard() {
        if (c) c->solo = false;
    }
};

#define API_BEGIN(ctx)                      \
    if (!(ctx)) return B200ZK_EINVAL;       \
    std::lock_guard<std::mutex> _lk((ctx)->c.mu); \
    try {                                   \
        cudaSetDevice((ctx)->c.device);             \
        (ctx)->c.arena.maybe_grow();
#define API_END(ctx)                        \
    }                                       \
    catch (const CudaError& e) {            \
        (ctx)->c.last_error = e.what();     \
        return B200ZK_ECUDA;                \
    }                                       \
    catch (const SynthesisError& e) {       \
        (ctx)->c.last_error = e.what();     \
        return B200ZK_ESYNTH;               \
    }                                       \
    catch (const std::invalid_argument& e) {\
        (ctx)->c.last_error = e.what();     \
        return B200ZK_EINVAL;               \
    }                                       \
    catch (const std::exception& e) {       \
        (ctx)->c.last_error = e.what();     \
        return B200ZK_ESTATE;               \
    }                                       \
    return B200ZK_OK;

extern "C" {

int b200zk_create(int device, b200zk_ctx** out) {
    if (!out) return B200ZK_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return B200ZK_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return B200ZK_ENODEV;
    b200zk_ctx* ctx = new b200zk_ctx;
    ctx->c.device = device;
    if (cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B200ZK_ECUDA;
    }
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    for (auto& st : ctx->c.aux_streams)
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) st = nullptr;
    for (auto& e : ctx->c.msm_events) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& e : ctx->c.msm_join) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->c.msm_fork, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&ctx->c.copy_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->c.copy_stream = nullptr;
    cudaEventCreateWithFlags(&ctx->c.copy_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->c.copy_done, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&ctx->c.comm_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->c.comm_stream = nullptr;
    if (cudaStreamCreateWithFlags(&ctx->c.ntt_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->c.ntt_stream = nullptr;
    cudaEventCreateWithFlags(&ctx->c.ntt_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->c.ntt_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->c.comm_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->c.comm_done, cudaEventDisableTiming);
    if (cudaHostAlloc((void**)&ctx->c.pinned_u32, 64, cudaHostAllocDefault) != cudaSuccess) {
        b200zk_destroy(ctx);
        return B200ZK_ECUDA;
    }
    ctx->c.arena.stream = ctx->c.stream;
    arena_register(ctx->c.stream, &ctx->c.arena);
    try {  // kernel attributes are per device: every context sets them for its own
        ntt_init_device();
        msm_init_device();
    } catch (const std::exception&) {
        b200zk_destroy(ctx);
        return B200ZK_ECUDA;
    }
    *out = ctx;
    return B200ZK_OK;
}
int b200zk_create_multi(const int* devices, int ndev, b200zk_ctx** out) {
    if (!out || !devices || ndev < 1 || ndev > 64) return B200ZK_EINVAL;
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return B200ZK_EINVAL;  // one rank per GPU
    std::vector<b200zk_ctx*> ctxs(ndev, nullptr);
    int rc = B200ZK_OK;
    for (int i = 0; i < ndev && rc == B200ZK_OK; ++i) rc = b200zk_create(devices[i], &ctxs[i]);
    auto fail = [&](int code) {
        for (auto* c : ctxs)
            if (c) b200zk_destroy(c);
        return code;
    };
    if (rc != B200ZK_OK) return fail(rc);
    if (ndev == 1) {
        *out = ctxs[0];
        return B200ZK_OK;
    }
    // the communicator: the id is made here and shared in memory; ncclCommInitRank blocks until every rank has joined,
    // hence one thread per device
    Nccl::UniqueId id;
    memset(&id, 0, sizeof(id));
    try {
        Nccl boot;
        boot.load();
        boot.check(boot.GetUniqueId(&id), "GetUniqueId");
    } catch (const std::exception&) {
        return fail(B200ZK_ESTATE);
    }
    std::vector<int> ok(ndev, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r)
        th.emplace_back([&, r]() {
            try {
                if (cudaSetDevice(devices[r]) != cudaSuccess) return;
                auto n = std::make_shared<Nccl>();
                n->init_with_id(r, ndev, id);
                ctxs[r]->c.rank = r;
                ctxs[r]->c.world = ndev;
                ctxs[r]->c.nccl = n;
                ok[r] = 1;
            } catch (const std::exception& e) {
                ctxs[r]->c.last_error = e.what();
            }
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r)
        if (!ok[r]) return fail(B200ZK_ESTATE);
    ctxs[0]->peers.assign(ctxs.begin() + 1, ctxs.end());
    ctxs[0]->group_advice.assign(ndev - 1, nullptr);
    ctxs[0]->group_advice_cap.assign(ndev - 1, 0);
    *out = ctxs[0];
    return B200ZK_OK;
}
int b200zk_group_size(b200zk_ctx* ctx) { return ctx ? ctx->group_size() : 0; }
int b200zk_destroy(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_EINVAL;
    if (!ctx->peers.empty()) {
        // communicators first, all ranks at once (ncclCommDestroy may wait for its peers), then the contexts
        std::vector<std::thread> th;
        for (int r = 0; r < ctx->group_size(); ++r)
            th.emplace_back([ctx, r]() {
                b200zk_ctx* rc = ctx->rank_ctx(r);
                cudaSetDevice(rc->c.device);
                cudaStreamSynchronize(rc->c.stream);
                rc->c.nccl.reset();
            });
        for (auto& t : th) t.join();
        for (size_t i = 0; i < ctx->peers.size(); ++i) {
            if (ctx->group_advice[i]) {
                cudaSetDevice(ctx->peers[i]->c.device);
                cudaFree(ctx->group_advice[i]);
            }
            b200zk_destroy(ctx->peers[i]);
        }
        ctx->peers.clear();
    }
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    cudaStream_t s = ctx->c.stream;
    ctx->c.srs.reset();
    ctx->c.tables.clear();
    ctx->c.custom_table.reset();
    ctx->c.domains.clear();
    ctx->c.scratch.release();
    cudaStreamSynchronize(s);
    arena_register(s, nullptr);
    ctx->c.arena.destroy();
    for (auto& st : ctx->c.aux_streams)
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (auto& e : ctx->c.msm_events)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->c.msm_join)
        if (e) cudaEventDestroy(e);
    if (ctx->c.msm_fork) cudaEventDestroy(ctx->c.msm_fork);
    if (ctx->c.copy_stream) {
        cudaStreamSynchronize(ctx->c.copy_stream);
        cudaStreamDestroy(ctx->c.copy_stream);
    }
    if (ctx->c.stage_buf) cudaFreeHost(ctx->c.stage_buf);
    for (auto& e : ctx->c.stage_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->c.comm_stream) {
        cudaStreamSynchronize(ctx->c.comm_stream);
        cudaStreamDestroy(ctx->c.comm_stream);
    }
    if (ctx->c.ntt_stream) {
        cudaStreamSynchronize(ctx->c.ntt_stream);
        ctx->c.side_scratch.release();
        cudaStreamDestroy(ctx->c.ntt_stream);
    }
    if (ctx->c.ntt_fork) cudaEventDestroy(ctx->c.ntt_fork);
    if (ctx->c.ntt_done) cudaEventDestroy(ctx->c.ntt_done);
    if (ctx->c.comm_fork) cudaEventDestroy(ctx->c.comm_fork);
    if (ctx->c.comm_done) cudaEventDestroy(ctx->c.comm_done);
    for (auto& e : ctx->c.column_events) cudaEventDestroy(e);
    if (ctx->c.copy_fork) cudaEventDestroy(ctx->c.copy_fork);
    if (ctx->c.copy_done) cudaEventDestroy(ctx->c.copy_done);
    if (ctx->c.pinned_u32) cudaFreeHost(ctx->c.pinned_u32);
    delete ctx;
    cudaStreamDestroy(s);
    return B200ZK_OK;
}
const char* b200zk_last_error(b200zk_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : "null context"; }
void* b200zk_stream(b200zk_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }
int b200zk_sync(b200zk_ctx* ctx) {
    GROUP_DISPATCH(ctx, b200zk_sync(rctx))
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    API_END(ctx)
}
unsigned long long b200zk_launch_count(void) { return g_launch_count; }
int b200zk_profile_enable(b200zk_ctx* ctx, int on) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    std::lock_guard<std::mutex> plk(g_prof_mu);
    for (auto& sp : g_prof_spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    g_prof_spans.clear();
    g_prof_enabled = on != 0;
    API_END(ctx)
}
int b200zk_profile_get(b200zk_ctx* ctx, int id, double* total_ms, unsigned long long* launches) {
    API_BEGIN(ctx)
    if (!total_ms || !launches) throw std::invalid_argument("profile_get: null argument");
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    double tot = 0;
    unsigned long long cnt = 0;
    std::lock_guard<std::mutex> plk(g_prof_mu);
    for (auto& sp : g_prof_spans) {
        if (sp.id != id) continue;
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        tot += ms;
        ++cnt;
    }
    *total_ms = tot;
    *launches = cnt;
    API_END(ctx)
}
int b200zk_profile_work(b200zk_ctx* ctx, int id, double* units) {
    API_BEGIN(ctx)
    if (!units) throw std::invalid_argument("profile_work: null argument");
    double tot = 0;
    std::lock_guard<std::mutex> plk(g_prof_mu);
    for (auto& sp : g_prof_spans)
        if (sp.id == id) tot += sp.work;
    *units = tot;
    API_END(ctx)
}
int b200zk_set_allgather(b200zk_ctx* ctx, int rank, int world, b200zk_allgather_fn fn, void* user) {
    if (ctx && !ctx->peers.empty()) return B200ZK_EINVAL;  // a device group has its own communicator
    API_BEGIN(ctx)
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) throw std::invalid_argument("set_allgather: bad arguments");
    ctx->c.rank = rank;
    ctx->c.world = world;
    ctx->c.allgather = fn;
    ctx->c.allgather_user = user;
    ctx->c.nccl.reset();
    API_END(ctx)
}
int b200zk_comm_init(b200zk_ctx* ctx) {
    if (ctx && !ctx->peers.empty()) return B200ZK_OK;  // a device group is born with its communicator
    API_BEGIN(ctx)
    if (ctx->c.world > 1) {
        Sharder sh(ctx->c);
        sh.nccl();
    }
    API_END(ctx)
}
int b200zk_set_compat(b200zk_ctx* ctx, uint32_t flags, uint32_t random_poly_chunks) {
    GROUP_DISPATCH(ctx, b200zk_set_compat(rctx, flags, random_poly_chunks))
    API_BEGIN(ctx)
    if (flags & ~7u) throw std::invalid_argument("set_compat: unknown flag bits");
    Compat& c = ctx->c.compat;
    c.draw_unused_blinds = !(flags & B200ZK_COMPAT_NO_UNUSED_BLIND_DRAWS);
    c.lookup_fill_from_end = !(flags & B200ZK_COMPAT_LOOKUP_FILL_ASCENDING);
    c.point_sign_bit = (flags & B200ZK_COMPAT_POINT_SIGN_BIT7) ? 7 : 6;
    c.random_poly_chunks = random_poly_chunks;
    API_END(ctx)
}
int b200zk_set_msm_affine_rounds(b200zk_ctx* ctx, int rounds) {
    GROUP_DISPATCH(ctx, b200zk_set_msm_affine_rounds(rctx, rounds))
    API_BEGIN(ctx)
    if (rounds < 0 || rounds > 6) throw std::invalid_argument("set_msm_affine_rounds: 0..6");
    ctx->c.msm_affine_rounds = rounds;
    API_END(ctx)
}
int b200zk_set_msm_tables(b200zk_ctx* ctx, int on) {
    GROUP_DISPATCH(ctx, b200zk_set_msm_tables(rctx, on))
    API_BEGIN(ctx)
    ctx->c.msm_tables_enabled = on != 0;
    srs_build_tables(ctx->c);
    API_END(ctx)
}
int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** out) {
    API_BEGIN(ctx)
    if (!out) throw std::invalid_argument("null out");
    CUDA_CHECK(cudaMalloc(out, bytes));
    API_END(ctx)
}
int b200zk_dev_free(b200zk_ctx* ctx, void* p) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    CUDA_CHECK(cudaFree(p));
    API_END(ctx)
}
int b200zk_h2d(b200zk_ctx* ctx, void* dst, const void* src, size_t bytes) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->c.stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    API_END(ctx)
}
int b200zk_d2h(b200zk_ctx* ctx, void* dst, const void* src, size_t bytes) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->c.stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    API_END(ctx)
}

}  // extern "C"

// ---- field / curve vector ops -------------------------------------------------------------------------------
template <class C>
__global__ void field_vec_kernel(int op, const Field<C>* a, const Field<C>* b, Field<C>* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Field<C> x = f_load(a + i), y = b ? f_load(b + i) : f_zero<C>(), r;
    switch (op) {
        case 0: r = f_add(x, y); break;
        case 1: r = f_sub(x, y); break;
        case 2: r = f_mul(x, y); break;
        case 3: r = f_inv(x); break;
        case 4: r = f_neg(x); break;
        case 5: r = f_from_mont(x); break;
        default: r = f_to_mont(x); break;
    }
    f_store(out + i, r);
}
__global__ void g1_vec_kernel(int op, const G1Affine* a, const void* b, G1Affine* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = f_load(&a[i].x);
    p.y = f_load(&a[i].y);
    G1X r;
    if (op == 0) {
        const G1Affine* bb = (const G1Affine*)b;
        G1Affine q;
        q.x = f_load(&bb[i].x);
        q.y = f_load(&bb[i].y);
        r = g1x_add_affine(g1x_from_affine(p), q);
    } else if (op == 1) {
        Fr s = f_from_mont(f_load((const Fr*)b + i));
        r = g1x_mul_bits(g1x_from_affine(p), s.l, 256);
    } else {
        r = g1x_dbl(g1x_from_affine(p));
    }
    G1Affine o = g1x_to_affine(r);
    f_store(&out[i].x, o.x);
    f_store(&out[i].y, o.y);
}

extern "C" {

int b200zk_field_vec_op(b200zk_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    API_BEGIN(ctx)
    if (!a || !out || op < 0 || op > 6 || (op <= 2 && !b)) throw std::invalid_argument("field_vec_op: bad arguments");
    if (n == 0) return B200ZK_OK;
    cudaStream_t s = ctx->c.stream;
    DevBuf<Fr> da(n, s), db(b ? n : 0, s), dout(n, s);
    CUDA_CHECK(cudaMemcpyAsync(da.get(), a, 32 * n, cudaMemcpyHostToDevice, s));
    if (b) CUDA_CHECK(cudaMemcpyAsync(db.get(), b, 32 * n, cudaMemcpyHostToDevice, s));
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (field == 0)
        field_vec_kernel<FrCfg><<<blocks, 128, 0, s>>>(op, da.get(), b ? db.get() : nullptr, dout.get(), n);
    else
        field_vec_kernel<FqCfg><<<blocks, 128, 0, s>>>(op, (const Fq*)da.get(), b ? (const Fq*)db.get() : nullptr, (Fq*)dout.get(), n);
    ++g_launch_count;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(out, dout.get(), 32 * n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_END(ctx)
}

int b200zk_g1_vec_op(b200zk_ctx* ctx, int op, const b200zk_g1_affine* a, const void* b, b200zk_g1_affine* out, size_t n) {
    API_BEGIN(ctx)
    if (!a || !out || op < 0 || op > 2 || (op <= 1 && !b)) throw std::invalid_argument("g1_vec_op: bad arguments");
    if (n == 0) return B200ZK_OK;
    cudaStream_t s = ctx->c.stream;
    size_t bbytes = op == 0 ? 64 * n : op == 1 ? 32 * n : 0;
    DevBuf<G1Affine> da(n, s), dout(n, s);
    DevBuf<uint8_t> db(bbytes, s);
    CUDA_CHECK(cudaMemcpyAsync(da.get(), a, 64 * n, cudaMemcpyHostToDevice, s));
    if (bbytes) CUDA_CHECK(cudaMemcpyAsync(db.get(), b, bbytes, cudaMemcpyHostToDevice, s));
    g1_vec_kernel<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(op, da.get(), db.get(), dout.get(), n);
    ++g_launch_count;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(out, dout.get(), 64 * n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_END(ctx)
}

// ---- NTT ------------------------------------------------------------------------------------------------------
static NttPlan plan_for_omega(Context& c, uint32_t log_n, const Fr& omega) {
    if (log_n > FrConsts::S) throw std::invalid_argument("ntt: log_n exceeds the 2-adicity of Fr (28)");
    const Fr w = FrConsts::root(log_n);
    if (f_eq(w, omega)) return make_plan(c, log_n, false);
    if (f_eq(f_inv(w), omega)) return make_plan(c, log_n, true);
    // arbitrary root: private table (cached while omega and size repeat)
    if (!c.custom_table || c.custom_table->log_n != log_n || !f_eq(c.custom_table->omega, omega)) {
        auto t = std::make_unique<TwiddleTable>();
        t->log_n = log_n;
        t->omega = omega;
        t->t.alloc_persistent(log_n == 0 ? 1 : (size_t)1 << (log_n - 1), c.stream);
        build_twiddle_table(t->t.get(), omega, log_n, c.stream);
        c.custom_table = std::move(t);
    }
    NttPlan p{};
    p.table = c.custom_table->t.get();
    p.table_log = log_n;
    p.log_n = log_n;
    p.inverse = false;
    return p;
}
static Fr load_fr(const b200zk_fr* p) {
    Fr r;
    memcpy(r.l, p->l, 32);
    return r;
}

int b200zk_ntt_batch_dev(b200zk_ctx* ctx, b200zk_fr* a_dev, uint32_t log_n, const b200zk_fr* omega, uint32_t batch, size_t stride) {
    API_BEGIN(ctx)
    if (!a_dev || !omega || batch == 0) throw std::invalid_argument("ntt: bad arguments");
    Context& c = ctx->c;
    NttPlan p = plan_for_omega(c, log_n, load_fr(omega));
    const size_t n = (size_t)1 << log_n;
    Fr* scratch = ntt_num_passes(log_n) > 1 ? c.get_scratch(n * batch) : nullptr;
    ntt_run_batch(p, (Fr*)a_dev, (Fr*)a_dev, scratch, batch, stride, stride, n, c.stream);
    API_END(ctx)
}
int b200zk_ntt_dev(b200zk_ctx* ctx, b200zk_fr* a_dev, uint32_t log_n, const b200zk_fr* omega) {
    return b200zk_ntt_batch_dev(ctx, a_dev, log_n, omega, 1, 0);
}
int b200zk_ntt(b200zk_ctx* ctx, b200zk_fr* a, uint32_t log_n, const b200zk_fr* omega) {
    if (!ctx || !a || log_n > 28) return B200ZK_EINVAL;
    const size_t n = (size_t)1 << log_n;
    void* d = nullptr;
    int rc = b200zk_dev_alloc(ctx, 32 * n, &d);
    if (rc) return rc;
    rc = b200zk_h2d(ctx, d, a, 32 * n);
    if (!rc) rc = b200zk_ntt_dev(ctx, (b200zk_fr*)d, log_n, omega);
    if (!rc) rc = b200zk_d2h(ctx, a, d, 32 * n);
    b200zk_dev_free(ctx, d);
    return rc;
}

int b200zk_lagrange_to_coeff_dev(b200zk_ctx* ctx, uint32_t k, b200zk_fr* a_dev, uint32_t batch, size_t stride) {
    API_BEGIN(ctx)
    if (!a_dev || k + 2 > 28 || batch == 0) throw std::invalid_argument("lagrange_to_coeff: bad arguments");
    dev_lagrange_to_coeff(ctx->c, k, (Fr*)a_dev, batch, stride);
    API_END(ctx)
}
int b200zk_coeff_to_extended_dev(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in_dev, b200zk_fr* out_dev, uint32_t batch, size_t stride_in,
                                 size_t stride_out) {
    API_BEGIN(ctx)
    if (!in_dev || !out_dev || k + 2 > 28 || batch == 0) throw std::invalid_argument("coeff_to_extended: bad arguments");
    dev_coeff_to_extended(ctx->c, k, (const Fr*)in_dev, (Fr*)out_dev, batch, stride_in, stride_out);
    API_END(ctx)
}
int b200zk_extended_to_coeff_dev(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in_dev, b200zk_fr* out_dev) {
    API_BEGIN(ctx)
    if (!in_dev || !out_dev || k + 2 > 28) throw std::invalid_argument("extended_to_coeff: bad arguments");
    dev_extended_to_coeff(ctx->c, k, (const Fr*)in_dev, (Fr*)out_dev);
    API_END(ctx)
}
static int staged(b200zk_ctx* ctx, const void* in, size_t in_bytes, void* out, size_t out_bytes, size_t dev_bytes,
                  int (*fn)(b200zk_ctx*, void*, void*), bool separate_out) {
    void *d = nullptr, *o = nullptr;
    int rc = b200zk_dev_alloc(ctx, dev_bytes, &d);
    if (rc) return rc;
    if (separate_out) {
        rc = b200zk_dev_alloc(ctx, out_bytes, &o);
        if (rc) {
            b200zk_dev_free(ctx, d);
            return rc;
        }
    } else {
        o = d;
    }
    rc = b200zk_h2d(ctx, d, in, in_bytes);
    if (!rc) rc = fn(ctx, d, o);
    if (!rc) rc = b200zk_d2h(ctx, out, o, out_bytes);
    b200zk_dev_free(ctx, d);
    if (separate_out) b200zk_dev_free(ctx, o);
    return rc;
}
static thread_local uint32_t tl_k;
int b200zk_lagrange_to_coeff(b200zk_ctx* ctx, uint32_t k, b200zk_fr* a) {
    if (!ctx || !a || k + 2 > 28) return B200ZK_EINVAL;
    tl_k = k;
    const size_t n = (size_t)1 << k;
    return staged(ctx, a, 32 * n, a, 32 * n, 32 * n, [](b200zk_ctx* c, void* d, void*) { return b200zk_lagrange_to_coeff_dev(c, tl_k, (b200zk_fr*)d, 1, 0); },
                  false);
}
int b200zk_coeff_to_extended(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in, b200zk_fr* out) {
    if (!ctx || !in || !out || k + 2 > 28) return B200ZK_EINVAL;
    tl_k = k;
    const size_t n = (size_t)1 << k;
    return staged(ctx, in, 32 * n, out, 128 * n, 32 * n,
                  [](b200zk_ctx* c, void* d, void* o) { return b200zk_coeff_to_extended_dev(c, tl_k, (const b200zk_fr*)d, (b200zk_fr*)o, 1, 0, 0); }, true);
}
int b200zk_extended_to_coeff(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in, b200zk_fr* out) {
    if (!ctx || !in || !out || k + 2 > 28) return B200ZK_EINVAL;
    tl_k = k;
    const size_t n = (size_t)1 << k;
    return staged(ctx, in, 128 * n, out, 96 * n, 128 * n,
                  [](b200zk_ctx* c, void* d, void* o) { return b200zk_extended_to_coeff_dev(c, tl_k, (const b200zk_fr*)d, (b200zk_fr*)o); }, true);
}

}  // extern "C"

// ---- SRS + MSM ----------------------------------------------------------------------------------------------------
extern "C" {

int b200zk_srs_load(b200zk_ctx* ctx, uint32_t k, const b200zk_g1_affine* g, const b200zk_g1_affine* g_lagrange) {
    GROUP_DISPATCH(ctx, b200zk_srs_load(rctx, k, g, g_lagrange))
    API_BEGIN(ctx)
    if (!g || !g_lagrange || k > 26) throw std::invalid_argument("srs_load: bad arguments");
    Context& c = ctx->c;
    auto srs = std::make_unique<Srs>();
    srs->k = k;
    srs->n = (size_t)1 << k;
    srs->g.alloc_persistent(srs->n, c.stream);
    srs->g_lagrange.alloc_persistent(srs->n, c.stream);
    CUDA_CHECK(cudaMemcpyAsync(srs->g.get(), g, 64 * srs->n, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(srs->g_lagrange.get(), g_lagrange, 64 * srs->n, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.srs = std::move(srs);
    srs_build_tables(c);
    API_END(ctx)
}
static const G1Affine* srs_basis(Context& c, int basis, size_t n) {
    if (!c.srs) throw std::runtime_error("no SRS loaded (call b200zk_srs_load / b200zk_srs_setup first)");
    if (n > c.srs->n) throw std::invalid_argument("msm: more scalars than SRS points");
    if (basis != 0 && basis != 1) throw std::invalid_argument("msm: basis must be 0 (g) or 1 (g_lagrange)");
    return basis == 0 ? c.srs->g.get() : c.srs->g_lagrange.get();
}
int b200zk_msm_bases_dev(b200zk_ctx* ctx, const b200zk_g1_affine* bases_dev, const b200zk_fr* scalars_dev, size_t n, b200zk_g1_affine* out) {
    SoloGuard solo(ctx);
    API_BEGIN(ctx)
    if (!out || (n && (!bases_dev || !scalars_dev))) throw std::invalid_argument("msm: null argument");
    G1Affine r = msm_run(ctx->c, (const G1Affine*)bases_dev, (const Fr*)scalars_dev, n);
    memcpy(out, &r, 64);
    API_END(ctx)
}
int b200zk_msm_dev(b200zk_ctx* ctx, int basis, const b200zk_fr* scalars_dev, size_t n, b200zk_g1_affine* out) {
    SoloGuard solo(ctx);
    API_BEGIN(ctx)
    if (!out || (n && !scalars_dev)) throw std::invalid_argument("msm: null argument");
    srs_basis(ctx->c, basis, n);
    G1Affine r = msm_run_srs(ctx->c, basis, (const Fr*)scalars_dev, n);
    memcpy(out, &r, 64);
    API_END(ctx)
}
int b200zk_msm_batch_dev(b200zk_ctx* ctx, int basis, const b200zk_fr* const* cols_dev, size_t ncols, size_t n, b200zk_g1_affine* out) {
    SoloGuard solo(ctx);
    API_BEGIN(ctx)
    if (!out || (ncols && !cols_dev)) throw std::invalid_argument("msm_batch: null argument");
    srs_basis(ctx->c, basis, n);
    std::vector<G1Affine> r(ncols);
    if (ncols) msm_batch_srs(ctx->c, basis, (const Fr* const*)cols_dev, ncols, n, r.data());
    memcpy(out, r.data(), 64 * ncols);
    API_END(ctx)
}
int b200zk_msm_batch(b200zk_ctx* ctx, int basis, const b200zk_fr* const* cols, size_t ncols, size_t n, b200zk_g1_affine* out) {
    if (ctx && !ctx->peers.empty() && !tls_group_worker) {  // every rank uploads the columns and takes its share; rank 0's results are returned
        std::vector<std::vector<b200zk_g1_affine>> scratch(ctx->group_size(), std::vector<b200zk_g1_affine>(out ? ncols : 0));
        return group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int { return b200zk_msm_batch(rctx, basis, cols, ncols, n, rank == 0 ? out : scratch[rank].data()); });
    }
    API_BEGIN(ctx)
    if (!out || (ncols && !cols)) throw std::invalid_argument("msm_batch: null argument");
    Context& c = ctx->c;
    srs_basis(c, basis, n);
    // groups of up to 8 columns share one staging buffer; within a group the uploads queue on the stream ahead of the MSM
    const size_t group = 8;
    DevBuf<Fr> d(std::min(group, ncols) * n, c.stream);
    StagedUpload stager(c);
    std::vector<G1Affine> r(ncols);
    for (size_t g0 = 0; g0 < ncols; g0 += group) {
        const size_t g = std::min(group, ncols - g0);
        std::vector<const Fr*> ptrs(g);
        for (size_t i = 0; i < g; ++i) {
            if (!cols[g0 + i]) throw std::invalid_argument("msm_batch: null column");
            ptrs[i] = d.get() + i * n;
            if (n * 32 >= ((size_t)8 << 20) && host_pointer_is_pageable(cols[g0 + i])) {  // pageable column: pinned staging (upload.cuh)
                stager.start(d.get() + i * n, cols[g0 + i], 32 * n, c.stream);
                stager.join();
            } else {
                CUDA_CHECK(cudaMemcpyAsync(d.get() + i * n, cols[g0 + i], 32 * n, cudaMemcpyHostToDevice, c.stream));
            }
        }
        msm_batch_srs(c, basis, ptrs.data(), g, n, r.data() + g0);
    }
    memcpy(out, r.data(), 64 * ncols);
    API_END(ctx)
}
int b200zk_msm(b200zk_ctx* ctx, int basis, const b200zk_fr* scalars, size_t n, b200zk_g1_affine* out) {
    if (ctx && !ctx->peers.empty() && !tls_group_worker) {  // point-range shards on every device, partial sums over NCCL
        std::vector<b200zk_g1_affine> scratch(ctx->group_size());
        return group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int { return b200zk_msm(rctx, basis, scalars, n, rank == 0 ? out : &scratch[rank]); });
    }
    API_BEGIN(ctx)
    if (!out || (n && !scalars)) throw std::invalid_argument("msm: null argument");
    Context& c = ctx->c;
    srs_basis(c, basis, n);
    DevBuf<Fr> d(n, c.stream);
    if (n) CUDA_CHECK(cudaMemcpyAsync(d.get(), scalars, 32 * n, cudaMemcpyHostToDevice, c.stream));
    G1Affine r = msm_run_srs(c, basis, d.get(), n);
    memcpy(out, &r, 64);
    API_END(ctx)
}
int b200zk_msm_bases(b200zk_ctx* ctx, const b200zk_g1_affine* bases, const b200zk_fr* scalars, size_t n, b200zk_g1_affine* out) {
    if (ctx && !ctx->peers.empty() && !tls_group_worker) {
        std::vector<b200zk_g1_affine> scratch(ctx->group_size());
        return group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int { return b200zk_msm_bases(rctx, bases, scalars, n, rank == 0 ? out : &scratch[rank]); });
    }
    API_BEGIN(ctx)
    if (!out || (n && (!scalars || !bases))) throw std::invalid_argument("msm: null argument");
    Context& c = ctx->c;
    DevBuf<Fr> d(n, c.stream);
    DevBuf<G1Affine> b(n, c.stream);
    if (n) {
        CUDA_CHECK(cudaMemcpyAsync(d.get(), scalars, 32 * n, cudaMemcpyHostToDevice, c.stream));
        CUDA_CHECK(cudaMemcpyAsync(b.get(), bases, 64 * n, cudaMemcpyHostToDevice, c.stream));
    }
    G1Affine r = msm_run(c, b.get(), d.get(), n);
    memcpy(out, &r, 64);
    API_END(ctx)
}

}  // extern "C"

// ---- keygen / create_proof ------------------------------------------------------------------------------------------
struct b200zk_pk {
    std::unique_ptr<ProvingKeyDev> pk;
    std::vector<b200zk_pk*> peers;  // device group: the keys of ranks 1..N-1 (every rank holds the whole key)
    const b200zk_pk* rank_pk(int r) const { return r == 0 ? this : peers[r - 1]; }
};

extern "C" {

int b200zk_keygen(b200zk_ctx* ctx, uint32_t k, uint32_t A, uint32_t L, uint32_t F, const b200zk_fr* fixed, const uint32_t* copies, size_t ncopies,
                  b200zk_pk** out) {
    if (ctx && !ctx->peers.empty() && !tls_group_worker) {
        if (!out) return B200ZK_EINVAL;
        std::vector<b200zk_pk*> pks(ctx->group_size(), nullptr);
        const int rc = group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int { return b200zk_keygen(rctx, k, A, L, F, fixed, copies, ncopies, &pks[rank]); });
        if (rc != B200ZK_OK) {
            for (int r = 0; r < ctx->group_size(); ++r)
                if (pks[r]) b200zk_pk_free(ctx->rank_ctx(r), pks[r]);
            return rc;
        }
        pks[0]->peers.assign(pks.begin() + 1, pks.end());
        *out = pks[0];
        return B200ZK_OK;
    }
    API_BEGIN(ctx)
    if (!fixed || !out || (ncopies && !copies)) throw std::invalid_argument("keygen: null argument");
    Shape sh{k, A, L, F};
    auto pk = keygen(ctx->c, sh, (const Fr*)fixed, copies, ncopies);
    *out = new b200zk_pk{std::move(pk)};
    API_END(ctx)
}
int b200zk_pk_free(b200zk_ctx* ctx, b200zk_pk* pk) {
    if (ctx && pk && !pk->peers.empty() && !tls_group_worker) {
        if (pk->peers.size() != ctx->peers.size()) return B200ZK_EINVAL;
        std::vector<b200zk_pk*> peers;
        peers.swap(pk->peers);
        for (size_t i = 0; i < peers.size(); ++i) b200zk_pk_free(ctx->peers[i], peers[i]);
    }
    API_BEGIN(ctx)
    delete pk;
    API_END(ctx)
}
static_assert(sizeof(b200zk_value_source) == sizeof(GateSrc) && sizeof(b200zk_calculation) == sizeof(GateCalc), "gate program ABI");
int b200zk_pk_set_gates(b200zk_ctx* ctx, b200zk_pk* pk, const b200zk_calculation* calcs, size_t ncalcs, const b200zk_fr* constants,
                        size_t nconstants, const uint32_t* results, size_t nresults) {
    if (ctx && pk && !pk->peers.empty() && !tls_group_worker) {
        if (pk->peers.size() != ctx->peers.size()) return B200ZK_EINVAL;
        return group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int {
            return b200zk_pk_set_gates(rctx, const_cast<b200zk_pk*>(pk->rank_pk(rank)), calcs, ncalcs, constants, nconstants, results, nresults);
        });
    }
    API_BEGIN(ctx)
    if (!pk) throw std::invalid_argument("null pk");
    pk_set_gates(ctx->c, *pk->pk, (const GateCalc*)calcs, ncalcs, (const Fr*)constants, nconstants, results, nresults);
    API_END(ctx)
}
int b200zk_pk_commitments(b200zk_ctx* ctx, const b200zk_pk* pk, b200zk_g1_affine* fixed_out, b200zk_g1_affine* perm_out) {
    API_BEGIN(ctx)
    if (!pk) throw std::invalid_argument("null pk");
    if (fixed_out) memcpy(fixed_out, pk->pk->fixed_commitments.data(), 64 * pk->pk->fixed_commitments.size());
    if (perm_out) memcpy(perm_out, pk->pk->perm_commitments.data(), 64 * pk->pk->perm_commitments.size());
    API_END(ctx)
}
int b200zk_pk_transcript_repr(b200zk_ctx* ctx, b200zk_pk* pk, b200zk_fr* get_out, const b200zk_fr* set_in) {
    API_BEGIN(ctx)
    if (!pk) throw std::invalid_argument("null pk");
    if (set_in) memcpy(pk->pk->transcript_repr.l, set_in->l, 32);
    if (get_out) memcpy(get_out->l, pk->pk->transcript_repr.l, 32);
    API_END(ctx)
}
int b200zk_pk_get_column(b200zk_ctx* ctx, const b200zk_pk* pk, int which, uint32_t idx, b200zk_fr* out) {
    API_BEGIN(ctx)
    if (!pk || !out) throw std::invalid_argument("null argument");
    const ProvingKeyDev& p = *pk->pk;
    const size_t n = p.shape.n(), en = 4 * n;
    const Fr* src = nullptr;
    size_t len = 0;
    switch (which) {
        case 0: if (idx < p.shape.num_perm()) { src = p.sigma_values.get() + (size_t)idx * n; len = n; } break;
        case 1: if (idx < p.shape.num_fixed()) { src = p.fixed_cosets.get() + (size_t)idx * en; len = en; } break;
        case 2: if (idx < p.shape.num_perm()) { src = p.sigma_cosets.get() + (size_t)idx * en; len = en; } break;
        case 3: if (idx < 3) { src = p.l_polys.get() + (size_t)idx * en; len = en; } break;
    }
    if (!src) throw std::invalid_argument("pk_get_column: bad selector");
    CUDA_CHECK(cudaMemcpyAsync(out, src, len * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->c.stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    API_END(ctx)
}
int b200zk_srs_setup_trapdoor(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* s) {
    GROUP_DISPATCH(ctx, b200zk_srs_setup_trapdoor(rctx, k, s))
    API_BEGIN(ctx)
    if (!s) throw std::invalid_argument("srs_setup: null trapdoor");
    srs_setup(ctx->c, k, load_fr(s));
    API_END(ctx)
}
int b200zk_srs_setup(b200zk_ctx* ctx, uint32_t k, const uint8_t seed[32], b200zk_fr* trapdoor_out) {
    GROUP_DISPATCH(ctx, b200zk_srs_setup(rctx, k, seed, rank == 0 ? trapdoor_out : nullptr))
    API_BEGIN(ctx)
    if (!seed) throw std::invalid_argument("srs_setup: null seed");
    host::FrRandomStream rng = host::FrRandomStream::chacha20_from_seed(seed);
    const Fr s = rng.next();
    if (trapdoor_out) memcpy(trapdoor_out->l, s.l, 32);
    srs_setup(ctx->c, k, s);
    API_END(ctx)
}
size_t b200zk_srs_file_size(uint32_t k, int format) { return srs_file_size(k, format); }
int b200zk_srs_read(b200zk_ctx* ctx, const uint8_t* data, size_t len, int format) {
    GROUP_DISPATCH(ctx, b200zk_srs_read(rctx, data, len, format))
    API_BEGIN(ctx)
    if (!data) throw std::invalid_argument("srs_read: null data");
    srs_read(ctx->c, data, len, format);
    API_END(ctx)
}
int b200zk_srs_write(b200zk_ctx* ctx, int format, uint8_t* out, size_t capacity, size_t* written) {
    API_BEGIN(ctx)
    if (!out || !written) throw std::invalid_argument("srs_write: null argument");
    *written = srs_write(ctx->c, format, out, capacity);
    API_END(ctx)
}
int b200zk_srs_download(b200zk_ctx* ctx, b200zk_g1_affine* g, b200zk_g1_affine* g_lagrange) {
    API_BEGIN(ctx)
    Context& c = ctx->c;
    if (!c.srs) throw std::runtime_error("no SRS loaded");
    if (g) CUDA_CHECK(cudaMemcpyAsync(g, c.srs->g.get(), 64 * c.srs->n, cudaMemcpyDeviceToHost, c.stream));
    if (g_lagrange) CUDA_CHECK(cudaMemcpyAsync(g_lagrange, c.srs->g_lagrange.get(), 64 * c.srs->n, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    API_END(ctx)
}
size_t b200zk_proof_size(uint32_t k, uint32_t A, uint32_t L, uint32_t F) { return Shape{k, A, L, F}.proof_size(); }
static int create_proof_impl(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice, bool on_device, uint64_t rng_seed, uint8_t* proof_out,
                             size_t* proof_len, double* timings) {
    if (ctx && !ctx->peers.empty() && !tls_group_worker) {
        // device group: every rank proves in lockstep (the sharded prover); rank 0's proof is returned — all ranks produce
        // the same bytes. A device-resident witness lives on rank 0's GPU and is copied to the peers over NVLink first.
        if (!pk || !advice || !proof_out || !proof_len || pk->peers.size() != ctx->peers.size()) return B200ZK_EINVAL;
        const int N = ctx->group_size();
        const size_t bytes = (size_t)pk->pk->shape.num_advice() * pk->pk->shape.n() * sizeof(Fr), psize = pk->pk->shape.proof_size();
        std::vector<std::vector<uint8_t>> proofs(N, std::vector<uint8_t>(psize));
        std::vector<size_t> lens(N, 0);
        std::vector<const b200zk_fr*> adv(N, advice);
        if (on_device)
            for (int r = 1; r < N; ++r) {
                b200zk_ctx* rc = ctx->rank_ctx(r);
                if (ctx->group_advice_cap[r - 1] < bytes) {
                    cudaSetDevice(rc->c.device);
                    if (ctx->group_advice[r - 1]) cudaFree(ctx->group_advice[r - 1]);
                    ctx->group_advice[r - 1] = nullptr;
                    ctx->group_advice_cap[r - 1] = 0;
                    if (cudaMalloc((void**)&ctx->group_advice[r - 1], bytes) != cudaSuccess) return B200ZK_ECUDA;
                    ctx->group_advice_cap[r - 1] = bytes;
                }
                if (cudaMemcpyPeer(ctx->group_advice[r - 1], rc->c.device, advice, ctx->c.device, bytes) != cudaSuccess) return B200ZK_ECUDA;
                adv[r] = (const b200zk_fr*)ctx->group_advice[r - 1];
            }
        std::vector<std::vector<double>> tms(N, std::vector<double>(10, 0.0));  // every rank runs in the same (timed or untimed) mode
        return group_call(ctx, [&](b200zk_ctx* rctx, int rank) -> int {
            return create_proof_impl(rctx, pk->rank_pk(rank), adv[rank], on_device, rng_seed, rank == 0 ? proof_out : proofs[rank].data(),
                                     rank == 0 ? proof_len : &lens[rank], !timings ? nullptr : rank == 0 ? timings : tms[rank].data());
        });
    }
    API_BEGIN(ctx)
    if (!pk || !advice || !proof_out || !proof_len) throw std::invalid_argument("create_proof: null argument");
    host::FrRandomStream rng = host::FrRandomStream::std_rng_seed_from_u64(rng_seed);
    ProofTimings tm;
    std::vector<uint8_t> proof = create_proof(ctx->c, *pk->pk, (const Fr*)advice, on_device, rng, timings ? &tm : nullptr);
    memcpy(proof_out, proof.data(), proof.size());
    *proof_len = proof.size();
    if (timings) {
        const double t[10] = {tm.upload, tm.msm, tm.ntt, tm.lookup, tm.products, tm.quotient, tm.evals, tm.shplonk, tm.other, tm.comm};
        memcpy(timings, t, sizeof(t));
    }
    API_END(ctx)
}
int b200zk_evaluate_h(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice_coeff, const b200zk_fr* perm_z_coeff, const b200zk_fr* lookup_coeff,
                      const b200zk_fr* y, const b200zk_fr* beta, const b200zk_fr* gamma, b200zk_fr* h_ext_out) {
    API_BEGIN(ctx)
    if (!pk || !advice_coeff || !perm_z_coeff || !y || !beta || !gamma || !h_ext_out) throw std::invalid_argument("evaluate_h: null argument");
    if (pk->pk->shape.L && !lookup_coeff) throw std::invalid_argument("evaluate_h: lookup_coeff is NULL but the shape has lookups");
    Fr ch[3];
    memcpy(&ch[0], y, 32);
    memcpy(&ch[1], beta, 32);
    memcpy(&ch[2], gamma, 32);
    evaluate_h(ctx->c, *pk->pk, (const Fr*)advice_coeff, (const Fr*)perm_z_coeff, (const Fr*)lookup_coeff, ch[0], ch[1], ch[2], (Fr*)h_ext_out);
    API_END(ctx)
}
uint32_t b200zk_num_sets(uint32_t A, uint32_t L, uint32_t F) {
    Shape sh{1, A, L, F};
    return sh.num_sets();
}
int b200zk_create_proof_rng(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice, b200zk_rng_fill_fn fill, void* user, uint8_t* proof_out,
                            size_t* proof_len, double* timings) {
    if (ctx && !ctx->peers.empty()) return B200ZK_EINVAL;  // one generator cannot feed several ranks in lockstep
    API_BEGIN(ctx)
    if (!pk || !advice || !proof_out || !proof_len || !fill) throw std::invalid_argument("create_proof_rng: null argument");
    host::FrRandomStream rng = host::FrRandomStream::from_callback(fill, user);
    ProofTimings tm;
    std::vector<uint8_t> proof = create_proof(ctx->c, *pk->pk, (const Fr*)advice, false, rng, timings ? &tm : nullptr);
    memcpy(proof_out, proof.data(), proof.size());
    *proof_len = proof.size();
    if (timings) {
        const double t[10] = {tm.upload, tm.msm, tm.ntt, tm.lookup, tm.products, tm.quotient, tm.evals, tm.shplonk, tm.other, tm.comm};
        memcpy(timings, t, sizeof(t));
    }
    API_END(ctx)
}
int b200zk_create_proof(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice, uint64_t rng_seed, uint8_t* proof_out, size_t* proof_len,
                        double* timings) {
    return create_proof_impl(ctx, pk, advice, false, rng_seed, proof_out, proof_len, timings);
}
int b200zk_create_proof_dev(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice_dev, uint64_t rng_seed, uint8_t* proof_out,
                            size_t* proof_len, double* timings) {
    return create_proof_impl(ctx, pk, advice_dev, true, rng_seed, proof_out, proof_len, timings);
}

}  // extern "C"

// ---- host-only helpers ------------------------------------------------------------------------------------------------
extern "C" {

int b200zk_g1_sum_host(const b200zk_g1_affine* points, size_t n, b200zk_g1_affine* out) {
    if (!out || (n && !points)) return B200ZK_EINVAL;
    G1X acc = g1x_identity();
    for (size_t i = 0; i < n; ++i) {
        G1Affine p;
        memcpy(&p, points + i, 64);
        acc = g1x_add_affine(acc, p);
    }
    const G1Affine r = g1x_to_affine(acc);
    memcpy(out, &r, 64);
    return B200ZK_OK;
}

int b200zk_host_selftest(uint64_t seed, size_t iters) {
    uint64_t st = seed * 0x9e3779b97f4a7c15ull + 1;
    auto next = [&]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return st;
    };
    auto rnd = [&]() {
        uint32_t w[16];
        for (int i = 0; i < 8; ++i) {
            uint64_t v = next();
            w[2 * i] = (uint32_t)v;
            w[2 * i + 1] = (uint32_t)(v >> 32);
        }
        return w[0] & 1 ? f_from_u512<FrCfg>(w) : f_from_u512<FrCfg>(w);
    };
    for (size_t it = 0; it < iters; ++it) {
        const Fr a = rnd(), b = rnd(), c = rnd();
        if (!f_eq(f_mul_chains(a, b), f_mul_host64(a, b))) return -10;                       // the multipliers agree
        if (!f_eq(f_mul_comba(a, b), f_mul_host64(a, b))) return -16;
        if (!f_eq(f_sqr_comba(a), f_mul_host64(a, a))) return -17;
        if (!f_eq(f_mul2_add(a, b, c, a), f_add(f_mul_host64(a, b), f_mul_host64(c, a)))) return -19;  // fused dual product
        if (!f_eq(f_mul(a, f_add(b, c)), f_add(f_mul(a, b), f_mul(a, c)))) return -11;       // distributivity
        if (!f_eq(f_sub(f_add(a, b), b), a)) return -12;
        if (!f_is_zero(a) && !f_eq(f_mul(a, f_inv(a)), f_one<FrCfg>())) return -13;
        if (!f_eq(f_to_mont(f_from_mont(a)), a)) return -14;
        Fq x, y;
        memcpy(x.l, a.l, 32);
        memcpy(y.l, b.l, 32);
        x.l[7] &= 0x0fffffffu;
        y.l[7] &= 0x0fffffffu;
        if (!f_eq(f_mul_chains(x, y), f_mul_host64(x, y))) return -15;
        if (!f_eq(f_mul_comba(x, y), f_mul_host64(x, y)) || !f_eq(f_sqr_comba(x), f_mul_host64(x, x))) return -18;
    }
    // group laws on the generator: (2G + G) + G == 2·(2G), 5G - 5G == 0, mixed add == full add
    G1Affine gen;
    gen.x = f_to_mont(Fq{{1, 0, 0, 0, 0, 0, 0, 0}});
    gen.y = f_to_mont(Fq{{2, 0, 0, 0, 0, 0, 0, 0}});
    const G1X g = g1x_from_affine(gen), g2 = g1x_dbl(g), g3 = g1x_add_affine(g2, gen), g4a = g1x_add(g3, g), g4b = g1x_dbl(g2);
    const G1Affine a4 = g1x_to_affine(g4a), b4 = g1x_to_affine(g4b);
    if (!f_eq(a4.x, b4.x) || !f_eq(a4.y, b4.y)) return -20;
    if (!g1_is_identity(g1x_add(g4a, g1x_neg(g4b)))) return -21;
    if (!g1_is_identity(g1x_add_affine(g1x_neg(g), gen))) return -22;
    const uint32_t five = 5;
    const G1Affine m5 = g1x_to_affine(g1x_mul_bits(g, &five, 3)), s5 = g1x_to_affine(g1x_add_affine(g4a, gen));
    if (!f_eq(m5.x, s5.x) || !f_eq(m5.y, s5.y)) return -23;
    // on-curve: y^2 = x^3 + 3
    const Fq three = f_to_mont(Fq{{3, 0, 0, 0, 0, 0, 0, 0}});
    if (!f_eq(f_sqr(m5.y), f_add(f_mul(f_sqr(m5.x), m5.x), three))) return -24;
    return 0;
}

}  // extern "C"

// ---- column primitives (rows E, F, I) ------------------------------------------------------------------------------------
extern "C" {

int b200zk_batch_invert(b200zk_ctx* ctx, b200zk_fr* a, size_t n) {
    API_BEGIN(ctx)
    if (n && !a) throw std::invalid_argument("batch_invert: null argument");
    cudaStream_t s = ctx->c.stream;
    DevBuf<Fr> d(n, s);
    if (n) {
        CUDA_CHECK(cudaMemcpyAsync(d.get(), a, 32 * n, cudaMemcpyHostToDevice, s));
        fr_batch_invert(d.get(), n, s);
        CUDA_CHECK(cudaMemcpyAsync(a, d.get(), 32 * n, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    API_END(ctx)
}
int b200zk_prefix_product(b200zk_ctx* ctx, const b200zk_fr* m, const b200zk_fr* first, b200zk_fr* z, size_t n) {
    API_BEGIN(ctx)
    if (!first || (n && (!m || !z))) throw std::invalid_argument("prefix_product: null argument");
    cudaStream_t s = ctx->c.stream;
    DevBuf<Fr> dm(n, s), dz(n, s);
    if (n) {
        CUDA_CHECK(cudaMemcpyAsync(dm.get(), m, 32 * n, cudaMemcpyHostToDevice, s));
        fr_prefix_product(dz.get(), dm.get(), load_fr(first), n, s);
        CUDA_CHECK(cudaMemcpyAsync(z, dz.get(), 32 * n, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    API_END(ctx)
}
int b200zk_eval_polynomial(b200zk_ctx* ctx, const b200zk_fr* poly, size_t n, const b200zk_fr* point, b200zk_fr* out) {
    API_BEGIN(ctx)
    if (!point || !out || (n && !poly)) throw std::invalid_argument("eval_polynomial: null argument");
    cudaStream_t s = ctx->c.stream;
    Fr r = f_zero<FrCfg>();
    if (n) {
        DevBuf<Fr> d(n, s);
        CUDA_CHECK(cudaMemcpyAsync(d.get(), poly, 32 * n, cudaMemcpyHostToDevice, s));
        std::vector<const Fr*> ps = {d.get()};
        fr_eval_many(ctx->c, ps, n, load_fr(point), &r);
    }
    memcpy(out->l, r.l, 32);
    API_END(ctx)
}
int b200zk_kate_division(b200zk_ctx* ctx, const b200zk_fr* a, size_t n, const b200zk_fr* b, b200zk_fr* q) {
    API_BEGIN(ctx)
    if (!a || !b || !q || n < 2) throw std::invalid_argument("kate_division: bad arguments");
    cudaStream_t s = ctx->c.stream;
    DevBuf<Fr> da(n, s), dq(n, s);
    CUDA_CHECK(cudaMemcpyAsync(da.get(), a, 32 * n, cudaMemcpyHostToDevice, s));
    fr_kate_division(ctx->c, da.get(), dq.get(), n, load_fr(b));
    CUDA_CHECK(cudaMemcpyAsync(q, dq.get(), 32 * (n - 1), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_END(ctx)
}
int b200zk_permute_expression_pair(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* input, const b200zk_fr* table, b200zk_fr* a_out, b200zk_fr* s_out) {
    API_BEGIN(ctx)
    if (!input || !table || !a_out || !s_out || k < 4 || k > 26) throw std::invalid_argument("permute_expression_pair: bad arguments");
    cudaStream_t s = ctx->c.stream;
    const size_t n = (size_t)1 << k, u = n - (Shape::blinding_factors + 1);
    DevBuf<Fr> din(n, s), dtab(n, s), da(n, s), ds(n, s);
    CUDA_CHECK(cudaMemcpyAsync(din.get(), input, 32 * n, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(dtab.get(), table, 32 * n, cudaMemcpyHostToDevice, s));
    const int st = lookup_permute(ctx->c, din.get(), dtab.get(), da.get(), ds.get(), n, u, ctx->c.compat.lookup_fill_from_end);
    if (st & LOOKUP_UNSUPPORTED) throw std::invalid_argument("permute_expression_pair: unsupported table — every table value must be < 2^k (range-style tables)");
    if (st & LOOKUP_NOT_IN_TABLE) throw SynthesisError("ConstraintSystemFailure: lookup input not in table");
    CUDA_CHECK(cudaMemcpyAsync(a_out, da.get(), 32 * u, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(s_out, ds.get(), 32 * u, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_END(ctx)
}

}  // extern "C"
