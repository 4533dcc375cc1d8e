// NCCL over NVLink/NVSwitch for the bulk exchanges of the sharded prover: broadcast of per-column
// results (coefficient forms, extended cosets) from the rank that computed them and the all-gather of h(X) row ranges.
// The library opens the NCCL that the host process already loaded (torch's bundled libnccl.so.2) with dlopen, so there is
// no link-time dependency and no second copy. One process per GPU (torchrun): the 128-byte ncclUniqueId is bootstrapped
// through the host all-gather callback registered with b200zk_set_allgather — the callback's ONLY use. One process for all
// GPUs (b200zk_create_multi): the id is shared in memory. Every later exchange, including the small host-side ones
// (partial MSM sums, evaluations, status words), is an ncclAllGather on the context's stream.
#pragma once
#include <dlfcn.h>

#include <chrono>

#include "context.cuh"

namespace b200zk {

struct Nccl {
    typedef struct { char internal[128]; } UniqueId;
    void* lib = nullptr;
    void* comm = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int rank = 0, world = 1;
    // staging of the small host all-gathers: one device buffer [world][cap] and pinned host mirrors
    uint8_t *stage_dev = nullptr, *stage_pin = nullptr;
    size_t stage_cap = 0;

    void check(int rc, const char* what) {
        if (rc != 0) throw std::runtime_error(std::string("NCCL ") + what + " failed: " + (GetErrorString ? GetErrorString(rc) : "?"));
    }
    template <class F>
    void sym(F& f, const char* name) {
        f = (F)dlsym(lib, name);
        if (!f) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
    }
    void load() {
        if (lib) return;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process (torch) already loaded
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
        if (!lib) throw std::runtime_error("libnccl.so.2 not found (multi-GPU needs NCCL in the host process)");
        sym(GetUniqueId, "ncclGetUniqueId");
        sym(CommInitRank, "ncclCommInitRank");
        sym(Broadcast, "ncclBroadcast");
        sym(AllGather, "ncclAllGather");
        sym(Send, "ncclSend");
        sym(Recv, "ncclRecv");
        sym(GroupStart, "ncclGroupStart");
        sym(GroupEnd, "ncclGroupEnd");
        sym(CommDestroy, "ncclCommDestroy");
        sym(GetErrorString, "ncclGetErrorString");
    }
    // one process per GPU: the id travels through the host callback
    void init(Context& ctx) {
        rank = ctx.rank;
        world = ctx.world;
        load();
        if (!ctx.allgather) throw std::runtime_error("multi-GPU context without a bootstrap callback (b200zk_set_allgather) or an in-process communicator");
        UniqueId id;
        memset(&id, 0, sizeof(id));
        if (rank == 0) check(GetUniqueId(&id), "GetUniqueId");
        std::vector<UniqueId> all(world);
        if (ctx.allgather(ctx.allgather_user, &id, sizeof(id), all.data()) != 0) throw std::runtime_error("NCCL bootstrap all-gather failed");
        check(CommInitRank(&comm, world, all[0], rank), "CommInitRank");
    }
    // one process, one thread per GPU: every thread calls this with the id that rank 0 made (the calls block until all joined)
    void init_with_id(int rank_, int world_, const UniqueId& id) {
        rank = rank_;
        world = world_;
        load();
        check(CommInitRank(&comm, world, id, rank), "CommInitRank");
    }
    void ensure_stage(size_t bytes_per_rank, cudaStream_t stream) {
        const size_t need = ((bytes_per_rank + 255) & ~(size_t)255);
        if (need <= stage_cap) return;
        CUDA_CHECK(cudaStreamSynchronize(stream));
        if (stage_dev) cudaFree(stage_dev);
        if (stage_pin) cudaFreeHost(stage_pin);
        stage_dev = stage_pin = nullptr;
        stage_cap = std::max<size_t>(need, 4096);
        CUDA_CHECK(cudaMalloc((void**)&stage_dev, stage_cap * world));
        CUDA_CHECK(cudaHostAlloc((void**)&stage_pin, stage_cap * (world + 1), cudaHostAllocDefault));
    }
    ~Nccl() {
        if (stage_dev) cudaFree(stage_dev);
        if (stage_pin) cudaFreeHost(stage_pin);
        if (comm && CommDestroy) CommDestroy(comm);
    }
};

// ownership of column-wise work in the sharded prover: item i of a family belongs to rank (i + off) mod world. Each column
// family uses its own offset so that the ranks taking the odd column of 17 advice / 9 product / 9 lookup columns differ.
struct Sharder {
    Context& ctx;
    bool enabled = true;  // false: this call works on the local GPU alone even inside a multi-rank job
    explicit Sharder(Context& c) : ctx(c) {}
    // brackets one collective when tracing: local work queued before it is drained first, so the span is the collective alone
    struct CommSpan {
        Context& c;
        std::chrono::steady_clock::time_point t0;
        explicit CommSpan(Context& ctx_) : c(ctx_) {
            if (!c.comm_trace) return;
            cudaStreamSynchronize(c.stream);
            t0 = std::chrono::steady_clock::now();
        }
        ~CommSpan() {
            if (!c.comm_trace) return;
            cudaStreamSynchronize(c.stream);
            c.comm_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            ++c.comm_calls;
        }
    };
    bool on() const { return enabled && ctx.sharded(); }
    int owner(size_t i, size_t off = 0) const { return on() ? (int)((i + off) % ctx.world) : 0; }
    bool mine(size_t i, size_t off = 0) const { return !on() || owner(i, off) == ctx.rank; }
    Nccl& nccl() {
        if (!ctx.nccl) {
            auto n = std::make_shared<Nccl>();
            n->init(ctx);
            ctx.nccl = n;
        }
        return *ctx.nccl;
    }
    // All-gather of `bytes` HOST bytes from every rank into recv[world][bytes] (host): H2D of this rank's piece, one
    // ncclAllGather on the context's stream, D2H of everything, one stream synchronisation. ~30 µs over NVLink instead of a
    // round trip through the host process's own collective layer.
    void host_allgather(const void* send, size_t bytes, void* recv) {
        Nccl& nc = nccl();
        CommSpan span(ctx);
        nc.ensure_stage(bytes, ctx.stream);
        uint8_t* pin_send = nc.stage_pin + nc.stage_cap * nc.world;
        memcpy(pin_send, send, bytes);
        CUDA_CHECK(cudaMemcpyAsync(nc.stage_dev + (size_t)nc.rank * bytes, pin_send, bytes, cudaMemcpyHostToDevice, ctx.stream));
        nc.check(nc.AllGather(nc.stage_dev + (size_t)nc.rank * bytes, nc.stage_dev, bytes, 1, nc.comm, ctx.stream), "AllGather");
        CUDA_CHECK(cudaMemcpyAsync(nc.stage_pin, nc.stage_dev, bytes * nc.world, cudaMemcpyDeviceToHost, ctx.stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
        memcpy(recv, nc.stage_pin, bytes * nc.world);
    }
    // the same for a piece that is already in DEVICE memory (the partial sums of a commit batch): no upload
    void dev_to_host_allgather(const void* send_dev, size_t bytes, void* recv) {
        Nccl& nc = nccl();
        CommSpan span(ctx);
        nc.ensure_stage(bytes, ctx.stream);
        nc.check(nc.AllGather(send_dev, nc.stage_dev, bytes, 1, nc.comm, ctx.stream), "AllGather");
        CUDA_CHECK(cudaMemcpyAsync(nc.stage_pin, nc.stage_dev, bytes * nc.world, cudaMemcpyDeviceToHost, ctx.stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
        memcpy(recv, nc.stage_pin, bytes * nc.world);
    }
    void group_start() {
        if (on()) nccl().check(nccl().GroupStart(), "GroupStart");
    }
    void group_end() {
        if (on()) nccl().check(nccl().GroupEnd(), "GroupEnd");
    }
    // every rank ends up with the owner's copy of buf[0..count)
    void broadcast(Fr* buf, size_t count, int root) {
        if (!on()) return;
        CommSpan span(ctx);
        nccl().check(nccl().Broadcast(buf, buf, count * sizeof(Fr), /*ncclUint8*/ 1, root, nccl().comm, ctx.stream), "Broadcast");
    }
    // Columns base[c·len .. (c+1)·len), c < ncols, each complete on rank owner(c, off) only: afterwards complete everywhere.
    // Uneven ownership: one in-place broadcast per column from its owner, all in ONE NCCL group (they run concurrently over
    // NVSwitch) — only the payload moves and nothing is staged. Even ownership: one all-gather over an owner-major staging
    // buffer (stage[r][j] = the j-th column owned by rank r = column first(r) + j·world).
    // `st` = the stream the collective runs on (the context's stream, or its comm stream for an overlapped exchange).
    void allgather_columns_on(cudaStream_t st, Fr* base, size_t ncols, size_t len, size_t off, bool broadcasts) {
        Nccl& nc = nccl();
        if (ncols % ctx.world != 0 || broadcasts) {  // NB: the choice must not depend on anything rank-local (e.g. tracing)
            nc.check(nc.GroupStart(), "GroupStart");
            for (size_t c = 0; c < ncols; ++c)
                nc.check(nc.Broadcast(base + c * len, base + c * len, len * sizeof(Fr), 1, owner(c, off), nc.comm, st), "Broadcast");
            nc.check(nc.GroupEnd(), "GroupEnd");
            return;
        }
        const size_t world = ctx.world, per = (ncols + world - 1) / world;
        auto first = [&](size_t r) { return (r + world - off % world) % world; };  // smallest column index owned by rank r
        DevBuf<Fr> stage(world * per * len, ctx.stream);
        for (size_t c = first(ctx.rank), j = 0; c < ncols; c += world, ++j)
            CUDA_CHECK(cudaMemcpyAsync(stage.get() + ((size_t)ctx.rank * per + j) * len, base + c * len, len * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx.stream));
        nc.check(nc.AllGather(stage.get() + (size_t)ctx.rank * per * len, stage.get(), per * len * sizeof(Fr), 1, nc.comm, ctx.stream), "AllGather");
        for (size_t r = 0; r < world; ++r) {
            if (r == (size_t)ctx.rank) continue;
            for (size_t c = first(r), j = 0; c < ncols; c += world, ++j)
                CUDA_CHECK(cudaMemcpyAsync(base + c * len, stage.get() + (r * per + j) * len, len * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx.stream));
        }
    }
    // rank r holds columns [ncols·r/world, ncols·(r+1)/world) of base: afterwards every rank holds all of them
    void broadcast_blocks(Fr* base, size_t ncols, size_t len) {
        if (!on() || ncols == 0) return;
        CommSpan span(ctx);
        Nccl& nc = nccl();
        nc.check(nc.GroupStart(), "GroupStart");
        for (int r = 0; r < ctx.world; ++r) {
            const size_t lo = ncols * (size_t)r / ctx.world, hi = ncols * (size_t)(r + 1) / ctx.world;
            if (hi > lo) nc.check(nc.Broadcast(base + lo * len, base + lo * len, (hi - lo) * len * sizeof(Fr), 1, r, nc.comm, ctx.stream), "Broadcast");
        }
        nc.check(nc.GroupEnd(), "GroupEnd");
    }
    void allgather_columns(Fr* base, size_t ncols, size_t len, size_t off = 0) {
        if (!on() || ncols == 0) return;
        CommSpan span(ctx);
        allgather_columns_on(ctx.stream, base, ncols, len, off, false);
    }
    // The same exchange on the context's comm stream: it starts once everything queued on the main stream so far is done and
    // runs beside the kernels launched afterwards; the main stream must not touch the columns until async_wait(). Several
    // async exchanges may be queued before one wait. (Traced calls run it synchronously so that the comm timing stays a
    // bracketed span.)
    void allgather_columns_async(Fr* base, size_t ncols, size_t len, size_t off = 0) {
        if (!on() || ncols == 0) return;
        if (ctx.comm_trace || !ctx.comm_stream) {  // same collectives (grouped broadcasts), on the main stream
            CommSpan span(ctx);
            allgather_columns_on(ctx.stream, base, ncols, len, off, true);
            return;
        }
        CUDA_CHECK(cudaEventRecord(ctx.comm_fork, ctx.stream));
        CUDA_CHECK(cudaStreamWaitEvent(ctx.comm_stream, ctx.comm_fork, 0));
        allgather_columns_on(ctx.comm_stream, base, ncols, len, off, true);
        ctx.comm_pending = true;
    }
    void async_wait() {
        if (!ctx.comm_pending) return;
        CUDA_CHECK(cudaEventRecord(ctx.comm_done, ctx.comm_stream));
        CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, ctx.comm_done, 0));
        ctx.comm_pending = false;
    }
    // Row-slice exchange for the h(X) stage: column c (length en, complete on owner(c) only) is needed by rank d only on
    // the extended rows d evaluates plus the rotation halo — rows [d·R − before, (d+1)·R + after) mod en, R = en / world.
    // Owners send exactly those rows (1/world of the column per peer instead of a full broadcast); receivers keep
    // full-length buffers and only that window is ever read.
    template <class OwnerFn>
    void exchange_row_slices(Fr* const* cols, size_t ncols, OwnerFn owner_of, size_t en, size_t before, size_t after) {
        if (!on()) return;
        Nccl& nc = nccl();
        CommSpan span(ctx);
        const size_t R = en / ctx.world;
        auto for_segments = [&](int d, auto&& fn) {  // contiguous pieces of rank d's window
            const long long start = (long long)(R * d) - (long long)before, end = (long long)(R * (d + 1)) + (long long)after;
            if (end - start >= (long long)en) {
                fn((size_t)0, en);
                return;
            }
            if (start < 0) {
                fn((size_t)(en + start), (size_t)(-start));
                fn((size_t)0, (size_t)end);
            } else if (end > (long long)en) {
                fn((size_t)start, (size_t)(en - start));
                fn((size_t)0, (size_t)(end - en));
            } else {
                fn((size_t)start, (size_t)(end - start));
            }
        };
        nc.check(nc.GroupStart(), "GroupStart");
        for (size_t c = 0; c < ncols; ++c) {
            const int o = owner_of(c);
            if (o == ctx.rank) {
                for (int d = 0; d < ctx.world; ++d)
                    if (d != ctx.rank)
                        for_segments(d, [&](size_t off, size_t len) {
                            nc.check(nc.Send(cols[c] + off, len * sizeof(Fr), 1, d, nc.comm, ctx.stream), "Send");
                        });
            } else {
                for_segments(ctx.rank, [&](size_t off, size_t len) {
                    nc.check(nc.Recv(cols[c] + off, len * sizeof(Fr), 1, o, nc.comm, ctx.stream), "Recv");
                });
            }
        }
        nc.check(nc.GroupEnd(), "GroupEnd");
    }
    // buf holds world equal chunks; this rank filled chunk `rank`
    void all_gather_inplace(Fr* buf, size_t chunk) {
        if (!on()) return;
        CommSpan span(ctx);
        nccl().check(nccl().AllGather(buf + (size_t)ctx.rank * chunk, buf, chunk * sizeof(Fr), 1, nccl().comm, ctx.stream), "AllGather");
    }
};

}  // namespace b200zk
