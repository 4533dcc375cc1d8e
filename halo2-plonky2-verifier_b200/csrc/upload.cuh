// Host→device upload of a PAGEABLE buffer through pinned staging chunks filled by worker threads.
// A Rust `Vec<Fr>` (what halo2's create_proof holds) is pageable: cudaMemcpyAsync from it is staged by the driver on the
// calling thread at ≈11 GB/s and blocks that thread, so a 570 MB witness cost 50 ms in front of the first kernel. Here
// WORKERS threads copy 4 MiB chunks into pinned slots and queue the DMA themselves; the caller only joins them when it
// needs the data, so kernels for the part that has already arrived can be launched meanwhile.
#pragma once
#include <atomic>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "context.cuh"

namespace b200zk {

inline bool host_pointer_is_pageable(const void* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

class StagedUpload {
public:
    static constexpr size_t CHUNK = (size_t)4 << 20;
    static constexpr int WORKERS = STAGE_WORKERS, SLOTS = STAGE_SLOTS;
    explicit StagedUpload(Context& c) : ctx(c) {}
    StagedUpload(const StagedUpload&) = delete;
    ~StagedUpload() { join_nothrow(); }
    // queue dst[0, bytes) <- src[0, bytes) on stream `st`; returns at once. With `piece` > 0 the buffer is seen as pieces of
    // `piece` bytes (columns) and on_piece(i) is called — from the worker thread that queued the last chunk overlapping
    // piece i, i.e. stream-ordered behind all of the piece's copies — so that the caller can queue follow-up work and
    // publish the piece as ready.
    void start(void* dst, const void* src, size_t bytes, cudaStream_t st, size_t piece = 0, std::function<void(size_t)> on_piece = nullptr) {
        join();
        if (bytes == 0) return;
        const size_t npieces = piece ? (bytes + piece - 1) / piece : 0;
        remaining.reset(npieces ? new std::atomic<long long>[npieces] : nullptr);
        for (size_t i = 0; i < npieces; ++i) remaining[i].store((long long)std::min(piece, bytes - i * piece));
        if (!ctx.stage_buf) {
            CUDA_CHECK(cudaHostAlloc((void**)&ctx.stage_buf, CHUNK * WORKERS * SLOTS, cudaHostAllocDefault));
            for (auto& e : ctx.stage_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
        const int device = ctx.device;
        for (int t = 0; t < WORKERS; ++t)
            threads.emplace_back([=]() {
                if (cudaSetDevice(device) != cudaSuccess) {
                    failed = 1;
                    return;
                }
                for (size_t i = t; i < nchunks; i += WORKERS) {
                    const int idx = t * SLOTS + (int)((i / WORKERS) % SLOTS);
                    uint8_t* slot = ctx.stage_buf + (size_t)idx * CHUNK;
                    const size_t off = i * CHUNK, len = std::min(CHUNK, bytes - off);
                    // the slot's previous DMA (this or an earlier upload) must have drained
                    if (cudaEventSynchronize(ctx.stage_ev[idx]) != cudaSuccess) failed = 1;
                    memcpy(slot, (const uint8_t*)src + off, len);
                    if (cudaMemcpyAsync((uint8_t*)dst + off, slot, len, cudaMemcpyHostToDevice, st) != cudaSuccess) failed = 1;
                    if (cudaEventRecord(ctx.stage_ev[idx], st) != cudaSuccess) failed = 1;
                    if (failed) return;
                    if (piece)  // pieces whose last chunk this was
                        for (size_t p = off / piece; p * piece < off + len; ++p) {
                            const size_t lo = std::max(off, p * piece), hi = std::min(off + len, (p + 1) * piece);
                            if (remaining[p].fetch_sub((long long)(hi - lo)) == (long long)(hi - lo) && on_piece) on_piece(p);
                        }
                }
            });
    }
    // every chunk has been queued on its stream (the DMA itself may still be running)
    void join() {
        join_nothrow();
        if (failed.exchange(0)) throw std::runtime_error("staged upload failed");
    }

private:
    void join_nothrow() {
        for (auto& t : threads)
            if (t.joinable()) t.join();
        threads.clear();
    }
    Context& ctx;
    std::vector<std::thread> threads;
    std::atomic<int> failed{0};
    std::unique_ptr<std::atomic<long long>[]> remaining;  // bytes of each piece not yet queued
};

}  // namespace b200zk
