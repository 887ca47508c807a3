// Host -> device ingest of PAGEABLE memory (replaces the reference's per-batch
// zero-copy slicing of a host table, table_batch_reader.cpp:5-16, as the way rows reach the operators).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one internal bounce buffer on
// the calling thread: ~6-10 GB/s, a fifth of what the PCIe Gen5 link carries.  A pyarrow.Table that a
// user builds from NumPy / pandas / a CSV file is pageable.  Here the copy is cut into pieces (option INGEST_PIECE_KB, default 2 MB);
// a small pool of worker threads copies each piece into a package-owned PINNED slot (two per worker)
// and queues the DMA of that slot on the caller's stream; a slot is refilled only after the event
// recorded behind its DMA has fired.  The CPU copy of piece k+1 overlaps the DMA of piece k, so the
// link stays busy as long as the workers together out-run it.  The call returns when every piece has
// been QUEUED on the stream (stream order then covers the kernels the caller launches next); it does
// not wait for the DMAs.
#include "vk_common.cuh"
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace vk {
namespace {

constexpr size_t MAX_PIECE_BYTES = 8u << 20;   // slot size; the piece actually used is option INGEST_PIECE_KB
constexpr int SLOTS_PER_WORKER = 2;

// Piece size: small enough that the workers' slots together stay in the host's last-level cache (the DMA
// engine then reads a piece from cache instead of DRAM: one pass over host memory per byte instead of
// three), large enough that the per-copy driver overhead (a few microseconds) stays small.
size_t piece_bytes() {
    int64_t kb = opt(OPT_INGEST_PIECE_KB);
    if (kb < 64) kb = 64;
    if (kb > (int64_t) (MAX_PIECE_BYTES >> 10)) kb = (int64_t) (MAX_PIECE_BYTES >> 10);
    return (size_t) kb << 10;
}

struct Job {
    std::mutex mu;
    std::condition_variable cv;
    int remaining = 0;
    cudaError_t error = cudaSuccess;
};
struct Piece {
    uint8_t* dst;
    const uint8_t* src;
    size_t bytes;
    cudaStream_t stream;
    int device;
    Job* job;
};
struct Slot {
    uint8_t* host = nullptr;
    cudaEvent_t ev = nullptr;
    int ev_device = -1;
    bool busy = false;
};

class IngestPool {
public:
    explicit IngestPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    // never destroyed (the process may exit with CUDA already torn down): leaked singleton
    void submit(const Piece& p) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(p);
        }
        cv_.notify_one();
    }
    int size() const { return (int) workers_.size(); }

private:
    void run() {
        Slot slots[SLOTS_PER_WORKER];
        int next = 0, cur_device = -1;
        for (;;) {
            Piece p;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                p = q_.front();
                q_.pop_front();
            }
            cudaError_t e = cudaSuccess;
            if (p.device != cur_device) {
                e = cudaSetDevice(p.device);
                cur_device = p.device;
            }
            Slot& s = slots[next];
            next = (next + 1) % SLOTS_PER_WORKER;
            if (e == cudaSuccess && s.host == nullptr) e = cudaHostAlloc((void**) &s.host, MAX_PIECE_BYTES, cudaHostAllocPortable);
            if (e == cudaSuccess && s.busy) e = cudaEventSynchronize(s.ev);   // the slot's previous DMA has read it
            if (e == cudaSuccess && s.ev_device != p.device) {
                // (the old event, if any, has fired: it was just waited for)
                if (s.ev) cudaEventDestroy(s.ev);
                e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming);
                s.ev_device = p.device;
            }
            if (e == cudaSuccess) {
                memcpy(s.host, p.src, p.bytes);
                e = cudaMemcpyAsync(p.dst, s.host, p.bytes, cudaMemcpyHostToDevice, p.stream);
            }
            if (e == cudaSuccess) {
                e = cudaEventRecord(s.ev, p.stream);
                s.busy = true;
            }
            {
                std::lock_guard<std::mutex> lk(p.job->mu);
                if (e != cudaSuccess && p.job->error == cudaSuccess) p.job->error = e;
                if (--p.job->remaining == 0) p.job->cv.notify_all();
            }
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Piece> q_;
    std::vector<std::thread> workers_;
};

std::mutex g_pool_mu;
IngestPool* g_pool = nullptr;

IngestPool* pool() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool == nullptr) {
        int n = (int) opt(OPT_INGEST_THREADS);
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = hw >= 16 ? 8 : (hw >= 4 ? (int) hw / 2 : 2);   // measured: 4 - 8 workers saturate one PCIe Gen5 link
        }
        if (n > 32) n = 32;
        g_pool = new IngestPool(n);
    }
    return g_pool;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace
}  // namespace vk

using namespace vk;

extern "C" {

int vk_ingest_threads(void) { return pool()->size(); }

int vk_memcpy_h2d_staged(void* dst, const void* host_src, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    VK_REQUIRE(dst && host_src, "vk_memcpy_h2d_staged: NULL pointer");
    int device = 0;
    VK_CUDA(cudaGetDevice(&device));
    IngestPool* pl = pool();
    Job job;
    const size_t piece = piece_bytes();
    const size_t n_pieces = (size_t) ((bytes + piece - 1) / piece);
    job.remaining = (int) n_pieces;
    for (size_t i = 0; i < n_pieces; ++i) {
        const size_t off = i * piece;
        const size_t len = bytes - off < piece ? (size_t) (bytes - off) : piece;
        pl->submit(Piece{static_cast<uint8_t*>(dst) + off, static_cast<const uint8_t*>(host_src) + off, len,
                         (cudaStream_t) stream, device, &job});
    }
    std::unique_lock<std::mutex> lk(job.mu);
    job.cv.wait(lk, [&job] { return job.remaining == 0; });
    if (job.error != cudaSuccess) return cuda_fail(job.error, "vk_memcpy_h2d_staged");
    return VK_OK;
}

int vk_memcpy_h2d_auto(void* dst, const void* host_src, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    // pinned / registered memory is DMA-able as it is; small pageable copies are not worth the hand-off
    if (bytes >= (1u << 20) && opt(OPT_INGEST_STAGED) != 0 && is_pageable(host_src))
        return vk_memcpy_h2d_staged(dst, host_src, bytes, stream);
    VK_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
    return VK_OK;
}

}  // extern "C"
