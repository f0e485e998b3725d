// tsdr_ring.cu -- the buffer ring between the radio (or file) producer thread and the processing
// thread, in page-locked host memory.
//
// It replaces AtomicCircularBuffer{T} (src/AtomicAbstractSDRs.jl:67-190) with the same observable
// behaviour -- `depth` slots of one recv! buffer each; the producer never waits for the consumer and
// overwrites the slot at its write pointer; the consumer waits until at least one buffer is marked
// new, reads the slot at ITS pointer and advances -- but
//   * the slots are cudaHostAlloc'ed, so the chain's H2D copy of a slot is a true asynchronous DMA at
//     PCIe speed (pageable Julia arrays are staged through a driver bounce buffer);
//   * the consumer can borrow the slot (acquire/release) instead of copying it out (circ_take! copies);
//   * waiting is a condition variable, not a yield() spin (wait_consData, :142-150).
// Host code only; no kernels.  pinned = 0 allocates ordinary memory (hosts without a GPU, CPU tests).
#include "tsdr_internal.cuh"

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <vector>

struct tsdr_ring {
    size_t slot_bytes;
    int depth;
    int pinned;
    unsigned char* mem;
    std::vector<std::mutex>* slot_lock;   // AtomicBuffer.lock[i]   (:46-56)
    std::mutex m;                         // guards the three counters below (ptr_write/ptr_read/t_new locks)
    std::condition_variable cv;
    int ptr_write, ptr_read, t_new;
    int borrowed_read, borrowed_write;    // slot currently lent to the consumer / producer, or -1
    uint64_t produced, consumed, overwritten;
};

using namespace tsdr;

extern "C" {

int tsdr_ring_create(tsdr_ring** out, size_t slot_bytes, int depth, int pinned) {
    TSDR_REQUIRE(out, "out is NULL");
    TSDR_REQUIRE(slot_bytes > 0 && depth >= 1, "ring needs slot_bytes > 0 and depth >= 1");
    tsdr_ring* r = new (std::nothrow) tsdr_ring();
    if (!r) return TSDR_ERR_NOMEM;
    r->slot_bytes = slot_bytes; r->depth = depth; r->pinned = pinned ? 1 : 0;
    r->ptr_write = r->ptr_read = r->t_new = 0;
    r->borrowed_read = r->borrowed_write = -1;
    r->produced = r->consumed = r->overwritten = 0;
    r->slot_lock = new (std::nothrow) std::vector<std::mutex>(depth);
    r->mem = nullptr;
    const size_t total = slot_bytes * (size_t)depth;
    if (r->pinned) {
        void* p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, total, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaHostAlloc(%zu) for the ring failed: %s", total, cudaGetErrorString(e));
            delete r->slot_lock; delete r;
            return e == cudaErrorMemoryAllocation ? TSDR_ERR_NOMEM : TSDR_ERR_CUDA;
        }
        r->mem = (unsigned char*)p;
    } else {
        r->mem = (unsigned char*)aligned_alloc(64, (total + 63) & ~(size_t)63);
        if (!r->mem) { delete r->slot_lock; delete r; set_error("ring allocation of %zu bytes failed", total); return TSDR_ERR_NOMEM; }
    }
    memset(r->mem, 0, total);   // zeros(T, nEch*depth), :51
    *out = r;
    return TSDR_OK;
}

int tsdr_ring_destroy(tsdr_ring* r) {
    if (!r) return TSDR_OK;
    if (r->pinned) cudaFreeHost(r->mem); else free(r->mem);
    delete r->slot_lock;
    delete r;
    return TSDR_OK;
}

// producer side of circ_put! (:159-170), zero copy: lend the slot at the write pointer ...
int tsdr_ring_acquire_write(tsdr_ring* r, void** slot) {
    TSDR_REQUIRE(r && slot, "NULL argument");
    TSDR_REQUIRE(r->borrowed_write < 0, "a write slot is already acquired");
    int pos;
    { std::lock_guard<std::mutex> g(r->m); pos = r->ptr_write; }
    (*r->slot_lock)[pos].lock();           // waits while the consumer still holds this very slot (atomic_write, :112-117)
    r->borrowed_write = pos;
    *slot = r->mem + (size_t)pos * r->slot_bytes;
    return TSDR_OK;
}

// ... and publish it: advance the write pointer, mark one more buffer new (saturating at depth, :122-126)
int tsdr_ring_commit(tsdr_ring* r) {
    TSDR_REQUIRE(r, "NULL argument");
    TSDR_REQUIRE(r->borrowed_write >= 0, "no write slot acquired");
    (*r->slot_lock)[r->borrowed_write].unlock();
    r->borrowed_write = -1;
    {
        std::lock_guard<std::mutex> g(r->m);
        r->ptr_write = (r->ptr_write + 1) % r->depth;
        if (r->t_new == r->depth) r->overwritten += 1;   // the consumer lags a whole ring: one unread buffer was lost
        else r->t_new += 1;
        r->produced += 1;
    }
    r->cv.notify_one();
    return TSDR_OK;
}

int tsdr_ring_put(tsdr_ring* r, const void* data, size_t bytes) {
    TSDR_REQUIRE(r && data, "NULL argument");
    // the reference asserts equal lengths (:113)
    TSDR_REQUIRE(bytes == r->slot_bytes, "buffer of %zu bytes does not match the ring's slot size %zu", bytes, r->slot_bytes);
    void* slot = nullptr;
    int rc = tsdr_ring_acquire_write(r, &slot);
    if (rc) return rc;
    memcpy(slot, data, bytes);
    return tsdr_ring_commit(r);
}

// consumer side of circ_take! (:176-189): wait for new data, lend the slot at the read pointer.
// timeout_ms < 0 waits forever; on timeout the status is TSDR_ERR_BOUNDS ("nothing there") and no slot is held.
int tsdr_ring_acquire_read(tsdr_ring* r, const void** slot, int timeout_ms) {
    TSDR_REQUIRE(r && slot, "NULL argument");
    TSDR_REQUIRE(r->borrowed_read < 0, "a read slot is already acquired");
    int pos;
    {
        std::unique_lock<std::mutex> g(r->m);
        if (timeout_ms < 0) r->cv.wait(g, [&] { return r->t_new > 0; });
        else if (!r->cv.wait_for(g, std::chrono::milliseconds(timeout_ms), [&] { return r->t_new > 0; })) {
            set_error("no new buffer in the ring within %d ms", timeout_ms);
            return TSDR_ERR_BOUNDS;
        }
        pos = r->ptr_read;
    }
    (*r->slot_lock)[pos].lock();
    r->borrowed_read = pos;
    *slot = r->mem + (size_t)pos * r->slot_bytes;
    return TSDR_OK;
}

int tsdr_ring_release_read(tsdr_ring* r) {
    TSDR_REQUIRE(r, "NULL argument");
    TSDR_REQUIRE(r->borrowed_read >= 0, "no read slot acquired");
    (*r->slot_lock)[r->borrowed_read].unlock();
    r->borrowed_read = -1;
    std::lock_guard<std::mutex> g(r->m);
    r->ptr_read = (r->ptr_read + 1) % r->depth;       // atomic_update, :103-107
    r->t_new = r->t_new > 0 ? r->t_new - 1 : 0;       // atomic_consData, :130-134
    r->consumed += 1;
    return TSDR_OK;
}

int tsdr_ring_take(tsdr_ring* r, void* out, size_t bytes, int timeout_ms) {
    TSDR_REQUIRE(r && out, "NULL argument");
    TSDR_REQUIRE(bytes == r->slot_bytes, "buffer of %zu bytes does not match the ring's slot size %zu", bytes, r->slot_bytes);
    const void* slot = nullptr;
    int rc = tsdr_ring_acquire_read(r, &slot, timeout_ms);
    if (rc) return rc;
    memcpy(out, slot, bytes);
    return tsdr_ring_release_read(r);
}

int tsdr_ring_stats(tsdr_ring* r, int* available, uint64_t* produced, uint64_t* consumed, uint64_t* overwritten) {
    TSDR_REQUIRE(r, "NULL argument");
    std::lock_guard<std::mutex> g(r->m);
    if (available) *available = r->t_new;
    if (produced) *produced = r->produced;
    if (consumed) *consumed = r->consumed;
    if (overwritten) *overwritten = r->overwritten;
    return TSDR_OK;
}

size_t tsdr_ring_slot_bytes(const tsdr_ring* r) { return r ? r->slot_bytes : 0; }

}  // extern "C"
