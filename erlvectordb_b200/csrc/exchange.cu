// exchange.cu -- the cross-GPU step of a row-sharded search as peer-memory stores.
//
// The reference has no sharded search (cluster_manager replicates whole stores, reference
// src/cluster_manager.erl:148-171); the exchange after each GPU's local top-k is this design's own
// step.  It is tiny (B*k*16 + B*8 bytes per rank) and latency-bound, so instead of
// an NCCL allgather followed by a merge launch, every rank
//   1. PUSHES its packed result blob straight into a mailbox slot in every peer's memory
//      (plain st.global over NVLink, peers mapped with CUDA IPC), then publishes an epoch flag
//      there (system-scope fence in between);
//   2. runs the merge kernel, whose CTAs first wait for all ranks' flags to reach the epoch.
// The waiting kernel depends only on kernels of OTHER GPUs (each rank's push precedes its merge in
// stream order), the wait is bounded (trap), and the mailboxes are double-buffered by epoch
// parity: a rank can be at most one search ahead of a peer, because its next push needs its own
// merge done, which needed that peer's previous push.
#include <string.h>

#include <new>

#include "internal.h"

namespace evdb {

static inline size_t box_words(const evdb_exchange *x) { return 2 * (size_t)x->world * x->slot_words; }
static inline size_t box_bytes(const evdb_exchange *x) { return (box_words(x) + 2 * (size_t)x->world + 8) * sizeof(uint64_t); }

__global__ void __launch_bounds__(256) exchange_push_kernel(const uint64_t *__restrict__ blob, size_t words,
                                                            uint64_t *const *__restrict__ peer_box, int rank,
                                                            int world, size_t slot_words, int parity,
                                                            unsigned long long epoch, unsigned int *done_counter) {
    // every peer's copy of my slot: [parity][rank]
    const size_t slot_off = ((size_t)parity * world + rank) * slot_words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t v = blob[i];
        for (int p = 0; p < world; ++p) peer_box[p][slot_off + i] = v;
    }
    __threadfence_system();   // my stores are visible system-wide before the flag below can be
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            // flags live behind the slots: [2][world]
            unsigned long long *f = reinterpret_cast<unsigned long long *>(peer_box[threadIdx.x] + 2 * (size_t)world * slot_words) +
                                    (size_t)parity * world + rank;
            *reinterpret_cast<volatile unsigned long long *>(f) = epoch;
        }
        if (threadIdx.x == 0) *done_counter = 0;
    }
}

int exchange_push_words(evdb_exchange *x, const void *d_blob, size_t words, cudaStream_t st) {
    if (!x || !d_blob || words == 0 || words > x->max_words) return EVDB_E_BAD_ARG;
    EVDB_CUDA(cudaSetDevice(x->device));
    x->epoch++;
    const int parity = (int)(x->epoch & 1);
    unsigned int *counter = reinterpret_cast<unsigned int *>(x->mailbox + box_words(x) + 2 * (size_t)x->world);
    int grid = (int)((words + 255) / 256);
    if (grid > 64) grid = 64;
    exchange_push_kernel<<<grid, 256, 0, st>>>((const uint64_t *)d_blob, words, x->d_peer_box, x->rank, x->world,
                                               x->slot_words, parity, x->epoch, counter);
    EVDB_CUDA(cudaGetLastError());
    return EVDB_OK;
}

PushTarget exchange_begin_push(evdb_exchange *x) {
    x->epoch++;
    const int parity = (int)(x->epoch & 1);
    PushTarget t;
    t.peer_box = x->d_peer_box;
    t.slot_off = ((unsigned long long)parity * x->world + x->rank) * x->slot_words;
    t.flag_off = 2ull * x->world * x->slot_words + (unsigned long long)parity * x->world + x->rank;
    t.epoch = x->epoch;
    t.counter = reinterpret_cast<unsigned int *>(x->mailbox + box_words(x) + 2 * (size_t)x->world) + 1;
    t.world = x->world;
    return t;
}

ExchangeView exchange_view(const evdb_exchange *x) {
    const int parity = (int)(x->epoch & 1);
    ExchangeView v;
    v.slots = x->mailbox + (size_t)parity * x->world * x->slot_words;
    v.stride = x->slot_words;
    v.flags = reinterpret_cast<const unsigned long long *>(x->mailbox + box_words(x)) + (size_t)parity * x->world;
    v.epoch = x->epoch;
    return v;
}

}  // namespace evdb

using namespace evdb;

extern "C" {

int evdb_exchange_create(int device, int rank, int world, uint64_t max_words, evdb_exchange **out,
                         void *ipc_handle_out) {
    if (!out || world < 1 || world > 64 || rank < 0 || rank >= world || max_words == 0) return EVDB_E_BAD_ARG;
    *out = nullptr;
    EVDB_TRY(check_device_public(device));
    EVDB_CUDA(cudaSetDevice(device));
    evdb_exchange *x = new (std::nothrow) evdb_exchange();
    if (!x) return EVDB_E_OOM;
    x->device = device; x->rank = rank; x->world = world;
    x->max_words = max_words;
    x->slot_words = (max_words + 31) / 32 * 32;
    cudaError_t e = cudaMalloc((void **)&x->mailbox, box_bytes(x));
    if (e == cudaSuccess) e = cudaMemset(x->mailbox, 0, box_bytes(x));
    if (e == cudaSuccess) e = cudaMalloc((void **)&x->d_peer_box, sizeof(uint64_t *) * world);
    if (e == cudaSuccess && ipc_handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, x->mailbox);
        if (e == cudaSuccess) memcpy(ipc_handle_out, &h, sizeof(h));
    }
    if (e != cudaSuccess) {
        set_cuda_error(e, __FILE__, __LINE__);
        cudaFree(x->mailbox); cudaFree(x->d_peer_box);
        delete x;
        return e == cudaErrorMemoryAllocation ? EVDB_E_OOM : EVDB_E_CUDA;
    }
    *out = x;
    return EVDB_OK;
}

/* all_handles: world cudaIpcMemHandle_t (64 bytes each) in rank order, as gathered from every
 * rank's evdb_exchange_create.  Other PROCESSES' mailboxes are opened; rank's own is used directly. */
int evdb_exchange_connect(evdb_exchange *x, const void *all_handles) {
    if (!x || !all_handles) return EVDB_E_BAD_ARG;
    EVDB_CUDA(cudaSetDevice(x->device));
    uint64_t *host[64];
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank) { host[p] = x->mailbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)all_handles + (size_t)p * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        EVDB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_base[p] = ptr;
        x->opened[p] = true;
        host[p] = (uint64_t *)ptr;
    }
    EVDB_CUDA(cudaMemcpy(x->d_peer_box, host, sizeof(uint64_t *) * x->world, cudaMemcpyHostToDevice));
    return EVDB_OK;
}

/* Same-process variant (tests: several "ranks" on one device): the mailbox base pointers directly. */
int evdb_exchange_connect_ptrs(evdb_exchange *x, const void *const *mailboxes) {
    if (!x || !mailboxes) return EVDB_E_BAD_ARG;
    EVDB_CUDA(cudaSetDevice(x->device));
    EVDB_CUDA(cudaMemcpy(x->d_peer_box, mailboxes, sizeof(uint64_t *) * x->world, cudaMemcpyHostToDevice));
    return EVDB_OK;
}

void *evdb_exchange_mailbox(evdb_exchange *x) { return x ? (void *)x->mailbox : nullptr; }

/* Enqueue the push of this rank's blob for the next search (no sync). */
int evdb_exchange_push(evdb_exchange *x, const void *d_local_blob, int B, int k, void *stream) {
    if (B <= 0 || k <= 0) return EVDB_E_BAD_ARG;
    return exchange_push_words(x, d_local_blob, 2 * (size_t)B * k + (size_t)B, (cudaStream_t)stream);
}

/* Enqueue the merge of the current search: waits (on the device) for every rank's push of this
 * epoch, then merges the world blobs of the local mailbox into d_out_blob (packed layout). */
int evdb_exchange_merge(evdb_exchange *x, int B, int k, void *d_out_blob, void *stream) {
    if (!x || !d_out_blob || B <= 0 || k <= 0) return EVDB_E_BAD_ARG;
    EVDB_CUDA(cudaSetDevice(x->device));
    const ExchangeView v = exchange_view(x);
    return launch_merge_topk_packed(v.slots, v.stride, x->world, B, k, (uint64_t *)d_out_blob, (cudaStream_t)stream,
                                    v.flags, v.epoch);
}

void evdb_exchange_destroy(evdb_exchange *x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < x->world; ++p)
        if (x->opened[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    cudaFree(x->mailbox);
    cudaFree(x->d_peer_box);
    cudaGetLastError();
    delete x;
}

}  // extern "C"
