// scan_common.cuh -- pieces of the float scans shared by scan.cu and the fused small-store kernel of
// select.cu: per-chunk accumulation (reference src/vector_store.erl:238-252, src/vector_utils.erl:38-43
// in fp32 as a candidate score), per-warp candidate buffers, the CTA merge.
#pragma once
#include "internal.h"
#include "topk.cuh"

namespace evdb {

// ----------------------------------------------------------------------------
// per-chunk accumulation
// ----------------------------------------------------------------------------
template <int METRIC>
__device__ __forceinline__ void acc_f4(float4 &a, const float4 v, const float4 q) {
    if (METRIC == EVDB_COSINE) {
        a.x = fmaf(v.x, q.x, a.x); a.y = fmaf(v.y, q.y, a.y);
        a.z = fmaf(v.z, q.z, a.z); a.w = fmaf(v.w, q.w, a.w);
    } else if (METRIC == EVDB_EUCLIDEAN) {
        float t0 = v.x - q.x, t1 = v.y - q.y, t2 = v.z - q.z, t3 = v.w - q.w;
        a.x = fmaf(t0, t0, a.x); a.y = fmaf(t1, t1, a.y);
        a.z = fmaf(t2, t2, a.z); a.w = fmaf(t3, t3, a.w);
    } else {
        a.x += fabsf(v.x - q.x); a.y += fabsf(v.y - q.y);
        a.z += fabsf(v.z - q.z); a.w += fabsf(v.w - q.w);
    }
}

__device__ __forceinline__ float4 bf16x4_lo(const uint4 w) {  // elements 0..3 of 8 bf16
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u),
                       __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xFFFF0000u));
}
__device__ __forceinline__ float4 bf16x4_hi(const uint4 w) {  // elements 4..7
    return make_float4(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xFFFF0000u),
                       __uint_as_float(w.w << 16), __uint_as_float(w.w & 0xFFFF0000u));
}

// ----------------------------------------------------------------------------
// shared epilogue: threshold test, warp-list insertion, CTA merge
// ----------------------------------------------------------------------------
__device__ __forceinline__ void offer(uint64_t key, uint64_t &thr, uint64_t *mylist, int KP,
                                      int lane) {
    unsigned m = __ballot_sync(0xffffffffu, key < thr);
    while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        uint64_t k2 = __shfl_sync(0xffffffffu, key, src);
        if (k2 < thr) thr = warp_list_insert(mylist, KP, k2, lane);
    }
}

__device__ __forceinline__ void cta_merge_and_store(uint64_t *lists, int KP, uint64_t *out) {
    __syncthreads();
    block_bitonic_sort(lists, kScanWarps * KP);
    for (int i = threadIdx.x; i < KP; i += blockDim.x) out[i] = lists[i];
}

// Per-warp candidate state: append-and-prune buffers (KP <= 128) or the sorted-list insert
// (wider escalation windows).  Region of warp w: lists + w * stride.
struct WarpCands {
    uint64_t *mine;
    uint64_t thr;
    int cnt, cap, KP;
    bool append;
    __device__ __forceinline__ void init(uint64_t *lists, int KP_, int warp) {
        KP = KP_;
        append = KP_ <= kAppendMaxKP;
        cap = append ? append_cap(KP_) : KP_;
        mine = lists + (size_t)warp * cap;
        thr = kKeyMax;
        cnt = 0;
    }
    __device__ __forceinline__ void offer(uint64_t key, int lane) {
        if (append) offer_append(key, thr, mine, cnt, cap, KP, lane);
        else evdb::offer(key, thr, mine, KP, lane);
    }
    // compact every warp's best <= KP keys to lists[w*KP ..], padded with kKeyMax (whole CTA calls)
    // `active` = false for a warp that only takes part in the barrier (the TMA producer warp)
    __device__ __forceinline__ void finish(uint64_t *lists, int warp, int lane, bool active = true) {
        if (!append) return;
        if (active && cnt > KP) warp_buf_prune(mine, cnt, thr, KP, lane);
        uint64_t e[kAppendMaxKP / 32];
#pragma unroll
        for (int r = 0; r < kAppendMaxKP / 32; ++r) {
            const int i = r * 32 + lane;
            e[r] = (active && i < cnt && i < KP) ? mine[i] : kKeyMax;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kAppendMaxKP / 32; ++r) {
            const int i = r * 32 + lane;
            if (active && i < KP) lists[(size_t)warp * KP + i] = e[r];
        }
    }
};

static inline size_t scan_list_bytes(int KP) {
    const int per = KP <= kAppendMaxKP ? append_cap(KP) : KP;
    return (size_t)kScanWarps * per * sizeof(uint64_t);
}

}  // namespace evdb
