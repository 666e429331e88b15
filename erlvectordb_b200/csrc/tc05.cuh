// tc05.cuh -- what the two tcgen05 candidate kernels share (gemm_tcgen05.cu: fp16 operands of float
// stores; gemm_i8.cu: u8 codes x query digit planes of quantization_8bit stores): PTX wrappers for
// mbarrier / TMA / tcgen05, the shared-memory matrix descriptors, and the accumulator-domain top-k
// epilogue (admission threshold, chunk filter, buffer pruning).  Everything is static / inline: each
// translation unit gets its own copy.
#pragma once
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"
#include "topk.cuh"

namespace evdb {

constexpr int GM = 128;        // queries per CTA tile (UMMA M)
constexpr int kEpiWarps = 16;       // tcgen05.ld is latency-bound per warp (tools/micro/ldtm_bench.cu): 4 warps per lane quarter
constexpr int kEpiParts = kEpiWarps / 4;   // column parts of a tile, one candidate list each
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;   // 4 control warps + the epilogue warps
constexpr int kGemmMaxKP = 128;
constexpr int kCandCapMax = 256;    // candidate buffer entries per (query, CTA, column part): 128 for KP <= 32, else 256
constexpr int kMaxSweeps = 8;       // query-block sweeps per launch (bounds the candidate buffers)

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants -----------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default .release.cta semantics, as for the local arrive: the TMEM reads this orders are fenced
    // by tcgen05.fence::before_thread_sync, and a cluster-scope release costs a MEMBAR + ERRBAR per tile per warp
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory, completion signalled on a barrier that may live in the peer
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address  [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                       // version = 1
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// K = 16 fp16 operand tile (32-byte rows), 32-byte swizzle: 8-row groups 256 bytes apart
__device__ __forceinline__ uint64_t make_sw32_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
    return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both K-major, M=128, N=256
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4)                 // c_format = F32
         | (0u << 7) | (0u << 10)    // a_format = b_format = F16
         | (0u << 15) | (0u << 16)   // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Largest accumulator value a (to a few ulps) with fma(a, c1, c0) >= tau, c1 < 0: a row whose
// accumulator is <= a has key score >= tau and can be skipped.  Never errs towards skipping more.
__device__ __forceinline__ float acc_threshold(float tau, float c0, float c1) {
    const float kInf = __int_as_float(0x7f800000);
    if (!(tau < kInf)) return -kInf;  // no threshold: admit everything
    uint32_t o = f32_orderable((tau - c0) / c1);
    uint32_t step = 1;
#pragma unroll 1
    for (int i = 0; i < 40 && fmaf(f32_from_orderable(o), c1, c0) < tau; ++i) {
        o = o > step ? o - step : 1u;
        step <<= 1;
        if (o <= 0x00800000u) return -kInf;  // ran off the float range: give up, admit everything
    }
    if (fmaf(f32_from_orderable(o), c1, c0) < tau) return -kInf;
#pragma unroll 1
    for (int i = 0; i < 4; ++i)
        if (fmaf(f32_from_orderable(o + 1), c1, c0) >= tau) ++o;
    return f32_from_orderable(o);
}

// Out-of-line (rare): reduce every buffer of this warp flagged in `need` to its KP best keys and
// tighten the owning lane's threshold.  wbase = the warp's first buffer column (entry i of lane e
// at wbase[i*GM + e]).  Software-pipelined: the next buffer's loads are in flight while the current
// one is selected and compacted (prunes come in bursts).  Returns the number of buffers pruned.
static __device__ __noinline__ int prune_buffers(uint64_t *wbase, unsigned need, const int KP, const int lane,
                                          const float c0, const float c1, int &cnt, float &tau, float &thrS) {
    auto load_buf = [&](int src, uint64_t (&x)[8]) {
        const int n = __shfl_sync(0xffffffffu, cnt, src);
        const uint64_t *b = wbase + src;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = r * 32 + lane;
            x[r] = e < n ? __ldcg(b + (size_t)e * GM) : kKeyMax;
        }
    };
    __syncwarp();
    int done = 0;
    int src = __ffs(need) - 1;
    need &= need - 1;
    uint64_t x[8];
    load_buf(src, x);  // > KP keys here
    while (true) {
        ++done;
        int nsrc = -1;
        uint64_t y[8];
        if (need) {
            nsrc = __ffs(need) - 1;
            need &= need - 1;
            load_buf(nsrc, y);
        }
        const uint32_t tau_o = warp_compact(x, KP, wbase + src, GM, lane);
        if (lane == src) {
            cnt = KP;
            tau = fminf(tau, f32_from_orderable(tau_o));
            thrS = fmaxf(thrS, acc_threshold(tau, c0, c1));
        }
        if (nsrc < 0) break;
#pragma unroll
        for (int r = 0; r < 8; ++r) x[r] = y[r];
        src = nsrc;
    }
    __syncwarp();
    return done;
}

// One 32-column chunk of one query's accumulator row: pooled maximum, and -- only when it beats
// the admission threshold -- the expansion that appends the hits.  Returns the chunk maximum.
// Columns past the end of the store hold acc = 0 (TMA zero fill); they can only cause a spurious
// expansion, the row bound is checked where a key is appended.  Kept small on purpose: the
// expansion is divergent code that every epilogue warp enters at a different time, and it has
// to stay resident in the instruction cache.
__device__ __forceinline__ float epi_chunk(const uint32_t (&v)[32], const float thrS, const uint32_t rowbase,
                                           const uint32_t nrows, const float c0, const float c1,
                                           uint64_t *__restrict__ mybuf, int &cnt, const bool admit) {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fmaxf(fmaxf(x[4 * i], x[4 * i + 1]), fmaxf(x[4 * i + 2], x[4 * i + 3]));
    const float m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
    if (admit && m > thrS) {
        // which 4-column groups hold a hit; each is fetched with selects (no dynamic register
        // index, no per-group branch) and expanded by one shared copy of the append code
        unsigned gm = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) gm |= (g[i] > thrS) ? (1u << i) : 0u;
#pragma unroll 1
        while (gm) {
            const int gi = __ffs(gm) - 1;
            gm &= gm - 1;
            float y0 = x[0], y1 = x[1], y2 = x[2], y3 = x[3];
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                const bool sel = gi == i;
                y0 = sel ? x[4 * i] : y0;
                y1 = sel ? x[4 * i + 1] : y1;
                y2 = sel ? x[4 * i + 2] : y2;
                y3 = sel ? x[4 * i + 3] : y3;
            }
            const uint32_t r0 = rowbase + 4 * gi;
            if (y0 > thrS && r0 < nrows) { mybuf[(size_t)cnt * GM] = make_key(fmaf(y0, c1, c0), r0); ++cnt; }
            if (y1 > thrS && r0 + 1 < nrows) { mybuf[(size_t)cnt * GM] = make_key(fmaf(y1, c1, c0), r0 + 1); ++cnt; }
            if (y2 > thrS && r0 + 2 < nrows) { mybuf[(size_t)cnt * GM] = make_key(fmaf(y2, c1, c0), r0 + 2); ++cnt; }
            if (y3 > thrS && r0 + 3 < nrows) { mybuf[(size_t)cnt * GM] = make_key(fmaf(y3, c1, c0), r0 + 3); ++cnt; }
        }
    }
    return m;
}


// ---- host side: the driver's tensor-map encoder through the runtime (no -lcuda) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

}  // namespace evdb
