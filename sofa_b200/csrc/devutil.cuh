// Block-level building blocks: warp-shuffle scans, staged (shared-memory) output
// tiles flushed with 128-bit coalesced stores, error slot.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sofa {

constexpr int PARSE_TPB = 256;
constexpr int EMIT_TPB = 128;
constexpr uint32_t OVERLAP_BYTES = 8192;      // a line may extend this far past its tile
constexpr uint32_t EMIT_STAGE_BYTES = 40960;  // shared-memory staging of one output tile

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (uint32_t)d) v += n;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan64(uint64_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (uint32_t)d) v += n;
    }
    return v;
}

// Exclusive scan of one value per thread over the whole block (blockDim.x multiple of 32,
// <= 1024).  sm: 33 words of shared memory.  All threads must call.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sm) {
    uint32_t lane = lane_id(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? sm[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    uint32_t res = sm[wid] + incl - v;
    *total = sm[32];
    __syncthreads();
    return res;
}
__device__ __forceinline__ uint64_t block_excl_scan64(uint64_t v, uint64_t* total, uint64_t* sm) {
    uint32_t lane = lane_id(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t incl = warp_incl_scan64(v);
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < nw ? sm[lane] : 0;
        uint64_t wi = warp_incl_scan64(w);
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    uint64_t res = sm[wid] + incl - v;
    *total = sm[32];
    __syncthreads();
    return res;
}

// ---- bulk-asynchronous copies (TMA engine, sm_90+): global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned addresses, size a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


// first failing record wins: (row << 8) | code, smaller row first
__device__ __forceinline__ void report_error(unsigned long long* err, uint64_t row, int code) {
    atomicMin(err, (unsigned long long)((row << 8) | (uint64_t)(code & 0xff)));
}

// Copy stage[a .. a+total) to gout[tile_off .. tile_off+total) where a == tile_off & 15 and
// gout_base is 16-byte aligned: interior with 128-bit stores, edges bytewise.
__device__ __forceinline__ void flush_staged(const unsigned char* stage, char* gout_base, uint64_t tile_off,
                                             uint32_t total) {
    uint32_t a = (uint32_t)(tile_off & 15);
    char* g = gout_base + (tile_off - a);                 // 16-aligned, g[i] <-> stage[i]
    uint32_t lo = a, hi = a + total;
    uint32_t v0 = (lo + 15) & ~15u, v1 = hi & ~15u;
    if (v1 < v0) { v0 = hi; v1 = hi; }                    // no aligned interior
    for (uint32_t i = lo + threadIdx.x; i < v0 && i < hi; i += blockDim.x) g[i] = (char)stage[i];
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);
    uint4* g4 = reinterpret_cast<uint4*>(g);
    for (uint32_t i = (v0 >> 4) + threadIdx.x; i < (v1 >> 4); i += blockDim.x) g4[i] = s4[i];
    uint32_t t0 = v1 > v0 ? v1 : v0;
    for (uint32_t i = t0 + threadIdx.x; i < hi; i += blockDim.x) g[i] = (char)stage[i];
}

// Same, with the 16-byte aligned interior written by ONE bulk-asynchronous copy (shared -> global, TMA engine) issued by
// thread 0 instead of 128-bit stores from every thread: the block's warps spend no issue slots on the flush.  The
// generic-proxy writes that filled the image are made visible to the async proxy first (fence.proxy.async), and thread 0
// waits until the engine has READ the image before the block may retire.  All threads must call; ends with a barrier.
__device__ __forceinline__ void flush_staged_bulk(const unsigned char* stage, char* gout_base, uint64_t tile_off, uint32_t total) {
    uint32_t a = (uint32_t)(tile_off & 15);
    char* g = gout_base + (tile_off - a);                 // 16-aligned, g[i] <-> stage[i]
    uint32_t lo = a, hi = a + total;
    uint32_t v0 = (lo + 15) & ~15u, v1 = hi & ~15u;
    if (v1 < v0) { v0 = hi; v1 = hi; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && v1 > v0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(g + v0), "r"(smem_u32(stage + v0)), "r"(v1 - v0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    for (uint32_t i = lo + threadIdx.x; i < v0 && i < hi; i += blockDim.x) g[i] = (char)stage[i];
    uint32_t t0 = v1 > v0 ? v1 : v0;
    for (uint32_t i = t0 + threadIdx.x; i < hi; i += blockDim.x) g[i] = (char)stage[i];
    if (threadIdx.x == 0 && v1 > v0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace sofa
