// Filter series of report.js gathered from the resident main series (included by tu_emit.cu only).
#pragma once
#include "kernels_emit.cuh"

namespace sofa {

// ---------------------------------------------------------------------------- filter series by gather
// A filter series of report.js holds the records of the main series whose name matches the keyword: the same
// bytes.  Block b takes EMIT_TPB selected candidates (ascending), sizes them from the main series' record
// offsets, finds its byte offset with the same decoupled look-back as k_emit_fused, copies the records into a
// shared-memory image (one warp per record, lanes stride the bytes: coalesced reads of a record that is already
// in L2 or HBM) and flushes the image with 128-bit stores.
struct GatherArgs {
    const char* main_out; const unsigned long long* rec_off;
    const uint32_t* index; uint64_t nsel;
    char* out; uint64_t cap;
    unsigned long long* state; unsigned int* ticket;
    unsigned long long* total; unsigned int* overflow;
    uint32_t slot_bytes;
};
static __global__ void __launch_bounds__(EMIT_TPB) k_gather_records(GatherArgs a) {
    extern __shared__ __align__(16) unsigned char stage[];
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_src[EMIT_TPB];
    __shared__ uint32_t s_ex[EMIT_TPB + 1];
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t b = s_tile;
    const uint64_t ci = (uint64_t)b * EMIT_TPB + threadIdx.x;
    uint32_t len = 0;
    unsigned long long so = 0;
    if (ci < a.nsel) {
        const uint32_t c = a.index[ci];
        so = a.rec_off[c];
        len = (uint32_t)(a.rec_off[c + 1] - so);
    }
    uint32_t total;
    const uint32_t ex = block_excl_scan(len, &total, sm);
    s_src[threadIdx.x] = so; s_ex[threadIdx.x] = ex;
    if (threadIdx.x == EMIT_TPB - 1) s_ex[EMIT_TPB] = ex + len;
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        unsigned long long base = 0;
        if (b > 0) {
            if (lane == 0) st_release_u64(&a.state[b], (1ull << 62) | (unsigned long long)total);
            long long j = (long long)b - 1;
            while (true) {
                long long idx = j - (long long)lane;
                unsigned long long v = idx >= 0 ? ld_acquire_u64(&a.state[idx]) : (2ull << 62);
                unsigned int f = (unsigned int)(v >> 62);
                unsigned int pmask = __ballot_sync(0xffffffffu, f == 2);
                unsigned int emask = __ballot_sync(0xffffffffu, f == 0);
                int first = pmask ? (__ffs(pmask) - 1) : 32;
                unsigned int need = first >= 31 ? 0xffffffffu : ((1u << (first + 1)) - 1u);
                if (emask & need) continue;
                unsigned long long contrib = ((int)lane <= first) ? (v & LB_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                base += contrib;
                if (first < 32) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st_release_u64(&a.state[b], (2ull << 62) | (base + total));
            s_base = base;
            if ((uint64_t)b + 1 == (a.nsel + EMIT_TPB - 1) / EMIT_TPB) *a.total = base + total;
        }
    }
    __syncthreads();
    const unsigned long long base = s_base;
    if (base + total > a.cap) { if (threadIdx.x == 0) atomicOr(a.overflow, 1u); return; }
    const bool staged = total + 16 <= a.slot_bytes;
    const uint32_t ph = (uint32_t)(base & 15);
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // every warp copies 32 consecutive records of the block, four at a time: the aligned 32-bit source words of the four
    // records (two per lane and record: 256 source bytes) are all requested before the first byte is stored, so that
    // one round trip to L2 / HBM serves four records (one record at a time ran at ~0.5 TB/s: latency bound); longer
    // records finish in a tail loop.  The bytes land one by one (source and destination are not mutually aligned).
    constexpr int RPW = EMIT_TPB / (EMIT_TPB / 32);      // records per warp = 32
    const uint32_t rbase = wid * RPW;
    for (uint32_t r0 = 0; r0 < RPW; r0 += 4) {
        uint32_t v[4][2], e0[4], l[4], sh[4], nwords[4];
        const uint32_t* sw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = rbase + r0 + u;
            e0[u] = s_ex[r]; l[u] = s_ex[r + 1] - e0[u];
            const unsigned long long sa = (unsigned long long)(a.main_out + s_src[r]);
            sh[u] = (uint32_t)(sa & 3);
            sw[u] = reinterpret_cast<const uint32_t*>(sa - sh[u]);
            nwords[u] = l[u] ? (sh[u] + l[u] + 3) >> 2 : 0;
#pragma unroll
            for (int q = 0; q < 2; q++) { const uint32_t wi = q * 32 + lane; v[u][q] = wi < nwords[u] ? __ldg(sw[u] + wi) : 0u; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t wi = q * 32 + lane;
                if (wi < nwords[u]) {
#pragma unroll
                    for (int bb = 0; bb < 4; bb++) {
                        const int p = (int)(wi * 4 + bb) - (int)sh[u];
                        if (p >= 0 && p < (int)l[u]) {
                            const unsigned char ch = (unsigned char)(v[u][q] >> (8 * bb));
                            if (staged) stage[ph + e0[u] + p] = ch; else a.out[base + e0[u] + p] = (char)ch;
                        }
                    }
                }
            }
            for (uint32_t wi = 64 + lane; wi < nwords[u]; wi += 32) {          // records longer than ~250 bytes
                const uint32_t w = __ldg(sw[u] + wi);
#pragma unroll
                for (int bb = 0; bb < 4; bb++) {
                    const int p = (int)(wi * 4 + bb) - (int)sh[u];
                    if (p >= 0 && p < (int)l[u]) {
                        const unsigned char ch = (unsigned char)(w >> (8 * bb));
                        if (staged) stage[ph + e0[u] + p] = ch; else a.out[base + e0[u] + p] = (char)ch;
                    }
                }
            }
        }
    }
    if (staged && total) flush_staged_bulk(stage, a.out, base, total);       // one bulk-asynchronous store per block
}

}  // namespace sofa
