// Persistent parse kernel with bulk-asynchronous (TMA) tile staging - the sm_100a variant of k_parse for the three
// trace sources.
//
// k_parse stages a tile with LDG -> STS loops, then a barrier, then parses: the copy and the parse of a block never
// overlap, and only other resident blocks hide the load latency.  Here a block keeps TWO tile buffers: while its
// threads tokenise tile i out of one buffer, the copy engine fills the other with tile i + gridDim.x through one
// cp.async.bulk (UBLKCP in SASS) that signals an mbarrier with its byte count; nobody spends issue slots on staging
// and the block never waits for HBM after its first tile.  Blocks are persistent (grid = resident blocks) and walk
// the tiles round-robin; tile row bases come from the count_lines + scan pass, so tiles are independent.
#pragma once
#include "kernels_ingest.cuh"

namespace sofa {

constexpr uint32_t TMA_STAGE_PAD = 64;
// Bytes staged past the end of a tile for the tile's last line.  k_parse stages the whole OVERLAP_BYTES (8 KB, the longest
// line supported); with two buffers per block that would cost a resident block per SM, so only 1 KB is staged here and
// the rare last line that runs past it is parsed straight out of global memory (generic pointers: the row codecs do
// not care which memory a line is in).
constexpr uint32_t TMA_OVERLAP_STAGED = 1024;

// bytes [s0, s0 + stage_len) of the text that tile t needs (same window as k_parse)
struct TileWin { uint64_t b0, b1, qlo, qhi, s0, s1, s1s; uint32_t stage_len; };
__device__ __forceinline__ TileWin tile_window(uint64_t t, uint64_t begin, uint64_t end, uint32_t tile_bytes) {
    TileWin w;
    w.b0 = begin + t * (uint64_t)tile_bytes;
    w.b1 = w.b0 + tile_bytes < end ? w.b0 + tile_bytes : end;
    w.qlo = (w.b0 == begin) ? begin : w.b0 - 1;
    w.qhi = (w.b1 - 1 < end - 1) ? w.b1 - 1 : end - 1;
    w.s0 = w.qlo & ~15ull;
    w.s1 = w.b1 + OVERLAP_BYTES < end ? w.b1 + OVERLAP_BYTES : end;                          // where a line may end at the latest
    w.s1s = w.b1 + TMA_OVERLAP_STAGED < end ? w.b1 + TMA_OVERLAP_STAGED : end;               // end of the staged bytes
    w.stage_len = (uint32_t)(((w.s1s + 15) & ~15ull) - w.s0);
    return w;
}

template <int KIND>
__global__ void __launch_bounds__(PARSE_TPB)
k_parse_tma(const char* __restrict__ text, uint64_t begin, uint64_t end, uint32_t tile_bytes, uint32_t max_lines, uint64_t ntiles,
            const uint64_t* __restrict__ tile_row_base, IngestCfg cfg, AllCols cols, unsigned long long* err) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t sm[33];
    __shared__ __align__(8) uint64_t bars[2];
    const uint32_t buf_bytes = (tile_bytes + TMA_OVERLAP_STAGED + TMA_STAGE_PAD + 127u) & ~127u;
    unsigned char* bufs[2] = {smem, smem + buf_bytes};
    uint16_t* lstart = reinterpret_cast<uint16_t*>(smem + 2 * (size_t)buf_bytes);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t t = blockIdx.x;
    if (t < ntiles && threadIdx.x == 0) {
        TileWin w = tile_window(t, begin, end, tile_bytes);
        mbar_expect_tx(&bars[0], w.stage_len);
        bulk_g2s(bufs[0], text + w.s0, w.stage_len, &bars[0]);
    }
    for (uint32_t it = 0; t < ntiles; it++, t += gridDim.x) {
        const uint32_t b = it & 1;
        const uint64_t tn = t + gridDim.x;
        if (tn < ntiles && threadIdx.x == 0) {             // buffer b^1 was last read before the barrier that ended iteration it-1
            TileWin wn = tile_window(tn, begin, end, tile_bytes);
            mbar_expect_tx(&bars[b ^ 1], wn.stage_len);
            bulk_g2s(bufs[b ^ 1], text + wn.s0, wn.stage_len, &bars[b ^ 1]);
        }
        const TileWin w = tile_window(t, begin, end, tile_bytes);
        mbar_wait(&bars[b], (it >> 1) & 1);
        unsigned char* stage = bufs[b];
        const uint64_t s0 = w.s0;
        // ---- newline scan (as k_parse): 8-byte words of the staged bytes, counts -> block scan -> line start table
        uint32_t scan_len = (uint32_t)(w.qhi > s0 ? w.qhi - s0 : 0);
        uint32_t chunk = ((scan_len + blockDim.x - 1) / blockDim.x + 7) & ~7u;
        if (((chunk >> 3) & 1u) == 0) chunk += 8;
        uint32_t c0 = threadIdx.x * chunk, c1 = c0 + chunk < scan_len ? c0 + chunk : scan_len;
        uint32_t lo_rel = (uint32_t)(w.qlo - s0);
        uint32_t cnt = 0;
        const uint64_t* st8 = reinterpret_cast<const uint64_t*>(stage);
        for (uint32_t i = c0; i < c1; i += 8) {
            uint32_t m = eq8(st8[i >> 3], '\n');
            if (i < lo_rel) m &= (lo_rel - i >= 8) ? 0u : ~((1u << (lo_rel - i)) - 1u);
            if (i + 8 > c1) m &= (1u << (c1 - i)) - 1u;
            cnt += __popc(m);
        }
        const bool first_tile = (w.b0 == begin && begin < end);
        if (threadIdx.x == 0 && first_tile) cnt += 1;
        uint32_t total;
        uint32_t ex = block_excl_scan(cnt, &total, sm);
        const bool too_many = total > max_lines;
        if (too_many) {
            if (threadIdx.x == 0) report_error(err, tile_row_base[t], E_TOO_MANY_LINES);
            total = 0;
        } else {
            uint32_t k = ex;
            if (threadIdx.x == 0 && first_tile) lstart[k++] = (uint16_t)(begin - s0);
            for (uint32_t i = c0; i < c1; i += 8) {
                uint32_t m = eq8(st8[i >> 3], '\n');
                if (i < lo_rel) m &= (lo_rel - i >= 8) ? 0u : ~((1u << (lo_rel - i)) - 1u);
                if (i + 8 > c1) m &= (1u << (c1 - i)) - 1u;
                while (m) { lstart[k++] = (uint16_t)(i + (uint32_t)ctz32(m) + 1); m &= m - 1; }
            }
        }
        __syncthreads();
        const uint32_t stage_end = (uint32_t)(w.s1s - s0);
        for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
            uint32_t ls = lstart[j], le;
            bool has_nl = true;
            int code = E_OK;
            const char* line = reinterpret_cast<const char*>(stage + ls);
            if (j + 1 < total) {
                le = (uint32_t)lstart[j + 1] - 1;
            } else {
                le = ls;
                while (le < stage_end && stage[le] != '\n') le++;
                if (le >= stage_end) {
                    if (w.s1s >= end) has_nl = false;
                    else {                                  // the line runs past the staged bytes: finish it in global memory
                        const char* g = text + s0;
                        const uint32_t lim = (uint32_t)(w.s1 - s0);
                        while (le < lim && g[le] != '\n') le++;
                        if (le >= lim) { if (w.s1 >= end) has_nl = false; else code = E_LINE_TOO_LONG; }
                        line = g + ls;
                    }
                }
            }
            const uint64_t row = tile_row_base[t] + j;
            const int n = (int)(le - ls);
            if (code == E_OK) {
                if (KIND == SRC_PERF) {
                    PerfRow r;
                    code = parse_perf(line, n, cfg, &r);
                    if (code == E_OK) {
                        cols.p.ts[row] = r.ts; cols.p.event[row] = r.event; cols.p.dur[row] = r.dur;
                        cols.p.pid[row] = r.pid; cols.p.tid[row] = r.tid;
                        cols.p.line_start[row] = s0 + ls;
                        uint4 pv;
                        pv.x = r.evt_off | ((uint32_t)r.evt_len << 16); pv.y = r.sym_off | ((uint32_t)r.sym_len << 16);
                        pv.z = r.dso_off | ((uint32_t)r.dso_len << 16); pv.w = r.va_len | ((uint32_t)r.vb_len << 16);
                        reinterpret_cast<uint4*>(cols.p.pieces)[row] = pv;
                    }
                } else if (KIND == SRC_GPU) {
                    GpuRow r;
                    code = parse_gpu(line, n + (has_nl ? 1 : 0), cfg, &r);
                    if (code == E_OK) {
                        cols.g.ts[row] = r.ts; cols.g.dur[row] = r.dur; cols.g.bw[row] = r.bw; cols.g.payload[row] = r.payload;
                        cols.g.dev[row] = r.dev; cols.g.tid[row] = r.tid; cols.g.src[row] = r.src; cols.g.dst[row] = r.dst;
                        cols.g.kind[row] = r.kind; cols.g.bwf[row] = r.bw_is_float;
                        cols.g.line_start[row] = s0 + ls; cols.g.noff[row] = r.name_off; cols.g.nlen[row] = r.name_len;
                    }
                } else {
                    NetRow r;
                    code = parse_net(line, n, cfg, &r);
                    if (code == E_OK) {
                        cols.n.ts[row] = r.ts; cols.n.payload[row] = r.payload; cols.n.src[row] = r.src; cols.n.dst[row] = r.dst;
                        cols.n.proto[row] = r.proto;
                    }
                }
            }
            if (code != E_OK) report_error(err, row, code);
        }
        __syncthreads();                                   // every thread is done with buffer b and with lstart
    }
}

}  // namespace sofa
