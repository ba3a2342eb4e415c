// Ingest: newline index + per-record parse into SoA columns.
//
// The text of one source sits in HBM as one byte buffer.  It is cut into fixed-size byte
// tiles; a line belongs to the tile its first byte falls in.  k_count_lines counts the
// line starts of every tile (128-bit loads, SWAR newline compare), a scan turns counts
// into first-row indices, and k_parse<KIND> stages each tile (+ an overlap window for the
// last line) in shared memory with coalesced 128-bit loads, finds the line starts with a
// block scan and parses one record per thread out of shared memory, writing the columns
// with coalesced stores (row i of the tile -> thread i).
#pragma once
#include "rows.cuh"
#include "side.cuh"
#include "devutil.cuh"

namespace sofa {

struct PerfCols {
    double *ts, *event, *dur; int32_t *pid, *tid; uint64_t* line_start; uint16_t* pieces;   // 8 u16 per row (one 16-byte vector)
};
struct GpuCols {
    double *ts, *dur, *bw; int64_t* payload; int32_t *dev, *tid, *src, *dst; uint8_t *kind, *bwf;
    uint64_t* line_start; uint16_t *noff, *nlen;
};
struct NetCols { double* ts; int64_t *payload, *src, *dst; uint8_t* proto; };

struct SideCols { double *ts, *event, *dur, *dev, *bw; int32_t* ratios; };      // ratios: 5 per row

// raw recorder side files (side.cuh): f = RAW_NF doubles per line, i = 4 int64 per line, nf = field count
// (-1: the line says "failed" - nvidia-smi query error)
struct RawCols { double* f; int64_t* i; int32_t* nf; };

struct AllCols { PerfCols p; GpuCols g; NetCols n; SideCols s; RawCols r; };

constexpr uint32_t CSVFIX_CHUNK = 4096;      // bytes per warp of the trace-CSV quote-parity pre-pass (below)

__device__ __forceinline__ uint32_t newline_mask4(uint32_t w) {       // 0xff in every byte equal to '\n'
    return __vcmpeq4(w, 0x0a0a0a0au);
}

#ifdef SOFA_K_COUNT
// Line starts owned by tile t: `begin` itself (first tile) and q+1 for every newline at
// q in [max(b0-1,begin), min(b1-1,end-1)).
__global__ void __launch_bounds__(PARSE_TPB)
k_count_lines(const char* __restrict__ text, uint64_t begin, uint64_t end, uint32_t tile_bytes,
              uint64_t* __restrict__ tile_counts, unsigned int* __restrict__ cr_flag) {
    __shared__ uint32_t sm[33];
    uint64_t t = blockIdx.x;
    uint64_t b0 = begin + t * (uint64_t)tile_bytes;
    uint64_t b1 = b0 + tile_bytes < end ? b0 + tile_bytes : end;
    uint64_t qlo = (b0 == begin) ? begin : b0 - 1;
    uint64_t qhi = (b1 - 1 < end - 1) ? b1 - 1 : end - 1;
    uint32_t cnt = 0;
    bool cr = false;
    uint64_t a0 = qlo & ~15ull;
    const uint4* t4 = reinterpret_cast<const uint4*>(text);
    // two 128-bit loads in flight per thread and iteration (the text buffer is padded past `end`)
    const uint64_t step = 16ull * blockDim.x;
    for (uint64_t p0 = a0 + 16ull * threadIdx.x; p0 < qhi; p0 += 2 * step) {
        const uint64_t pp[2] = {p0, p0 + step};
        uint4 vv[2];
        vv[0] = t4[pp[0] >> 4];
        vv[1] = pp[1] < qhi ? t4[pp[1] >> 4] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint64_t p = pp[u];
            if (p >= qhi) continue;
            uint32_t w[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t m = newline_mask4(w[k]);
                uint32_t c = __vcmpeq4(w[k], 0x0d0d0d0du);
                uint64_t q = p + 4 * k;
                if (q < qlo || q + 4 > qhi) {            // partial word: drop bytes outside [qlo, qhi)
                    uint32_t keep = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (q + j >= qlo && q + j < qhi) keep |= 0xffu << (8 * j);
                    m &= keep; c &= keep;
                }
                cnt += __popc(m) >> 3;
                cr |= (c != 0);
            }
        }
    }
    if (threadIdx.x == 0 && b0 == begin && begin < end) cnt += 1;
    uint32_t total;
    block_excl_scan(cnt, &total, sm);
    if (threadIdx.x == 0) tile_counts[t] = total;
    if (cr) atomicOr(cr_flag, 1u);
}

// ---- trace CSV re-ingest: newlines inside quoted name fields (nvprof kernel names keep their '\n') become 0x1f, so
// that rows are lines for the tile machinery above.  Quote parity of a byte = number of '"' before it, mod 2:
// counted per chunk, prefix-summed over the chunks, then resolved inside every chunk by one warp with ballots.
__global__ void __launch_bounds__(256) k_csv_quote_count(const char* __restrict__ text, uint64_t begin, uint64_t end, uint64_t* counts) {
    const uint64_t chunk = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint64_t a = begin + chunk * CSVFIX_CHUNK;
    if (a >= end) return;
    const uint64_t b = a + CSVFIX_CHUNK < end ? a + CSVFIX_CHUNK : end;
    uint32_t c = 0;
    for (uint64_t p = a + (threadIdx.x & 31); p < b; p += 32) c += text[p] == '"';
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) counts[chunk] = c;
}
__global__ void __launch_bounds__(256) k_csv_mask_newlines(char* text, uint64_t begin, uint64_t end, const uint64_t* prefix) {
    const uint64_t chunk = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint64_t a = begin + chunk * CSVFIX_CHUNK;
    if (a >= end) return;
    const uint64_t b = a + CSVFIX_CHUNK < end ? a + CSVFIX_CHUNK : end;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t inq = (uint32_t)(prefix[chunk] & 1);
    for (uint64_t p0 = a; p0 < b; p0 += 32) {
        const uint64_t p = p0 + lane;
        const char ch = p < b ? text[p] : 0;
        const unsigned qb = __ballot_sync(0xffffffffu, ch == '"');
        const uint32_t mine = inq ^ (__popc(qb & ((1u << lane) - 1u)) & 1u);
        if (ch == '\n' && mine) text[p] = 0x1f;
        inq ^= __popc(qb) & 1u;
    }
}
#endif  // SOFA_K_COUNT

// In-place exclusive scan of a[0..n) (single block); a[n] = total.  (static: every translation unit that scans gets its own copy)
static __global__ void __launch_bounds__(1024) k_scan_u64(uint64_t* a, uint64_t n) {
    __shared__ uint64_t sm[33];
    constexpr int PER = 8;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n; base += 1024ull * PER) {
        uint64_t v[PER];
        uint64_t i0 = base + (uint64_t)threadIdx.x * PER;
        uint64_t s = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { v[k] = (i0 + k < n) ? a[i0 + k] : 0; s += v[k]; }
        uint64_t total;
        uint64_t ex = block_excl_scan64(s, &total, sm) + carry;
#pragma unroll
        for (int k = 0; k < PER; k++) { if (i0 + k < n) a[i0 + k] = ex; ex += v[k]; }
        carry += total;
    }
    if (threadIdx.x == 0) a[n] = carry;
}

// Large scans in three launches: every block scans 8192 elements in place and leaves its total in
// part[block]; k_scan_u64 scans the totals; k_scan_add adds them back (and stores a[n] = grand total).
constexpr uint32_t SCAN_BLOCK_ELEMS = 8192;
static __global__ void __launch_bounds__(1024) k_scan_part(uint64_t* a, uint64_t n, uint64_t* part) {
    __shared__ uint64_t sm[33];
    constexpr int PER = 8;
    uint64_t v[PER];
    uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLOCK_ELEMS + (uint64_t)threadIdx.x * PER;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { v[k] = (i0 + k < n) ? a[i0 + k] : 0; s += v[k]; }
    uint64_t total;
    uint64_t ex = block_excl_scan64(s, &total, sm);
#pragma unroll
    for (int k = 0; k < PER; k++) { if (i0 + k < n) a[i0 + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(1024) k_scan_add(uint64_t* a, uint64_t n, const uint64_t* __restrict__ part, uint32_t nparts) {
    uint64_t off = part[blockIdx.x];
    uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLOCK_ELEMS + (uint64_t)threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) if (i0 + k < n) a[i0 + k] += off;
    if (blockIdx.x == 0 && threadIdx.x == 0) a[n] = part[nparts];
}

// CSV = true: the text is a trace CSV (stand-alone `sofa analyze`); a separate instantiation so that the hot raw-log
// kernels carry none of its code (one merged kernel cost 46 -> 52 ms per 100 M events)
// BULK: the tile is brought into shared memory by ONE cp.async.bulk issued by thread 0 (the copy engine moves the 24-32 KB
// while no warp of the block spends issue slots on it; UBLKCP.S.G + SYNCS in SASS) instead of an LDG -> STS loop over all
// threads; the block waits on the copy's mbarrier where it used to wait at the barrier after the loop.
template <int KIND, bool CSV = false, bool BULK = (KIND <= SRC_NET && !CSV)>
__global__ void __launch_bounds__(PARSE_TPB)
k_parse(const char* __restrict__ text, uint64_t begin, uint64_t end, uint32_t tile_bytes, uint32_t max_lines,
        const uint64_t* __restrict__ tile_row_base, IngestCfg cfg, AllCols cols, unsigned long long* err) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t sm[33];
    uint64_t t = blockIdx.x;
    uint64_t b0 = begin + t * (uint64_t)tile_bytes;
    uint64_t b1 = b0 + tile_bytes < end ? b0 + tile_bytes : end;
    uint64_t qlo = (b0 == begin) ? begin : b0 - 1;
    uint64_t qhi = (b1 - 1 < end - 1) ? b1 - 1 : end - 1;
    uint64_t s0 = qlo & ~15ull;
    uint64_t s1 = b1 + OVERLAP_BYTES < end ? b1 + OVERLAP_BYTES : end;
    uint32_t stage_len = (uint32_t)(((s1 + 15) & ~15ull) - s0);
    unsigned char* stage = smem;
    uint32_t* lstart = reinterpret_cast<uint32_t*>(smem + (((size_t)tile_bytes + OVERLAP_BYTES + 48) & ~(size_t)15));
    if constexpr (BULK) {
        __shared__ __align__(8) uint64_t bar;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, stage_len);
            bulk_g2s(stage, text + s0, stage_len, &bar);       // (s0 and stage_len are multiples of 16; the text is padded past `end`)
        }
        mbar_wait(&bar, 0);
    } else {   // coalesced 128-bit staging (buffer is padded past `end`)
        const uint4* g4 = reinterpret_cast<const uint4*>(text + s0);
        uint4* s4 = reinterpret_cast<uint4*>(stage);
        for (uint32_t i = threadIdx.x; i < (stage_len >> 4); i += blockDim.x) s4[i] = g4[i];
        __syncthreads();
    }
    // each thread scans an 8-byte-aligned chunk of the staged bytes for newlines in [qlo, qhi),
    // one 64-bit shared-memory load per 8 bytes (odd word stride: conflict free)
    uint32_t scan_len = (uint32_t)(qhi > s0 ? qhi - s0 : 0);
    uint32_t chunk = ((scan_len + blockDim.x - 1) / blockDim.x + 7) & ~7u;
    if (((chunk >> 3) & 1u) == 0) chunk += 8;
    uint32_t c0 = threadIdx.x * chunk, c1 = c0 + chunk < scan_len ? c0 + chunk : scan_len;
    uint32_t lo_rel = (uint32_t)(qlo - s0);
    uint32_t cnt = 0;
    const uint64_t* st8 = reinterpret_cast<const uint64_t*>(stage);
    for (uint32_t i = c0; i < c1; i += 8) {
        uint32_t m = eq8(st8[i >> 3], '\n');
        if (i < lo_rel) m &= (lo_rel - i >= 8) ? 0u : ~((1u << (lo_rel - i)) - 1u);
        if (i + 8 > c1) m &= (1u << (c1 - i)) - 1u;
        cnt += __popc(m);
    }
    if (threadIdx.x == 0 && b0 == begin && begin < end) cnt += 1;
    uint32_t total;
    uint32_t ex = block_excl_scan(cnt, &total, sm);
    if (total > max_lines) {
        if (threadIdx.x == 0) report_error(err, tile_row_base[t], E_TOO_MANY_LINES);
        return;
    }
    {
        uint32_t k = ex;
        if (threadIdx.x == 0 && b0 == begin && begin < end) lstart[k++] = (uint32_t)(begin - s0);
        for (uint32_t i = c0; i < c1; i += 8) {
            uint32_t m = eq8(st8[i >> 3], '\n');
            if (i < lo_rel) m &= (lo_rel - i >= 8) ? 0u : ~((1u << (lo_rel - i)) - 1u);
            if (i + 8 > c1) m &= (1u << (c1 - i)) - 1u;
            while (m) { lstart[k++] = i + (uint32_t)ctz32(m) + 1; m &= m - 1; }
        }
    }
    __syncthreads();
    uint32_t stage_end = (uint32_t)(s1 - s0);           // staged bytes that belong to the text
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        uint32_t ls = lstart[j], le;
        bool has_nl = true;
        int code = E_OK;
        if (j + 1 < total) {
            le = lstart[j + 1] - 1;
        } else {
            le = ls;
            while (le < stage_end && stage[le] != '\n') le++;
            if (le >= stage_end) {
                if (s1 >= end) has_nl = false;            // last line of the file without '\n'
                else code = E_LINE_TOO_LONG;
            }
        }
        uint64_t row = tile_row_base[t] + j;
        const char* line = reinterpret_cast<const char*>(stage + ls);
        int n = (int)(le - ls);
        if (code == E_OK) {
            if (KIND == SRC_PERF) {
                PerfRow r;
                if constexpr (CSV) code = parse_csv_perf(line, n, &r); else code = parse_perf(line, n, cfg, &r);
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
                if constexpr (CSV) code = parse_csv_gpu(line, n, &r); else code = parse_gpu(line, n + (has_nl ? 1 : 0), cfg, &r);
                if (code == E_OK) {
                    cols.g.ts[row] = r.ts; cols.g.dur[row] = r.dur; cols.g.bw[row] = r.bw; cols.g.payload[row] = r.payload;
                    cols.g.dev[row] = r.dev; cols.g.tid[row] = r.tid; cols.g.src[row] = r.src; cols.g.dst[row] = r.dst;
                    cols.g.kind[row] = r.kind; cols.g.bwf[row] = r.bw_is_float;
                    cols.g.line_start[row] = s0 + ls; cols.g.noff[row] = r.name_off; cols.g.nlen[row] = r.name_len;
                }
            } else if (KIND == SRC_RAW) {
                RawRow r = RawRow();
                if (cfg.side_table == RAW_MPSTAT) code = parse_raw_mpstat(line, n, &r);
                else if (cfg.side_table == RAW_NVSMI) code = parse_raw_nvsmi(line, n, has_nl, &r);
                else if (cfg.side_table == RAW_NETSTAT) code = parse_raw_netstat(line, n, &r);
                else if (cfg.side_table == RAW_VMSTAT) code = parse_raw_vmstat(line, n, &r);
                else code = parse_raw_dmon(line, n, &r);
                if (code == E_OK) {
                    for (int q = 0; q < RAW_NF; q++) cols.r.f[row * RAW_NF + q] = r.f[q];
                    for (int q = 0; q < 4; q++) cols.r.i[row * 4 + q] = r.i[q];
                    cols.r.nf[row] = r.flag ? -1 : r.nf;
                }
            } else if (KIND == SRC_SIDE) {
                SideRow r;
                code = parse_side(line, n, cfg, &r);
                if (code == E_OK) {
                    cols.s.ts[row] = r.ts; cols.s.event[row] = r.event; cols.s.dur[row] = r.dur; cols.s.dev[row] = r.dev;
                    cols.s.bw[row] = r.bw;
                    for (int q = 0; q < 5; q++) cols.s.ratios[row * 5 + q] = r.r[q];
                }
            } else {
                NetRow r;
                if constexpr (CSV) code = parse_csv_net(line, n, &r); else code = parse_net(line, n, cfg, &r);
                if (code == E_OK) {
                    cols.n.ts[row] = r.ts; cols.n.payload[row] = r.payload; cols.n.src[row] = r.src; cols.n.dst[row] = r.dst;
                    cols.n.proto[row] = r.proto;
                }
            }
        }
        if (code != E_OK) report_error(err, row, code);
    }
}

}  // namespace sofa
