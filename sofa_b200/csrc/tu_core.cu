// libsofa_b200.so, translation unit 1: context lifecycle, transfers, line index + parse driver, dtype flags,
// statistics.  The kernels live in inst/*.cu (one instantiation per file, compiled in parallel).
#include "ctx.cuh"
#include "kernels_ingest_tma.cuh"

thread_local std::string g_last_error;
static const char* KERNEL_NAMES[K_NKINDS] = {
    "count_lines", "scan", "parse", "row_flags", "csv_emit", "json_emit", "filter",
    "viz_dict", "analyze", "sort", "synth", "json_gather"};

int need_parsed(sofa_ctx* c, int k) {
    if (!c || k < 0 || k > 2) FAIL(SOFA_ERR_ARG, "bad source");
    if (!c->src[k].parsed) FAIL(SOFA_ERR_STATE, "source not parsed");
    CK(cudaSetDevice(c->device));
    c->pending.clear(); c->stage_cur = 0;          // drop read-backs orphaned by an earlier failed call
    return 0;
}

extern "C" {

int sofa_b200_abi_version(void) { return SOFA_B200_ABI_VERSION; }
const char* sofa_b200_last_error(void) { return g_last_error.c_str(); }
const char* sofa_b200_kernel_name(int kind) { return (kind >= 0 && kind < K_NKINDS) ? KERNEL_NAMES[kind] : ""; }

int sofa_b200_error_row(sofa_ctx* c, uint64_t* row, int32_t* code) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    if (row) *row = c->err_row;
    if (code) *code = c->err_code;
    return 0;
}

int sofa_b200_create(sofa_ctx** out, int device) {
    if (!out) FAIL(SOFA_ERR_ARG, "null out");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) FAIL(SOFA_ERR_CUDA, "no CUDA device: libsofa_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) FAIL(SOFA_ERR_ARG, "bad device index");
    CK(cudaSetDevice(device));
    sofa_ctx* c = new sofa_ctx();
    c->device = device;
    memset(&c->stats, 0, sizeof c->stats);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&c->out_ev[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->out_alt_ev[i], cudaEventDisableTiming)); }
    for (int i = 0; i < 11; i++) CK(cudaEventCreateWithFlags(&c->up_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) { CK(cudaEventCreateWithFlags(&c->fetch_ev[i], cudaEventDisableTiming)); c->fetch_used[i] = false; }
    CK(cudaHostAlloc((void**)&c->stage_h, c->stage_cap, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&c->stage_d, c->stage_h, 0));
    CK(cudaMalloc(&c->d_err, 8 * 16));
    CK(cudaMalloc(&c->d_flags, 4 * 8));
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    { int e_attr = set_kernel_attrs(); if (e_attr) { g_last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString((cudaError_t)e_attr); return SOFA_ERR_CUDA; } }
    *out = c;
    return 0;
}

int sofa_b200_destroy(sofa_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamDestroy(c->h2d); }
    if (c->d2h) { cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->d2h); }
    for (int i = 0; i < 2; i++) { if (c->out_ev[i]) cudaEventDestroy(c->out_ev[i]); if (c->out_alt_ev[i]) cudaEventDestroy(c->out_alt_ev[i]); c->out_alt[i].release(); }
    for (int i = 0; i < 11; i++) if (c->up_ev[i]) cudaEventDestroy(c->up_ev[i]);
    for (int i = 0; i < 16; i++) cudaEventDestroy(c->fetch_ev[i]);
    for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto& e : c->pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) if (c->mark[i]) cudaEventDestroy(c->mark[i]);
    for (int k = 0; k < 11; k++) {
        Source& s = c->src[k];
        s.text.release(); s.mhz.release(); s.mask.release(); s.vizhash.release(); s.viz_id.release();
        s.dict_hash.release(); s.dict_off.release(); s.dict_blob.release();
        s.json_main.release(); s.rec_off.release();
        for (auto& g : c->gather) g.release();
        for (auto& b : s.cols) b.release();
        if (k < 3) c->out[k].release();
    }
    c->s2.release();
    for (auto& b : c->held) b.release();
    c->tiles.release(); c->rowlen.release(); c->blockoff.release(); c->scratch.release(); c->ftable.release();
    c->keys.release(); c->rows.release();
    c->sk0.release(); c->sv0.release(); c->sk1.release(); c->sv1.release(); c->hist.release(); c->segs.release();
    c->sums.release(); c->small.release(); c->arena.release(); c->scanpart.release();
    cudaFree(c->d_err); cudaFree(c->d_flags);
    if (c->stage_h) cudaFreeHost(c->stage_h);
    if (c->pin_h) cudaFreeHost(c->pin_h);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

void* sofa_b200_host_alloc(uint64_t nbytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void sofa_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

int sofa_b200_upload(sofa_ctx* c, int k, const void* text, uint64_t nbytes, uint64_t data_off) {
    if (!c || k < 0 || k > 10 || (!text && nbytes)) FAIL(SOFA_ERR_ARG, "upload: bad argument");
    if (data_off > nbytes) FAIL(SOFA_ERR_ARG, "upload: data_off beyond the text");
    CK(cudaSetDevice(c->device));
    Source& s = c->src[k];
    ENSURE(s.text, nbytes + 64);
    if (nbytes) CK(cudaMemcpyAsync(s.text.p, text, nbytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(s.text.as<char>() + nbytes, '\n', 64, c->stream));
    CK_S(stream_sync(c));
    c->stats.h2d_bytes += nbytes;
    s.nbytes = nbytes; s.data_off = data_off; s.uploaded = true; s.parsed = false; s.viz_ready = false; s.n_rows = 0;
    s.mask_stride = 0; s.nkw = 0;
    return 0;
}

// A source file read in pieces (pinned staging buffer of the caller): piece [off, off + n) of a text of `total` bytes.
// The first piece (off == 0) sizes the device buffer, the last one (off + n == total) makes the source ready to parse.
int sofa_b200_upload_chunk(sofa_ctx* c, int k, uint64_t off, const void* piece, uint64_t n, uint64_t total, uint64_t data_off) {
    if (!c || k < 0 || k > 10 || (!piece && n) || off + n > total) FAIL(SOFA_ERR_ARG, "upload_chunk: bad argument");
    if (data_off > total) FAIL(SOFA_ERR_ARG, "upload_chunk: data_off beyond the text");
    CK(cudaSetDevice(c->device));
    Source& s = c->src[k];
    if (off == 0) {
        CK_S(stream_sync(c));
        ENSURE(s.text, total + 64);
        s.uploaded = false; s.parsed = false; s.viz_ready = false; s.n_rows = 0; s.mask_stride = 0; s.nkw = 0;
    } else if (s.text.cap < total + 64) FAIL(SOFA_ERR_STATE, "upload_chunk: pieces must start at offset 0");
    if (n) {
        CK(cudaMemcpyAsync(s.text.as<char>() + off, piece, n, cudaMemcpyHostToDevice, c->stream));
        CK_S(stream_sync(c));                                     // the caller refills its staging buffer next
        c->stats.h2d_bytes += n;
    }
    if (off + n == total) {
        CK(cudaMemsetAsync(s.text.as<char>() + total, '\n', 64, c->stream));
        CK_S(stream_sync(c));
        s.nbytes = total; s.data_off = data_off; s.uploaded = true;
    }
    return 0;
}

// Park a copy of output `which` (0 csv, 1 report.js records without their leading ", ") on the device; read it back in
// pieces later (the host writes report.js in series order, not in the order the series are produced).
int sofa_b200_hold_output(sofa_ctx* c, int which, int32_t* handle) {
    if (!c || which < 0 || which > 1 || !handle) FAIL(SOFA_ERR_ARG, "hold_output: bad argument");
    CK(cudaSetDevice(c->device));
    uint64_t n = c->out_n[which], skip = 0;
    if (which == 1) { skip = n >= 2 ? 2 : n; n -= skip; }
    DevBuf b;
    if (n) {
        if (b.ensure(n)) FAIL(SOFA_ERR_CUDA, "cudaMalloc failed (held output)");
        CK(cudaMemcpyAsync(b.p, c->out_ptr[which] + skip, n, cudaMemcpyDeviceToDevice, c->stream));
        CK_S(stream_sync(c));
    }
    int32_t h = -1;
    for (size_t i = 0; i < c->held.size(); i++) if (!c->held[i].p && c->held_n[i] == ~0ull) { h = (int32_t)i; break; }
    if (h < 0) { c->held.push_back(DevBuf()); c->held_n.push_back(~0ull); h = (int32_t)c->held.size() - 1; }
    c->held[h] = b; c->held_n[h] = n;
    *handle = h;
    return 0;
}
int sofa_b200_hold_fetch(sofa_ctx* c, int32_t handle, uint64_t off, uint64_t n, void* host_out) {
    if (!c || handle < 0 || (size_t)handle >= c->held.size() || c->held_n[handle] == ~0ull || (!host_out && n))
        FAIL(SOFA_ERR_ARG, "hold_fetch: bad handle");
    if (off + n > c->held_n[handle]) FAIL(SOFA_ERR_CAP, "hold_fetch: beyond the held output");
    CK(cudaSetDevice(c->device));
    if (n) {
        CK(cudaMemcpyAsync(host_out, c->held[handle].as<char>() + off, n, cudaMemcpyDeviceToHost, c->stream));
        CK_S(stream_sync(c));
        c->stats.d2h_bytes += n;
    }
    return 0;
}
int sofa_b200_hold_release(sofa_ctx* c, int32_t handle) {
    if (!c || handle < -1 || (handle >= 0 && (size_t)handle >= c->held.size())) FAIL(SOFA_ERR_ARG, "hold_release: bad handle");
    CK(cudaSetDevice(c->device));
    for (size_t i = 0; i < c->held.size(); i++)
        if (handle < 0 || (size_t)handle == i) { c->held[i].release(); c->held_n[i] = ~0ull; }
    return 0;
}

static uint32_t pick_tile_bytes(int k) { return (k == SRC_NET || k >= 3) ? 16384u : 24576u; }

int sofa_b200_parse(sofa_ctx* c, int k, const sofa_ingest_params_t* p, uint64_t* n_rows) {
    if (!c || k < 0 || k > 10 || !p) FAIL(SOFA_ERR_ARG, "parse: bad argument");
    Source& s = c->src[k];
    if (!s.uploaded) FAIL(SOFA_ERR_STATE, "parse: nothing uploaded for this source");
    CK(cudaSetDevice(c->device));
    c->pending.clear(); c->stage_cur = 0;
    IngestCfg cfg;
    cfg.time_base = p->time_base; cfg.t_offset = p->t_offset; cfg.absolute_timestamp = p->absolute_timestamp;
    cfg.perf_has_cycles = p->perf_has_cycles; cfg.n_mhz = p->n_mhz > 0 ? p->n_mhz : 1;
    cfg.ts_rescale = p->ts_rescale; cfg.dt_rescale = p->dt_rescale; cfg.payload_unit = p->payload_unit;
    cfg.n_cudaproc = p->n_cudaproc;
    cfg.idx_start = p->idx[0]; cfg.idx_dur = p->idx[1]; cfg.idx_size = p->idx[2]; cfg.idx_thr = p->idx[3];
    cfg.idx_ctx = p->idx[4]; cfg.idx_stream = p->idx[5]; cfg.idx_name = p->idx[6]; cfg.idx_srcctx = p->idx[7];
    cfg.idx_dstctx = p->idx[8];
    cfg.api_mode = p->api_mode == 1 ? 1 : 0;
    cfg.csv_mode = (k <= 2 && p->api_mode == 3) ? 1 : 0;
    cfg.side_table = k >= 6 ? k - 6 : (k >= 3 ? k - 3 : 0);
    if (k >= 3 && k < 6 && p->api_mode == 2) cfg.side_table = 3;          // vmstat.csv: bi / bo / in / cs out of the name column
    if (k == SRC_GPU && !cfg.csv_mode && (cfg.idx_start < 0 || cfg.idx_dur < 0 || cfg.idx_name < 0))
        FAIL(SOFA_ERR_ARG, "parse: gpu trace needs Start, Duration and Name columns");
    ENSURE(s.mhz, 16 * (size_t)cfg.n_mhz + 16);
    if (p->n_mhz > 0 && p->mhz_xp && p->mhz_fp) {
        CK_S(small_h2d(c, s.mhz.p, p->mhz_xp, 8 * (size_t)cfg.n_mhz));
        CK_S(small_h2d(c, s.mhz.as<double>() + cfg.n_mhz, p->mhz_fp, 8 * (size_t)cfg.n_mhz));
    } else {
        double d[2] = {0.0, 3000.0};
        CK_S(small_h2d(c, s.mhz.p, &d[0], 8));
        CK_S(small_h2d(c, s.mhz.as<double>() + 1, &d[1], 8));
        CK_S(stream_sync(c));
    }
    cfg.mhz_xp = s.mhz.as<double>(); cfg.mhz_fp = s.mhz.as<double>() + cfg.n_mhz;
    s.cfg = cfg;
    s.parsed = false; s.viz_ready = false; s.mask_stride = 0;
    s.flags_stride[0] = s.flags_stride[1] = 0;
    if (wait_upload(c, k)) return SOFA_ERR_CUDA;
    uint64_t begin = s.data_off, end = s.nbytes;
    if (end <= begin) { s.n_rows = 0; s.parsed = true; if (n_rows) *n_rows = 0; return 0; }
    uint32_t tile = pick_tile_bytes(k);
    uint64_t ntiles = cdiv(end - begin, tile);
    if (ntiles > 0x7fffffffull) FAIL(SOFA_ERR_ARG, "parse: source too large for one call");
    ENSURE(c->tiles, (ntiles + 2) * 8);
    if (reset_err(c)) return SOFA_ERR_CUDA;
    if (cfg.csv_mode) {
        // rows of a trace CSV are lines once the newlines inside quoted names are masked (in place, on the device copy)
        uint64_t nchunks = cdiv(end - begin, CSVFIX_CHUNK);
        ENSURE(c->blockoff, (nchunks + 2) * 8);
        CK_S(csv_mask_newlines(c, s.text.as<char>(), begin, end, c->blockoff.as<uint64_t>(), nchunks, launch_scan));
    }
    LAUNCH(c, K_COUNT, launch_count_lines((unsigned)ntiles, c->stream, s.text.as<char>(), begin, end, tile, c->tiles.as<uint64_t>(), c->d_flags));
    CK_S(launch_scan(c, c->tiles.as<uint64_t>(), ntiles));
    uint64_t total = 0; unsigned int cr = 0;
    CK_S(small_d2h(c, &total, c->tiles.as<uint64_t>() + ntiles, 8));
    CK_S(small_d2h(c, &cr, c->d_flags, 4));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    if (cr) { c->err_row = 0; c->err_code = E_CR; FAIL(SOFA_ERR_ROW, "parse: carriage return in trace text (not supported)"); }
    if (alloc_cols(c, k, total)) return SOFA_ERR_CUDA;
    if (k >= 6 && total) CK(cudaMemsetAsync(s.cols[RC_NF].p, 0, total * 4, c->stream));
    s.n_rows = total;
    uint32_t max_lines = tile / 8;
    size_t smem = (((size_t)tile + OVERLAP_BYTES + 48) & ~(size_t)15) + (size_t)max_lines * 4;
    AllCols cols = all_cols(c, (k >= 3 && k < 6) ? k : 3, k >= 6 ? k : 6);
    const int tpl = k >= 6 ? SRC_RAW : (k >= 3 ? SRC_SIDE : k);
    // SOFA_B200_PARSE_TMA=1: persistent blocks with two tile buffers filled by cp.async.bulk (kernels_ingest_tma.cuh).  Measured
    // on B200 at 100 M events: 52.6 ms against 46.0 ms for the plain kernel below (64 / 128 registers against 57 / 108 cost
    // more than the overlapped staging gives back; the parse is issue-bound, not load-bound) - so it is not the default.
    static const bool use_tma = getenv("SOFA_B200_PARSE_TMA") && atoi(getenv("SOFA_B200_PARSE_TMA")) != 0;
    if (k <= 2 && cfg.csv_mode) {
        LAUNCH(c, K_PARSE, launch_parse_csv(k, (unsigned)ntiles, smem, c->stream, s.text.as<char>(), begin, end, tile, max_lines,
                                            c->tiles.as<uint64_t>(), cfg, cols, c->d_err));
    } else if (k <= 2 && use_tma) {
        // trace sources: persistent blocks, two tile buffers filled by cp.async.bulk (kernels_ingest_tma.cuh)
        const size_t buf = ((size_t)tile + TMA_OVERLAP_STAGED + TMA_STAGE_PAD + 127) & ~(size_t)127;
        const size_t smem_tma = 2 * buf + (size_t)max_lines * 2 + 16;
        const int occ = parse_tma_occupancy(k, smem_tma);
        const uint64_t resident = (uint64_t)c->sm_count * (uint64_t)occ;
        const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
        LAUNCH(c, K_PARSE, launch_parse_tma(k, grid, smem_tma, c->stream, s.text.as<char>(), begin, end, tile, max_lines, ntiles,
                                            c->tiles.as<uint64_t>(), cfg, cols, c->d_err));
    } else {
        LAUNCH(c, K_PARSE, launch_parse(tpl, (unsigned)ntiles, smem, c->stream, s.text.as<char>(), begin, end, tile, max_lines,
                                        c->tiles.as<uint64_t>(), cfg, cols, c->d_err));
    }
    int r = check_kernel_error(c, "parse");
    if (r) return r;
    s.parsed = true;
    if (n_rows) *n_rows = total;
    return 0;
}


int sofa_b200_row_flags(sofa_ctx* c, int k, uint32_t stride, uint32_t* flags) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0 || !flags) FAIL(SOFA_ERR_ARG, "row_flags: bad argument");
    Source& s = c->src[k];
    *flags = 0;
    if (k == SRC_PERF) return 0;
    for (int i = 0; i < 2; i++) if (s.flags_stride[i] == stride) { *flags = s.flags_val[i]; return 0; }
    if (s.n_rows == 0) return 0;
    uint64_t ncand = cdiv(s.n_rows, stride);
    CK(cudaMemsetAsync(c->d_flags + 1, 0, 4, c->stream));
    SrcView v = make_view(c, k, false);
    if (k == SRC_GPU) LAUNCH(c, K_FLAGS, (k_row_flags<SRC_GPU><<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(v, ncand, stride, c->d_flags + 1)));
    else LAUNCH(c, K_FLAGS, (k_row_flags<SRC_NET><<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(v, ncand, stride, c->d_flags + 1)));
    CK_S(small_d2h(c, flags, c->d_flags + 1, 4));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    { int slot = s.flags_stride[0] == 0 ? 0 : 1; s.flags_stride[slot] = stride; s.flags_val[slot] = *flags; }
    return 0;
}


int sofa_b200_fetch(sofa_ctx* c, int which, void* host_out, uint64_t cap) {
    if (!c || which < 0 || which > 2 || !host_out) FAIL(SOFA_ERR_ARG, "fetch: bad argument");
    CK(cudaSetDevice(c->device));
    uint64_t n = c->out_n[which], skip = 0;
    if (which == 1) { skip = n >= 2 ? 2 : n; n -= skip; }
    if (cap < n) FAIL(SOFA_ERR_CAP, "fetch: buffer too small");
    if (n) {
        const char* srcp = which < 2 ? c->out_ptr[which] : c->out[which].as<char>();
        CK(cudaMemcpyAsync(host_out, srcp + skip, n, cudaMemcpyDeviceToHost, c->stream));
        CK_S(stream_sync(c));
        c->stats.d2h_bytes += n;
    }
    return 0;
}

int sofa_b200_fetch_range(sofa_ctx* c, int which, uint64_t off, uint64_t n, void* host_out) {
    if (!c || which < 0 || which > 1 || (!host_out && n)) FAIL(SOFA_ERR_ARG, "fetch_range: bad argument");
    CK(cudaSetDevice(c->device));
    if (off + n > c->out_n[which]) FAIL(SOFA_ERR_CAP, "fetch_range: beyond the output");
    if (n) {
        CK(cudaMemcpyAsync(host_out, c->out_ptr[which] + off, n, cudaMemcpyDeviceToHost, c->stream));
        CK_S(stream_sync(c));
        c->stats.d2h_bytes += n;
    }
    return 0;
}

int sofa_b200_emit_block_end(sofa_ctx* c, uint64_t block, uint64_t* end_bytes) {
    if (!c || !end_bytes) FAIL(SOFA_ERR_ARG, "emit_block_end: bad argument");
    CK(cudaSetDevice(c->device));
    if ((block + 1) * 8 > c->tiles.cap) FAIL(SOFA_ERR_ARG, "emit_block_end: no such block");
    unsigned long long v = 0;
    CK_S(small_d2h(c, &v, c->tiles.as<unsigned long long>() + block, 8));
    CK_S(stream_sync(c));
    if ((v >> 62) != 2) FAIL(SOFA_ERR_STATE, "emit_block_end: the look-back state of the last emit is gone");
    *end_bytes = v & ((1ull << 62) - 1);
    return 0;
}

int sofa_b200_upload_async(sofa_ctx* c, int k, const void* text, uint64_t nbytes, uint64_t data_off) {
    if (!c || k < 0 || k > 10 || (!text && nbytes)) FAIL(SOFA_ERR_ARG, "upload_async: bad argument");
    if (data_off > nbytes) FAIL(SOFA_ERR_ARG, "upload_async: data_off beyond the text");
    CK(cudaSetDevice(c->device));
    Source& s = c->src[k];
    if (nbytes + 64 > s.text.cap) { CK_S(stream_sync(c)); CK(cudaStreamSynchronize(c->h2d)); ENSURE(s.text, nbytes + 64); }
    // the copy must not overtake kernels of an earlier pass that still read this buffer
    CK(cudaEventRecord(c->up_ev[k], c->stream));
    CK(cudaStreamWaitEvent(c->h2d, c->up_ev[k], 0));
    if (nbytes) CK(cudaMemcpyAsync(s.text.p, text, nbytes, cudaMemcpyHostToDevice, c->h2d));
    CK(cudaMemsetAsync(s.text.as<char>() + nbytes, '\n', 64, c->h2d));
    CK(cudaEventRecord(c->up_ev[k], c->h2d));
    c->up_pending[k] = true;
    c->stats.h2d_bytes += nbytes;
    s.nbytes = nbytes; s.data_off = data_off; s.uploaded = true; s.parsed = false; s.viz_ready = false; s.n_rows = 0;
    s.mask_stride = 0; s.nkw = 0;
    return 0;
}

int sofa_b200_fetch_async(sofa_ctx* c, int which, void* host_out, uint64_t cap, int32_t* ticket) {
    if (!c || which < 0 || which > 1 || !host_out) FAIL(SOFA_ERR_ARG, "fetch_async: output buffer 0 or 1");
    CK(cudaSetDevice(c->device));
    uint64_t n = c->out_n[which], skip = 0;
    if (which == 1) { skip = n >= 2 ? 2 : n; n -= skip; }
    if (cap < n) FAIL(SOFA_ERR_CAP, "fetch_async: buffer too small");
    int t = c->fetch_next;
    c->fetch_next = (c->fetch_next + 1) & 15;
    if (c->fetch_used[t]) CK(cudaEventSynchronize(c->fetch_ev[t]));
    if (n) CK(cudaMemcpyAsync(host_out, c->out_ptr[which] + skip, n, cudaMemcpyDeviceToHost, c->d2h));
    CK(cudaEventRecord(c->out_ev[which], c->d2h));
    CK(cudaEventRecord(c->fetch_ev[t], c->d2h));
    c->out_busy[which] = true;
    c->fetch_used[t] = true;
    c->async_fetches++;
    c->stats.d2h_bytes += n;
    if (ticket) *ticket = t;
    return 0;
}
int sofa_b200_wait_fetch(sofa_ctx* c, int32_t ticket) {
    if (!c || ticket < 0 || ticket > 15) FAIL(SOFA_ERR_ARG, "wait_fetch: bad ticket");
    CK(cudaSetDevice(c->device));
    if (c->fetch_used[ticket]) CK(cudaEventSynchronize(c->fetch_ev[ticket]));
    return 0;
}
int sofa_b200_sync(sofa_ctx* c) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->h2d));
    CK_S(stream_sync(c));
    CK(cudaStreamSynchronize(c->d2h));
    c->arena_cur = 0; c->arena_gen++; c->async_fetches = 0;
    return 0;
}
int sofa_b200_set_output_arena(sofa_ctx* c, uint64_t nbytes) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK_S(stream_sync(c));
    CK(cudaStreamSynchronize(c->d2h));
    c->arena_gen++;
    if (nbytes == 0) { c->arena.release(); c->arena_on = false; c->arena_cur = 0; return 0; }
    ENSURE(c->arena, nbytes);
    c->arena_on = true; c->arena_cur = 0;
    return 0;
}

int sofa_b200_get_column(sofa_ctx* c, int k, int col, void* host_out, uint64_t cap_bytes) {
    int r = need_parsed(c, k); if (r) return r;
    Source& s = c->src[k];
    int slot = -1; size_t w = 0;
    if (k == SRC_PERF) {
        switch (col) { case 0: slot = PC_TS; w = 8; break; case 1: slot = PC_EVENT; w = 8; break; case 2: slot = PC_DUR; w = 8; break;
                       case 9: slot = PC_PID; w = 4; break; case 10: slot = PC_TID; w = 4; break; }
    } else if (k == SRC_GPU) {
        switch (col) { case 0: slot = GC_TS; w = 8; break; case 2: slot = GC_DUR; w = 8; break; case 6: slot = GC_BW; w = 8; break;
                       case 5: slot = GC_PAYLOAD; w = 8; break; case 3: slot = GC_DEV; w = 4; break; case 10: slot = GC_TID; w = 4; break;
                       case 7: slot = GC_SRC; w = 4; break; case 8: slot = GC_DST; w = 4; break; case 4: slot = GC_KIND; w = 1; break; }
    } else {
        switch (col) { case 0: slot = NC_TS; w = 8; break; case 5: slot = NC_PAYLOAD; w = 8; break; case 7: slot = NC_SRC; w = 8; break;
                       case 8: slot = NC_DST; w = 8; break; case 4: slot = NC_PROTO; w = 1; break; }
    }
    if (slot < 0) FAIL(SOFA_ERR_ARG, "get_column: no such stored column for this source");
    uint64_t nb = s.n_rows * w;
    if (cap_bytes < nb) FAIL(SOFA_ERR_CAP, "get_column: buffer too small");
    if (nb) { CK(cudaMemcpyAsync(host_out, s.cols[slot].p, nb, cudaMemcpyDeviceToHost, c->stream)); CK_S(stream_sync(c)); }
    return 0;
}

int sofa_b200_mark(sofa_ctx* c, int which) {
    if (!c || which < 0 || which > 1) FAIL(SOFA_ERR_ARG, "mark: bad argument");
    CK(cudaSetDevice(c->device));
    if (!c->mark[which]) CK(cudaEventCreate(&c->mark[which]));
    CK(cudaEventRecord(c->mark[which], c->stream));
    return 0;
}
int sofa_b200_elapsed_ms(sofa_ctx* c, double* ms) {
    if (!c || !ms || !c->mark[0] || !c->mark[1]) FAIL(SOFA_ERR_ARG, "elapsed_ms: marks not set");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->mark[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->mark[0], c->mark[1]));
    *ms = f;
    return 0;
}

int sofa_b200_set_debug(sofa_ctx* c, int key, uint64_t value) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    if (key == SOFA_DEBUG_NAME_HASH_MASK) { c->debug_hash_mask = value; return 0; }
    FAIL(SOFA_ERR_ARG, "set_debug: unknown key");
}

int sofa_b200_set_profiling(sofa_ctx* c, int on) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    c->profiling = on != 0;
    return 0;
}
int sofa_b200_get_stats(sofa_ctx* c, sofa_stats_t* out) {
    if (!c || !out) FAIL(SOFA_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    *out = c->stats;
    return 0;
}
int sofa_b200_reset_stats(sofa_ctx* c) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    memset(&c->stats, 0, sizeof c->stats);
    return 0;
}

int sofa_b200_text_fetch(sofa_ctx* c, int k, void* host_out, uint64_t cap) {
    if (!c || k < 0 || k > 2 || !host_out) FAIL(SOFA_ERR_ARG, "text_fetch: bad argument");
    Source& s = c->src[k];
    if (!s.uploaded) FAIL(SOFA_ERR_STATE, "text_fetch: no text");
    if (cap < s.nbytes) FAIL(SOFA_ERR_CAP, "text_fetch: buffer too small");
    CK(cudaSetDevice(c->device));
    if (s.nbytes) { CK(cudaMemcpyAsync(host_out, s.text.p, s.nbytes, cudaMemcpyDeviceToHost, c->stream)); CK_S(stream_sync(c)); }
    return 0;
}

static __global__ void k_add_u64(uint64_t* a, uint64_t n, uint64_t delta) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += delta;
}
// (slot, element bytes) of the parsed columns of a trace source, in a fixed order
static int shard_layout(int k, int* slots, uint32_t* elem) {
    if (k == SRC_PERF) {
        const int s[] = {PC_TS, PC_EVENT, PC_DUR, PC_PID, PC_TID, PC_LS, PC_PIECES}; const uint32_t e[] = {8, 8, 8, 4, 4, 8, 16};
        for (int i = 0; i < 7; i++) { slots[i] = s[i]; elem[i] = e[i]; }
        return 7;
    }
    if (k == SRC_GPU) {
        const int s[] = {GC_TS, GC_DUR, GC_BW, GC_PAYLOAD, GC_DEV, GC_TID, GC_SRC, GC_DST, GC_KIND, GC_BWF, GC_LS, GC_NOFF, GC_NLEN};
        const uint32_t e[] = {8, 8, 8, 8, 4, 4, 4, 4, 1, 1, 8, 2, 2};
        for (int i = 0; i < 13; i++) { slots[i] = s[i]; elem[i] = e[i]; }
        return 13;
    }
    const int s[] = {NC_TS, NC_PAYLOAD, NC_SRC, NC_DST, NC_PROTO}; const uint32_t e[] = {8, 8, 8, 8, 1};
    for (int i = 0; i < 5; i++) { slots[i] = s[i]; elem[i] = e[i]; }
    return 5;
}

int sofa_b200_shard_export(sofa_ctx* c, int k, sofa_shard_t* out) {
    int r = need_parsed(c, k); if (r) return r;
    if (!out) FAIL(SOFA_ERR_ARG, "shard_export: null out");
    CK_S(stream_sync(c));
    Source& s = c->src[k];
    memset(out, 0, sizeof *out);
    int slots[16]; uint32_t elem[16];
    out->n_cols = shard_layout(k, slots, elem);
    out->n_rows = s.n_rows; out->text_bytes = s.nbytes; out->data_off = s.data_off; out->text = s.text.p;
    for (int i = 0; i < out->n_cols; i++) { out->cols[i] = s.cols[slots[i]].p; out->col_elem[i] = elem[i]; }
    return 0;
}

/* rank 0 of a sharded trace: fresh buffers for `total_rows` rows and `total_text` text bytes of source k (the shard it
 * parsed itself stays valid until sofa_b200_shard_commit); the caller fills them shard by shard (its own shard with a
 * device copy, the others with NCCL receives straight into place) and commits. */
int sofa_b200_shard_reserve(sofa_ctx* c, int k, uint64_t total_rows, uint64_t total_text, sofa_shard_t* out) {
    int r = need_parsed(c, k); if (r) return r;
    if (!out) FAIL(SOFA_ERR_ARG, "shard_reserve: null out");
    CK_S(stream_sync(c));
    memset(out, 0, sizeof *out);
    int slots[16]; uint32_t elem[16];
    out->n_cols = shard_layout(k, slots, elem);
    // (the buffers are kept between calls: after a commit they hold the previous shard's storage, which the next parse of
    //  this source reuses through the swap - a steady-state step allocates nothing)
    uint64_t m = total_rows ? total_rows : 1;
    ENSURE(c->gather[0], total_text + 64);
    for (int i = 0; i < out->n_cols; i++) ENSURE(c->gather[1 + i], m * elem[i] + 64);
    out->n_rows = total_rows; out->text_bytes = total_text; out->text = c->gather[0].p;
    for (int i = 0; i < out->n_cols; i++) { out->cols[i] = c->gather[1 + i].p; out->col_elem[i] = elem[i]; }
    return 0;
}

/* shard_rows[i] / shard_text[i]: rows and text bytes of shard i as laid out in the reserved buffers, in rank order.
 * The row -> line offsets of shard i are rebased onto its position in the gathered text; the gathered buffers become
 * the source's parsed state (as if one context had parsed the whole file). */
int sofa_b200_shard_commit(sofa_ctx* c, int k, int32_t nshards, const uint64_t* shard_rows, const uint64_t* shard_text) {
    int r = need_parsed(c, k); if (r) return r;
    if (nshards < 1 || !shard_rows || !shard_text) FAIL(SOFA_ERR_ARG, "shard_commit: bad argument");
    Source& s = c->src[k];
    int slots[16]; uint32_t elem[16];
    int nc = shard_layout(k, slots, elem);
    uint64_t rows = 0, text = 0;
    int ls_slot = k == SRC_PERF ? PC_LS : (k == SRC_GPU ? GC_LS : -1);
    for (int i = 0; i < nshards; i++) {
        if (ls_slot >= 0 && text && shard_rows[i]) {
            int gi = 0;
            for (int q = 0; q < nc; q++) if (slots[q] == ls_slot) gi = 1 + q;
            uint64_t* ls = c->gather[gi].as<uint64_t>() + rows;
            k_add_u64<<<(unsigned)cdiv(shard_rows[i], 256), 256, 0, c->stream>>>(ls, shard_rows[i], text);
        }
        rows += shard_rows[i]; text += shard_text[i];
    }
    CK(cudaMemsetAsync(c->gather[0].as<char>() + text, '\n', 64, c->stream));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    std::swap(s.text, c->gather[0]);
    for (int q = 0; q < nc; q++) std::swap(s.cols[slots[q]], c->gather[1 + q]);
    s.n_rows = rows; s.nbytes = text; s.data_off = 0; s.parsed = true; s.viz_ready = false; s.mask_stride = 0; s.nkw = 0; s.gather_ok = false;
    return 0;
}

/* one logical trace split over several contexts (ranks): the dtype flags of a column are a property of the whole
 * file, so every shard formats with the OR of all shards' flags */
int sofa_b200_set_global_flags(sofa_ctx* c, int k, uint32_t stride, uint32_t flags) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "set_global_flags: stride must be >= 1");
    Source& s = c->src[k];
    int slot = (s.flags_stride[0] == stride || s.flags_stride[0] == 0) ? 0 : 1;
    s.flags_stride[slot] = stride; s.flags_val[slot] = flags;
    if (s.n_rows == 0) { s.flags_stride[slot] = stride; }
    return 0;
}

}  // extern "C"

int fmask_for(sofa_ctx* c, int k, uint32_t stride, uint32_t* fmask) {
    if (k == SRC_PERF) { *fmask = 0x7; return 0; }
    uint32_t fl = 0;
    int r = sofa_b200_row_flags(c, k, stride, &fl);
    if (r) return r;
    if (k == SRC_GPU) { *fmask = 0x5u | ((fl & 1) ? 0x40u : 0u); return 0; }
    if (fl & 2) {
        if (!(fl & 4)) FAIL(SOFA_ERR_ROW, "every selected packet row is empty: the reference raises (0 columns)");
        *fmask = 0x17ff;
    } else *fmask = 0x45;
    return 0;
}
