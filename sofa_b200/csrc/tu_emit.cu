// libsofa_b200.so, translation unit: text emission driver (output buffers, capacity retry, look-back state) over
// k_emit_fused<KIND, FMT> (instantiated in inst/emit_*.cu).
#include "ctx.cuh"
#include "kernels_gather.cuh"

// report.js main series with fused keyword matching / record offsets (see FusedArgs)
struct JsonFuse { unsigned long long* rec_off; bool keep_output; };

template <int FMT>
static int emit_fused_t(sofa_ctx* c, int k, uint32_t stride, const uint8_t* mask, int bit, uint32_t fmask,
                      int outslot, uint64_t sel_hint, uint64_t* nbytes, uint64_t* nrec, const SideEmit* se = nullptr,
                      const JsonFuse* jf = nullptr) {
    Source& s = c->src[k];
    uint64_t ncand = se ? se->ncand : cdiv(s.n_rows, stride);
    c->out_n[outslot] = 0;
    if (nbytes) *nbytes = 0;
    if (nrec) *nrec = 0;
    if (flip_output(c, outslot)) return SOFA_ERR_CUDA;
    if (ncand == 0) return 0;
    if (ncand >= 0xffffffffull) FAIL(SOFA_ERR_ARG, "emit: too many rows for one call");
    uint64_t text_share = s.nbytes > s.data_off ? s.nbytes - s.data_off : 0;
    if (se) text_share = 0;                      // names are generated, not copied from the text
    const uint32_t* index = nullptr;
    uint64_t nitems = ncand;
    if (mask) {
        // filtered series: compact the selected candidates first so the emit grid is dense
        uint64_t nb = cdiv(ncand, 256);
        ENSURE(c->blockoff, (nb + 2) * 8);
        LAUNCH(c, K_FILTER, (k_mask_count<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>())));
        CK_S(launch_scan(c, c->blockoff.as<uint64_t>(), nb));
        uint64_t cnt = 0;
        CK_S(small_d2h(c, &cnt, c->blockoff.as<uint64_t>() + nb, 8));
        CK_S(stream_sync(c));
        if (cnt == 0) return 0;
        ENSURE(c->rowlen, cnt * 4 + 16);
        LAUNCH(c, K_FILTER, (k_mask_write<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>(), c->rowlen.as<uint32_t>())));
        index = c->rowlen.as<uint32_t>();
        nitems = cnt;
        text_share = (uint64_t)((double)text_share * ((double)cnt / (double)ncand) * 2.0) + (1 << 20);
        sel_hint = cnt;
    }
    uint64_t nblocks = cdiv(nitems, EMIT_TPB);
    ENSURE(c->tiles, (nblocks + 2) * 8);
    SrcView v = make_view(c, k, true);
    if (se) v.side = se->v;
    uint64_t per_row = FMT == FMT_CSV ? 96 : ((FMT == FMT_COMM || FMT == FMT_AUX) ? 64 : 230);
    if (FMT == FMT_COMM) text_share = 0;
    uint64_t cap = sel_hint * per_row + text_share + (1 << 16);
    // shared-memory image per block: sized for 128 typical rows of the format (larger ranges fall back
    // to direct global writes); smaller images = more resident blocks per SM
    // bytes of the per-block image; measured at 100 M events: CSV 24K 37.5 ms, 28K 26.5, 32K 27.8, 40K 30.2; JSON 36K 73.9,
    // 40K 57.5, 44K 57.5 (a block whose rows do not fit falls back to direct global stores - the cliff).  Env overrides for tuning.
    static const uint32_t slot_csv = getenv("SOFA_B200_SLOT_CSV") ? (uint32_t)atoi(getenv("SOFA_B200_SLOT_CSV")) : 32u * 1024u;
    static const uint32_t slot_json = getenv("SOFA_B200_SLOT_JSON") ? (uint32_t)atoi(getenv("SOFA_B200_SLOT_JSON")) : 40u * 1024u;
    uint32_t slot_bytes = (FMT == FMT_COMM || FMT == FMT_AUX) ? 16 * 1024 : (FMT == FMT_JSON ? slot_json : slot_csv);
    if (!se && nitems) {
        // traces with long names (C++ symbols of several hundred bytes): grow the image with the average line
        // length so that a block's 128 rows still fit - the direct-to-global fallback is several times slower
        double avg_line = (double)(s.nbytes > s.data_off ? s.nbytes - s.data_off : 0) / (double)(s.n_rows ? s.n_rows : 1);
        double est_row = (FMT == FMT_JSON ? 200.0 : 70.0) + avg_line;
        uint32_t want = (uint32_t)(est_row * EMIT_TPB * 1.1);
        if (FMT != FMT_COMM && want > slot_bytes + slot_bytes / 8) slot_bytes = (want + 1023u) & ~1023u;     // only when clearly needed
    }
    if (slot_bytes < 4096u) slot_bytes = 4096u;
    if (slot_bytes > 63u * 1024u) slot_bytes = 63u * 1024u;
    for (int attempt = 0; attempt < 6; attempt++, cap *= 2) {
        char* outp = nullptr;
        if (jf && jf->keep_output && !c->arena_on) {
            // the main series stays resident in its own buffer so that the filter series can be gathered from it
            if (c->async_fetches) { CK(cudaStreamSynchronize(c->d2h)); c->async_fetches = 0; }
            ENSURE(s.json_main, cap + 64);
            outp = s.json_main.as<char>();
            c->out_ptr[outslot] = outp;
        } else { int ra = out_alloc(c, outslot, cap, &outp); if (ra) return ra; }
        if (reset_err(c)) return SOFA_ERR_CUDA;
        CK(cudaMemsetAsync(c->tiles.p, 0, (nblocks + 1) * 8, c->stream));
        FusedArgs a;
        a.ncand = nitems; a.stride = stride; a.fmask = fmask; a.mask = nullptr; a.bit = 0; a.index = index;
        a.out = outp; a.cap = cap;
        a.state = c->tiles.as<unsigned long long>(); a.ticket = c->d_flags + 4;
        a.err = c->d_err; a.count = c->d_err + 1; a.total = c->d_err + 11; a.overflow = c->d_flags + 2;
        a.slot_bytes = slot_bytes;
        a.rec_off = jf ? jf->rec_off : nullptr;
        int kk = FMT == FMT_JSON ? K_JSON_WRITE : K_CSV_WRITE;
        unsigned g = (unsigned)nblocks;
        if (FMT == FMT_COMM && k != SRC_GPU) FAIL(SOFA_ERR_ARG, "comm.csv is a gpu-trace output");
        if (FMT == FMT_AUX && !se) FAIL(SOFA_ERR_ARG, "aux rows are a side-source output");
        const int tk = (se || FMT == FMT_AUX) ? SRC_SIDEOUT : k;
        LAUNCH(c, kk, launch_emit(tk, FMT, g, slot_bytes + 32, c->stream, v, a));
        unsigned long long total = 0, cnt = 0; unsigned int ovf = 0;
        CK_S(small_d2h(c, &total, c->d_err + 11, 8));
        CK_S(small_d2h(c, &cnt, c->d_err + 1, 8));
        CK_S(small_d2h(c, &ovf, c->d_flags + 2, 4));
        int r = check_kernel_error(c, FMT == FMT_CSV ? "emit_csv" : "emit_json");
        if (r) return r;
        if (ovf) { if (total + 64 > cap * 2) cap = total / 2 + 64; continue; }
        if (jf && jf->keep_output) {
            s.gather_ptr = outp; s.gather_gen = c->arena_on ? c->arena_gen : ~0ull;
        }
        out_commit(c, outslot, total);
        if (nbytes) *nbytes = total;
        if (nrec) *nrec = cnt;
        return 0;
    }
    FAIL(SOFA_ERR_CAP, "emit: output capacity estimate failed repeatedly");
}

int sofa_emit_fused(sofa_ctx* c, int fmt, int k, uint32_t stride, const uint8_t* mask, int bit, uint32_t fmask,
                    int outslot, uint64_t sel_hint, uint64_t* nbytes, uint64_t* nrec, const SideEmit* se) {
    switch (fmt) {
        case FMT_CSV: return emit_fused_t<FMT_CSV>(c, k, stride, mask, bit, fmask, outslot, sel_hint, nbytes, nrec, se);
        case FMT_JSON: return emit_fused_t<FMT_JSON>(c, k, stride, mask, bit, fmask, outslot, sel_hint, nbytes, nrec, se);
        case FMT_COMM: return emit_fused_t<FMT_COMM>(c, k, stride, mask, bit, fmask, outslot, sel_hint, nbytes, nrec, se);
        default: return emit_fused_t<FMT_AUX>(c, k, stride, mask, bit, fmask, outslot, sel_hint, nbytes, nrec, se);
    }
}

// filter series `bit` of source k gathered from the resident main series (k_gather_records)
static int gather_series(sofa_ctx* c, int k, int bit, uint64_t* nbytes, uint64_t* nrec) {
    Source& s = c->src[k];
    const int outslot = 1;
    uint64_t ncand = cdiv(s.n_rows, s.gather_stride);
    c->out_n[outslot] = 0;
    if (nbytes) *nbytes = 0;
    if (nrec) *nrec = 0;
    if (flip_output(c, outslot)) return SOFA_ERR_CUDA;
    if (ncand == 0) return 0;
    const uint8_t* mask = s.mask.as<uint8_t>();
    uint64_t nb = cdiv(ncand, 256);
    ENSURE(c->blockoff, (nb + 2) * 8);
    LAUNCH(c, K_FILTER, (k_mask_count<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>())));
    CK_S(launch_scan(c, c->blockoff.as<uint64_t>(), nb));
    uint64_t cnt = 0;
    CK_S(small_d2h(c, &cnt, c->blockoff.as<uint64_t>() + nb, 8));
    CK_S(stream_sync(c));
    if (cnt == 0) return 0;
    ENSURE(c->rowlen, cnt * 4 + 16);
    LAUNCH(c, K_FILTER, (k_mask_write<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>(), c->rowlen.as<uint32_t>())));
    uint64_t nblocks = cdiv(cnt, EMIT_TPB);
    ENSURE(c->tiles, (nblocks + 2) * 8);
    double avg = s.gather_total / (double)(ncand ? ncand : 1);
    uint64_t cap = (uint64_t)(cnt * (avg * 1.15 + 16.0)) + (1 << 16);
    const uint32_t slot_bytes = 48u * 1024u;
    for (int attempt = 0; attempt < 6; attempt++, cap *= 2) {
        uint64_t gen0 = c->arena_gen;
        char* outp = nullptr;
        { int ra = out_alloc(c, outslot, cap, &outp); if (ra) return ra; }
        if (c->arena_on && c->arena_gen != gen0) return 1;           // the arena wrapped: the main series is gone, format instead
        if (reset_err(c)) return SOFA_ERR_CUDA;
        CK(cudaMemsetAsync(c->tiles.p, 0, (nblocks + 1) * 8, c->stream));
        GatherArgs a;
        a.main_out = s.gather_ptr; a.rec_off = s.rec_off.as<unsigned long long>();
        a.index = c->rowlen.as<uint32_t>(); a.nsel = cnt; a.out = outp; a.cap = cap;
        a.state = c->tiles.as<unsigned long long>(); a.ticket = c->d_flags + 4;
        a.total = c->d_err + 11; a.overflow = c->d_flags + 2; a.slot_bytes = slot_bytes;
        LAUNCH(c, K_GATHER, (k_gather_records<<<(unsigned)nblocks, EMIT_TPB, slot_bytes + 32, c->stream>>>(a)));
        unsigned long long total = 0; unsigned int ovf = 0;
        CK_S(small_d2h(c, &total, c->d_err + 11, 8));
        CK_S(small_d2h(c, &ovf, c->d_flags + 2, 4));
        int r = check_kernel_error(c, "emit_json (gather)");
        if (r) return r;
        if (ovf) { if (total + 64 > cap * 2) cap = total / 2 + 64; continue; }
        out_commit(c, outslot, total);
        if (nbytes) *nbytes = total;
        if (nrec) *nrec = cnt;
        return 0;
    }
    FAIL(SOFA_ERR_CAP, "emit_json (gather): output capacity estimate failed repeatedly");
}

int set_gather_attr(void) {
    return (int)cudaFuncSetAttribute(k_gather_records, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

extern "C" {

int sofa_b200_emit_json_series(sofa_ctx* c, int k, uint32_t stride, int32_t nkw, const char* const* kws, uint64_t* counts,
                               uint64_t* nbytes, uint64_t* nrec) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0 || nkw < 0 || nkw > 8 || (nkw && !kws)) FAIL(SOFA_ERR_ARG, "emit_json_series: stride >= 1, 0..8 keywords");
    Source& s = c->src[k];
    if (k == SRC_PERF && !(s.viz_ready && s.viz_stride == stride))
        FAIL(SOFA_ERR_STATE, "emit_json_series(perf): call viz_mangled + set_demangled for this stride first");
    s.gather_ok = false;
    for (int i = 0; i < nkw; i++) {
        size_t l = strlen(kws[i]);
        if (l == 0 || l > 32) FAIL(SOFA_ERR_ARG, "emit_json_series: keyword length must be 1..32 bytes");
    }
    if (nkw == 0) return sofa_b200_emit_json(c, k, stride, -1, nbytes, nrec);
    // masks + hit counts of every keyword in one k_filter pass over the names (sofa_b200_set_filters), then the main
    // series with the byte offset of every record kept
    r = sofa_b200_set_filters(c, k, stride, nkw, kws, counts); if (r) return r;
    uint64_t ncand = cdiv(s.n_rows, stride);
    ENSURE(s.rec_off, (ncand + 2) * 8);
    uint32_t fmask = 0;
    r = fmask_for(c, k, stride, &fmask); if (r) return r;
    JsonFuse jf{s.rec_off.as<unsigned long long>(), true};
    uint64_t nb = 0, nr = 0;
    r = emit_fused_t<FMT_JSON>(c, k, stride, nullptr, 0, fmask, 1, ncand, &nb, &nr, nullptr, &jf);
    if (r) return r;
    s.gather_ok = nr > 0; s.gather_stride = stride; s.gather_total = (double)nb;
    if (nbytes) *nbytes = nb >= 2 ? nb - 2 : 0;
    if (nrec) *nrec = nr;
    return 0;
}

int sofa_b200_emit_csv(sofa_ctx* c, int k, uint32_t stride, uint64_t* nbytes) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "emit_csv: stride must be >= 1");
    uint32_t fmask = 0;
    r = fmask_for(c, k, stride, &fmask); if (r) return r;
    return emit_fused_t<FMT_CSV>(c, k, stride, nullptr, 0, fmask, 0, cdiv(c->src[k].n_rows, stride), nbytes, nullptr);
}

int sofa_b200_emit_comm_csv(sofa_ctx* c, uint64_t* nbytes) {
    int r = need_parsed(c, SRC_GPU); if (r) return r;
    uint32_t fmask = 0;
    r = fmask_for(c, SRC_GPU, 1, &fmask); if (r) return r;
    return emit_fused_t<FMT_COMM>(c, SRC_GPU, 1, nullptr, 0, fmask, 0, c->src[SRC_GPU].n_rows, nbytes, nullptr);
}

int sofa_b200_emit_json(sofa_ctx* c, int k, uint32_t stride, int32_t filter_idx, uint64_t* nbytes, uint64_t* nrec) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "emit_json: stride must be >= 1");
    Source& s = c->src[k];
    if (k == SRC_PERF && !(s.viz_ready && s.viz_stride == stride))
        FAIL(SOFA_ERR_STATE, "emit_json(perf): call viz_mangled + set_demangled for this stride first");
    const uint8_t* mask = nullptr;
    if (filter_idx >= 0) {
        if (filter_idx >= s.nkw || s.mask_stride != stride) FAIL(SOFA_ERR_STATE, "emit_json: filter not set for this stride");
        mask = s.mask.as<uint8_t>();
    }
    if (filter_idx < 0) s.gather_ok = false;              // a plain main-series emit replaces the resident one
    if (filter_idx >= 0 && s.gather_ok && s.gather_stride == stride &&
        (s.gather_gen == ~0ull || (c->arena_on && s.gather_gen == c->arena_gen))) {
        uint64_t gb = 0, gr = 0;
        r = gather_series(c, k, filter_idx, &gb, &gr);
        if (r == 0) {
            if (nbytes) *nbytes = gb >= 2 ? gb - 2 : 0;
            if (nrec) *nrec = gr;
            return 0;
        }
        if (r < 0) return r;
        s.gather_ok = false;                                 // r == 1: main series no longer resident
    }
    uint32_t fmask = 0;
    r = fmask_for(c, k, stride, &fmask); if (r) return r;
    uint64_t nb = 0, nr = 0;
    uint64_t hint = filter_idx >= 0 ? s.filter_counts[filter_idx] : cdiv(s.n_rows, stride);
    r = emit_fused_t<FMT_JSON>(c, k, stride, mask, filter_idx >= 0 ? filter_idx : 0, fmask, 1, hint, &nb, &nr);
    if (r) return r;
    if (nbytes) *nbytes = nb >= 2 ? nb - 2 : 0;       // the leading ", " is dropped by fetch
    if (nrec) *nrec = nr;
    return 0;
}

}  // extern "C"
