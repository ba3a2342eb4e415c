// libsofa_b200.so, translation unit: keyword filters and the viz-name dictionary (demangling round trip).
#define SOFA_K_VIZ
#include "ctx.cuh"

extern "C" {

int sofa_b200_set_filters(sofa_ctx* c, int k, uint32_t stride, int32_t nkw, const char* const* kws, uint64_t* counts) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0 || nkw < 0 || nkw > 8 || (nkw && !kws)) FAIL(SOFA_ERR_ARG, "set_filters: 0..8 keywords");
    Source& s = c->src[k];
    if (k == SRC_PERF && !(s.viz_ready && s.viz_stride == stride))
        FAIL(SOFA_ERR_STATE, "set_filters(perf): call viz_mangled + set_demangled for this stride first");
    FilterTable ft;
    memset(&ft, 0, sizeof ft);
    size_t lens[8] = {0};
    for (int i = 0; i < nkw; i++) {
        lens[i] = strlen(kws[i]);
        if (lens[i] == 0 || lens[i] > 32) FAIL(SOFA_ERR_ARG, "set_filters: keyword length must be 1..32 bytes");
    }
    filter_table_build(&ft, nkw, kws, lens);
    uint64_t ncand = cdiv(s.n_rows, stride);
    ENSURE(c->ftable, sizeof(FilterTable));
    ENSURE(s.mask, ncand + 16);
    if (reset_err(c)) return SOFA_ERR_CUDA;
    CK_S(small_h2d(c, c->ftable.p, &ft, sizeof ft));
    CK_S(stream_sync(c));
    if (ncand) {
        SrcView v = make_view(c, k, true);
        unsigned g = (unsigned)cdiv(ncand, 256);
        const FilterTable* dft = c->ftable.as<FilterTable>();
        if (k == SRC_PERF) LAUNCH(c, K_FILTER, (k_filter<SRC_PERF><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
        else if (k == SRC_GPU) LAUNCH(c, K_FILTER, (k_filter<SRC_GPU><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
        else LAUNCH(c, K_FILTER, (k_filter<SRC_NET><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
    }
    unsigned long long cn[8] = {0};
    CK_S(small_d2h(c, cn, c->d_err + 2, 64));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    if (counts) for (int i = 0; i < nkw; i++) counts[i] = cn[i];
    for (int i = 0; i < 8; i++) s.filter_counts[i] = i < nkw ? cn[i] : 0;
    s.mask_stride = stride; s.nkw = nkw;
    return 0;
}

int sofa_b200_viz_mangled(sofa_ctx* c, uint32_t stride, uint64_t* nbytes, uint32_t* n_names) {
    int r = need_parsed(c, SRC_PERF); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "viz_mangled: stride must be >= 1");
    Source& s = c->src[SRC_PERF];
    uint64_t ncand = cdiv(s.n_rows, stride);
    s.viz_ready = false; s.viz_stride = stride;
    c->out_n[2] = 0;
    if (nbytes) *nbytes = 0;
    if (n_names) *n_names = 0;
    ENSURE(s.vizhash, (ncand + 1) * 8);
    ENSURE(s.viz_id, (ncand + 1) * 4);
    if (ncand == 0) return 0;
    SrcView v = make_view(c, SRC_PERF, false);
    for (uint64_t cap = 1ull << 16;; cap <<= 2) {
        if (cap > (1ull << 30)) FAIL(SOFA_ERR_CUDA, "viz dictionary overflow");
        ENSURE(c->keys, cap * 8); ENSURE(c->rows, cap * 8);
        if (reset_err(c)) return SOFA_ERR_CUDA;
        CK(cudaMemsetAsync(c->keys.p, 0, cap * 8, c->stream));
        CK(cudaMemsetAsync(c->rows.p, 0xff, cap * 8, c->stream));
        LAUNCH(c, K_VIZ, (k_viz_hash<<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(
            v, ncand, stride, s.vizhash.as<uint64_t>(), c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(),
            cap - 1, c->d_flags + 2)));
        unsigned int ovf = 0;
        CK_S(small_d2h(c, &ovf, c->d_flags + 2, 4));
        CK_S(stream_sync(c));
        CK(cudaGetLastError());
        if (!ovf) {
            LAUNCH(c, K_VIZ, (k_dict_len<<<(unsigned)cdiv(cap, 256), 256, 0, c->stream>>>(
                v, c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(), cap, c->d_err + 1, c->d_flags + 3)));
            unsigned long long total = 0; unsigned int n = 0;
            CK_S(small_d2h(c, &total, c->d_err + 1, 8));
            CK_S(small_d2h(c, &n, c->d_flags + 3, 4));
            CK_S(stream_sync(c));
            // occupancy above 50% makes probing slow and overflow likely next time: grow
            if ((uint64_t)n * 2 > cap) continue;
            ENSURE(c->out[2], total + 64);
            CK(cudaMemsetAsync(c->d_err + 1, 0, 8, c->stream));
            if (total)
                LAUNCH(c, K_VIZ, (k_dict_write<<<(unsigned)cdiv(cap, 256), 256, 0, c->stream>>>(
                    v, c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(), cap, c->d_err + 1, c->out[2].as<char>())));
            CK_S(stream_sync(c));
            CK(cudaGetLastError());
            c->out_n[2] = total;
            if (nbytes) *nbytes = total;
            if (n_names) *n_names = n;
            return 0;
        }
    }
}

int sofa_b200_set_demangled(sofa_ctx* c, uint32_t n, const char* mangled, const uint64_t* moff,
                            const char* demangled, const uint64_t* doff) {
    int r = need_parsed(c, SRC_PERF); if (r) return r;
    Source& s = c->src[SRC_PERF];
    if (s.viz_stride == 0) FAIL(SOFA_ERR_STATE, "set_demangled: call viz_mangled first");
    if (n && (!mangled || !moff || !demangled || !doff)) FAIL(SOFA_ERR_ARG, "set_demangled: null table");
    std::vector<std::pair<uint64_t, uint32_t>> order(n);
    for (uint32_t i = 0; i < n; i++) {
        HashSink h;
        for (uint64_t j = moff[i]; j < moff[i + 1]; j++) h.put(mangled[j]);
        order[i] = {h.h ? h.h : 1, i};
    }
    std::sort(order.begin(), order.end());
    std::vector<uint64_t> hashes(n + 1), offs(n + 2);
    std::string blob;
    for (uint32_t i = 0; i < n; i++) {
        if (i && order[i].first == order[i - 1].first) FAIL(SOFA_ERR_ARG, "set_demangled: duplicate or colliding names");
        hashes[i] = order[i].first;
        offs[i] = blob.size();
        uint32_t src = order[i].second;
        blob.append(demangled + doff[src], demangled + doff[src + 1]);
    }
    offs[n] = blob.size();
    ENSURE(s.dict_hash, (n + 1) * 8); ENSURE(s.dict_off, (n + 2) * 8); ENSURE(s.dict_blob, blob.size() + 16);
    if (n) {
        CK_S(small_h2d(c, s.dict_hash.p, hashes.data(), n * 8));
        CK_S(small_h2d(c, s.dict_blob.p, blob.data(), blob.size()));
    }
    CK_S(small_h2d(c, s.dict_off.p, offs.data(), (n + 1) * 8));
    CK_S(stream_sync(c));
    s.dict_n = (int)n;
    uint64_t ncand = cdiv(s.n_rows, s.viz_stride);
    if (ncand) {
        if (reset_err(c)) return SOFA_ERR_CUDA;
        DictView d{(int)n, s.dict_hash.as<uint64_t>(), s.dict_off.as<uint64_t>(), s.dict_blob.as<char>()};
        LAUNCH(c, K_VIZ, (k_viz_resolve<<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(
            s.vizhash.as<uint64_t>(), ncand, d, s.viz_id.as<int32_t>(), c->d_err)));
        r = check_kernel_error(c, "set_demangled (a mangled viz name has no demangled entry)");
        if (r) return r;
    }
    s.viz_ready = true;
    return 0;
}

}  // extern "C"
