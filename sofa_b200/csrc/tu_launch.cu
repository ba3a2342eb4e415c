// libsofa_b200.so, translation unit: dispatch to the per-instantiation launchers of inst/*.cu; count_lines, row flags,
// keyword filter instantiations (small kernels) live here.
#define SOFA_K_COUNT
#include "ctx.cuh"
using namespace sofa;

void inst_parse_perf(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_attr_perf(void);
void inst_parse_gpu(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_attr_gpu(void);
void inst_parse_net(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_attr_net(void);
void inst_parse_side(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_attr_side(void);
void inst_parse_raw(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_attr_raw(void);
void inst_emit_perf_csv(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_perf_csv(void);
void inst_emit_perf_json(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_perf_json(void);
void inst_emit_gpu_csv(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_gpu_csv(void);
void inst_emit_gpu_json(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_gpu_json(void);
void inst_emit_net_csv(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_net_csv(void);
void inst_emit_net_json(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_net_json(void);
void inst_emit_gpu_comm(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_gpu_comm(void);
void inst_emit_sideout_csv(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_sideout_csv(void);
void inst_emit_sideout_json(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_sideout_json(void);
void inst_emit_sideout_aux(unsigned, size_t, cudaStream_t, const SrcView&, const FusedArgs&);
int inst_emit_attr_sideout_aux(void);
void inst_parse_tma_perf(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, uint64_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_tma_attr_perf(void);
int inst_parse_tma_occ_perf(size_t);
void inst_parse_tma_gpu(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, uint64_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_tma_attr_gpu(void);
int inst_parse_tma_occ_gpu(size_t);
void inst_parse_tma_net(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, uint64_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_tma_attr_net(void);
int inst_parse_tma_occ_net(size_t);

void inst_parse_csv_perf(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_csv_attr_perf(void);
void inst_parse_csv_gpu(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_csv_attr_gpu(void);
void inst_parse_csv_net(unsigned, size_t, cudaStream_t, const char*, uint64_t, uint64_t, uint32_t, uint32_t, const uint64_t*, const IngestCfg&, const AllCols&, unsigned long long*);
int inst_parse_csv_attr_net(void);
void launch_parse_csv(int kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                      uint32_t tile_bytes, uint32_t max_lines, const uint64_t* tile_row_base, const IngestCfg& cfg,
                      const AllCols& cols, unsigned long long* err) {
    if (kind == SRC_PERF) inst_parse_csv_perf(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err);
    else if (kind == SRC_GPU) inst_parse_csv_gpu(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err);
    else inst_parse_csv_net(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err);
}

int parse_tma_occupancy(int kind, size_t smem) {
    return kind == SRC_PERF ? inst_parse_tma_occ_perf(smem) : (kind == SRC_GPU ? inst_parse_tma_occ_gpu(smem) : inst_parse_tma_occ_net(smem));
}
void launch_parse_tma(int kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                      uint32_t tile_bytes, uint32_t max_lines, uint64_t ntiles, const uint64_t* tile_row_base, const IngestCfg& cfg,
                      const AllCols& cols, unsigned long long* err) {
    if (kind == SRC_PERF) inst_parse_tma_perf(grid, smem, st, text, begin, end, tile_bytes, max_lines, ntiles, tile_row_base, cfg, cols, err);
    else if (kind == SRC_GPU) inst_parse_tma_gpu(grid, smem, st, text, begin, end, tile_bytes, max_lines, ntiles, tile_row_base, cfg, cols, err);
    else inst_parse_tma_net(grid, smem, st, text, begin, end, tile_bytes, max_lines, ntiles, tile_row_base, cfg, cols, err);
}

void launch_parse(int tpl_kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                  uint32_t tile_bytes, uint32_t max_lines, const uint64_t* tile_row_base, const IngestCfg& cfg,
                  const AllCols& cols, unsigned long long* err) {
    switch (tpl_kind) {
        case SRC_PERF: inst_parse_perf(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err); break;
        case SRC_GPU: inst_parse_gpu(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err); break;
        case SRC_NET: inst_parse_net(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err); break;
        case SRC_SIDE: inst_parse_side(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err); break;
        default: inst_parse_raw(grid, smem, st, text, begin, end, tile_bytes, max_lines, tile_row_base, cfg, cols, err); break;
    }
}
void launch_count_lines(unsigned grid, cudaStream_t st, const char* text, uint64_t begin, uint64_t end, uint32_t tile_bytes,
                        uint64_t* tile_counts, unsigned int* cr_flag) {
    k_count_lines<<<grid, PARSE_TPB, 0, st>>>(text, begin, end, tile_bytes, tile_counts, cr_flag);
}
// quote-parity pre-pass of a trace CSV (see kernels_ingest.cuh); `counts` has room for nchunks + 2 values
int csv_mask_newlines(sofa_ctx* c, char* text, uint64_t begin, uint64_t end, uint64_t* counts, uint64_t nchunks, int (*scan)(sofa_ctx*, uint64_t*, uint64_t)) {
    if (nchunks == 0) return 0;
    unsigned g = (unsigned)((nchunks + 7) / 8);
    k_csv_quote_count<<<g, 256, 0, c->stream>>>(text, begin, end, counts);
    int r = scan(c, counts, nchunks); if (r) return r;
    k_csv_mask_newlines<<<g, 256, 0, c->stream>>>(text, begin, end, counts);
    return 0;
}

void launch_emit(int kind, int fmt, unsigned grid, size_t smem, cudaStream_t st, const SrcView& v, const FusedArgs& a) {
    if (fmt == FMT_COMM) { inst_emit_gpu_comm(grid, smem, st, v, a); return; }
    if (fmt == FMT_AUX) { inst_emit_sideout_aux(grid, smem, st, v, a); return; }
    if (fmt == FMT_CSV) {
        if (kind == SRC_SIDEOUT) inst_emit_sideout_csv(grid, smem, st, v, a);
        else if (kind == SRC_PERF) inst_emit_perf_csv(grid, smem, st, v, a);
        else if (kind == SRC_GPU) inst_emit_gpu_csv(grid, smem, st, v, a);
        else inst_emit_net_csv(grid, smem, st, v, a);
    } else {
        if (kind == SRC_SIDEOUT) inst_emit_sideout_json(grid, smem, st, v, a);
        else if (kind == SRC_PERF) inst_emit_perf_json(grid, smem, st, v, a);
        else if (kind == SRC_GPU) inst_emit_gpu_json(grid, smem, st, v, a);
        else inst_emit_net_json(grid, smem, st, v, a);
    }
}
int set_kernel_attrs(void) {
    int e;
    if ((e = inst_parse_attr_perf())) return e;
    if ((e = inst_parse_attr_gpu())) return e;
    if ((e = inst_parse_attr_net())) return e;
    if ((e = inst_parse_attr_side())) return e;
    if ((e = inst_parse_attr_raw())) return e;
    if ((e = inst_emit_attr_perf_csv())) return e;
    if ((e = inst_emit_attr_perf_json())) return e;
    if ((e = inst_emit_attr_gpu_csv())) return e;
    if ((e = inst_emit_attr_gpu_json())) return e;
    if ((e = inst_emit_attr_net_csv())) return e;
    if ((e = inst_emit_attr_net_json())) return e;
    if ((e = inst_emit_attr_gpu_comm())) return e;
    if ((e = inst_emit_attr_sideout_csv())) return e;
    if ((e = inst_emit_attr_sideout_json())) return e;
    if ((e = inst_emit_attr_sideout_aux())) return e;
    if ((e = set_gather_attr())) return e;
    if ((e = inst_parse_csv_attr_perf())) return e;
    if ((e = inst_parse_csv_attr_gpu())) return e;
    if ((e = inst_parse_csv_attr_net())) return e;
    if ((e = inst_parse_tma_attr_perf())) return e;
    if ((e = inst_parse_tma_attr_gpu())) return e;
    if ((e = inst_parse_tma_attr_net())) return e;
    return 0;
}
