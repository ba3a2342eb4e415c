/* libsofa_b200.so - C ABI of the B200-native SOFA trace preprocess/analyze hot path.
 *
 * The reference (cyliustack/sofa) has no FFI: its hot path is the pair of Python entry
 * points sofa_preprocess(cfg) (bin/sofa_preprocess.py:377) and sofa_analyze(cfg)
 * (bin/sofa_analyze.py:793).  This library is what those entry points bind through
 * ctypes; each entry below names the reference code it replaces.
 *
 * Conventions: every function returns 0 on success or a negative SOFA_ERR_* code; the
 * message is available from sofa_b200_last_error().  No exception crosses the boundary.
 * The caller owns host buffers, the library owns device buffers behind an opaque
 * sofa_ctx (one per GPU; not thread-safe per ctx, safe across ctxs).  All calls are
 * synchronous on the ctx's stream.  Plain pointers and sizes only - no torch types.
 */
#ifndef SOFA_B200_H
#define SOFA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOFA_B200_ABI_VERSION 2

#define SOFA_OK 0
#define SOFA_ERR_CUDA (-1)   /* CUDA runtime error */
#define SOFA_ERR_ARG (-2)    /* bad argument */
#define SOFA_ERR_ROW (-3)    /* a record the reference would raise on; see sofa_b200_error_row */
#define SOFA_ERR_STATE (-4)  /* call order (e.g. emit before parse) */
#define SOFA_ERR_CAP (-5)    /* caller buffer too small */

enum { SOFA_SRC_PERF = 0, SOFA_SRC_GPU = 1, SOFA_SRC_NET = 2,
       /* 13-column trace CSV bodies read by concurrency_breakdown (upload + parse like a source) */
       SOFA_SRC_MPSTAT_CSV = 3, SOFA_SRC_NVSMI_CSV = 4, SOFA_SRC_NETSTAT_CSV = 5,
       /* recorder side files as written by `sofa record` (upload, then sofa_b200_side_preprocess) */
       SOFA_RAW_MPSTAT = 6, SOFA_RAW_NVSMI_QUERY = 7, SOFA_RAW_NETSTAT = 8, SOFA_RAW_VMSTAT = 9, SOFA_RAW_NVSMI_DMON = 10 };

typedef struct sofa_ctx sofa_ctx;

/* Scalars that parameterise one source's ingest (computed once per file by the host,
 * reference sofa_preprocess.py:418-436, 1385-1419, 1545-1551, 1766-1796). */
typedef struct {
    double time_base;            /* cfg.time_base */
    double t_offset;             /* perf: unix-uptime; gpu: t_glb_gpu_base - t_base; net: 0 */
    int32_t absolute_timestamp;  /* cfg.absolute_timestamp */
    int32_t perf_has_cycles;     /* cfg.perf_events.find('cycles') != -1 */
    int32_t n_mhz;               /* len(cpu_mhz_xp), >= 1 */
    int32_t n_cudaproc;          /* pid column of gpu rows */
    const double* mhz_xp;        /* host pointers, n_mhz entries each */
    const double* mhz_fp;
    double ts_rescale, dt_rescale, payload_unit;
    int32_t idx[9];              /* Start, Duration, Size, Throughput, Context, Stream, Name, Src Ctx, Dst Ctx; -1 = absent */
    int32_t api_mode;            /* 1: gpu source holds `nvprof --print-api-trace` rows (cuda_api_trace_read);
                                  * 2 (side CSV sources 3..5): vmstat.csv name fields;
                                  * 3 (trace sources 0..2): the text is <src>trace.csv itself, header line skipped by data_off -
                                  *   the rows pd.read_csv would give (stand-alone `sofa analyze`, sofa_analyze.py:898-977) */
} sofa_ingest_params_t;

/* gpu_profile (sofa_analyze.py:343-377) + comm_profile (sofa_common.py:23-177) results */
typedef struct {
    double gpu_time, kernel_time, nccl_time;
    int64_t num_gpus;            /* number of distinct deviceId values */
    int64_t n_gpus_max;          /* max(deviceId) (comm_profile n_gpus); 0 -> no comm report */
    double payload[4];           /* H2D, D2H, D2D, P2P summed payload */
    double bandwidth[4];         /* mean bandwidth where payload > 64000 */
    double duration[4];          /* summed duration */
    /* (1+n) x (1+n) row-major matrices, n = min(n_gpus_max, 15): payload sum, count, bw sum, bw count */
    double accum[256], accum_count[256], accum_bw[256], accum_bw_count[256];
    /* the "[id]: time" lines of --verbose (:349-354): the first 64 deviceIds in ascending order and their summed durations */
    int64_t per_gpu_id[64]; double per_gpu_time[64];
} sofa_gpu_profile_t;

typedef struct {
    uint64_t launches[32];       /* kernel launches per kernel kind */
    double ms[32];               /* CUDA-event time per kernel kind (profiling on) */
    uint64_t h2d_bytes, d2h_bytes;
} sofa_stats_t;

int sofa_b200_abi_version(void);
const char* sofa_b200_last_error(void);
/* row index and row error code of the last SOFA_ERR_ROW */
int sofa_b200_error_row(sofa_ctx*, uint64_t* row, int32_t* code);

int sofa_b200_create(sofa_ctx** out, int device);
int sofa_b200_destroy(sofa_ctx*);

/* pinned host memory for callers that want full-speed H2D/D2H */
void* sofa_b200_host_alloc(uint64_t nbytes);
void sofa_b200_host_free(void*);

/* Copy one source's raw text to the device; records start at data_off (nvprof: after the
 * three header rows).  Replaces f.readlines() of perf.script / gputrace.tmp / net.tmp
 * (sofa_preprocess.py:1194, 1383, 1788). */
int sofa_b200_upload(sofa_ctx*, int src, const void* text, uint64_t nbytes, uint64_t data_off);

/* Line split + per-record parse into SoA columns on the device.  Replaces
 * Pool.map(cpu_trace_read | gpu_trace_read | net_trace_read) (:110-341, 1196-1202,
 * 1420-1432, 1791-1799). */
int sofa_b200_parse(sofa_ctx*, int src, const sofa_ingest_params_t*, uint64_t* n_rows);

/* Column dtype flags over rows i % stride == 0 (pandas dtype inference):
 * bit0 some Throughput parsed as float (gpu), bit1 some [] packet row (net),
 * bit2 some valid packet row (net). */
int sofa_b200_row_flags(sofa_ctx*, int src, uint32_t stride, uint32_t* flags);

/* <src>trace.csv body (no header line) of rows i % stride == 0, formatted on the device
 * into output buffer 0.  Replaces DataFrame(...).to_csv(float_format='%.6f')
 * (:1204-1211, 1433-1444, 1800-1807). */
int sofa_b200_emit_csv(sofa_ctx*, int src, uint32_t stride, uint64_t* nbytes);

/* comm.csv body (no header): index, timestamp, pkt_src, pkt_dst, payload, bandwidth of every gpu
 * row with pandas' default float text (repr of the gputrace.csv value), into output buffer 0.
 * Replaces df_gpu.to_csv(columns=[...]) (sofa_common.py:160-167). */
int sofa_b200_emit_comm_csv(sofa_ctx*, uint64_t* nbytes);

/* Keyword filters, Series.str.contains(literal) (:1446-1452 gpu, full frame;
 * :1819-1827 cpu, down-sampled demangled viz frame).  counts[k] = matching rows. */
int sofa_b200_set_filters(sofa_ctx*, int src, uint32_t stride, int32_t nkw, const char* const* keywords,
                          uint64_t* counts);

/* Unique mangled (_Z...) viz names of perf rows i % stride == 0, '\n'-separated, into
 * output buffer 2 (host demangles them once each: cxxfilt.demangle, :1811-1816). */
int sofa_b200_viz_mangled(sofa_ctx*, uint32_t stride, uint64_t* nbytes, uint32_t* n_names);
int sofa_b200_set_demangled(sofa_ctx*, uint32_t n, const char* mangled, const uint64_t* moff,
                            const char* demangled, const uint64_t* doff);

/* report.js records ("{...}, {...}") of one series into output buffer 1: rows
 * i % stride == 0, restricted to filter `filter_idx` when >= 0.  Replaces
 * json.loads(df.to_json(orient='records')) + json.dump (:343-374). */
int sofa_b200_emit_json(sofa_ctx*, int src, uint32_t stride, int32_t filter_idx, uint64_t* nbytes,
                        uint64_t* n_records);

/* Main series of a source plus its keyword filters: sofa_b200_set_filters (counts[k] = matching rows, masks kept
 * for sofa_b200_emit_json(filter_idx)), then the records of the main series go to output buffer 1 and the byte
 * offset of every record is kept with that output, so that the following sofa_b200_emit_json(src, stride, k) calls
 * GATHER series k from it instead of formatting the rows again (a filter series holds the very same record bytes,
 * sofa_preprocess.py:1446-1452, 1819-1827).  Valid when the filters run on the frame of the main series: always
 * for the cpu filters, for the gpu filters when plot_ratio == 1. */
int sofa_b200_emit_json_series(sofa_ctx*, int src, uint32_t stride, int32_t nkw, const char* const* keywords,
                               uint64_t* counts, uint64_t* nbytes, uint64_t* n_records);

/* D2H of an output buffer (0 csv, 1 json, 2 names) */
int sofa_b200_fetch(sofa_ctx*, int which, void* host_out, uint64_t cap);

/* Overlapped transfers (pinned host memory): uploads run on an H2D copy stream (parse waits
 * for its source's upload), fetches of output buffers 0/1 run on a D2H copy stream while the
 * next emit fills the other half of a double buffer.  `ticket` identifies the fetch for
 * sofa_b200_wait_fetch; sofa_b200_sync drains every stream of the ctx. */
/* Optional device arena for csv/json outputs: every emit bump-allocates its buffer from it,
 * so any number of asynchronous fetches can be in flight; sofa_b200_sync rewinds it. */
int sofa_b200_set_output_arena(sofa_ctx*, uint64_t nbytes);
int sofa_b200_upload_async(sofa_ctx*, int src, const void* text, uint64_t nbytes, uint64_t data_off);
int sofa_b200_fetch_async(sofa_ctx*, int which, void* host_out, uint64_t cap, int32_t* ticket);
int sofa_b200_wait_fetch(sofa_ctx*, int32_t ticket);
int sofa_b200_sync(sofa_ctx*);

/* Raw device column to host (parity tests). perf: 0 ts,1 event,2 dur (f64), 9 pid, 10 tid (i32);
 * gpu: 0 ts,2 dur,6 bw (f64), 5 payload (i64), 3 dev,10 tid,7 src,8 dst (i32), 4 kind (u8);
 * net: 0 ts (f64), 5 payload,7 src,8 dst (i64), 4 proto (u8). */
int sofa_b200_get_column(sofa_ctx*, int src, int col, void* host_out, uint64_t cap_bytes);

/* gpu_profile + comm_profile features on the parsed gpu rows (sofa_analyze.py:343-377,
 * sofa_common.py:23-177).  want_matrices: also fill the (src,dst) matrices that the
 * reference prints only with --verbose (python-loop sums, row order). */
int sofa_b200_gpu_profile(sofa_ctx*, sofa_gpu_profile_t* out, int want_matrices);

/* get_top_k_events (sofa_analyze.py:34-44): names by summed duration, descending.
 * names: '\0'-terminated, in order, into output buffer 2; sums[k] their totals. */
int sofa_b200_top_k(sofa_ctx*, int src, uint32_t k, double* sums, uint64_t* nbytes, uint32_t* n_out);

/* cpu_profile (sofa_analyze.py:694-710): total duration and the k largest-duration rows
 * (descending duration, ties by row index): row index, timestamp and duration as read from
 * cputrace.csv; their names '\0'-terminated into output buffer 2. */
int sofa_b200_cpu_profile(sofa_ctx*, uint32_t k, double* total, uint64_t* rows, double* ts, double* dur,
                          uint64_t* names_nbytes, uint32_t* n_out);

/* Per-window aggregates of concurrency_breakdown (sofa_analyze.py:75-243) over the parsed side
 * tables: window k = (ends[k] - delta, ends[k]].  mpstat: row count, max usr/sys/idl/iow, sums of
 * usr/sys/irq, max/min of usr+sys+irq (started at 0 / 100 as the reference does); nvsmi SM rows
 * (event 0): count, max/min/sum duration, distinct deviceIds; netstat: tx (event 0) / rx (event 1)
 * bandwidth sums (numpy pairwise, row order).  num_gpus = distinct deviceIds of the whole nvsmi table. */
typedef struct {
    uint32_t mp_count; int32_t usr_max, sys_max, idl_max, iow_max, cpu_max, cpu_min;
    int64_t usr_sum, sys_sum, irq_sum;
    uint32_t nv_count, nv_ndev; double nv_max, nv_min; int64_t nv_sum;
    uint32_t tx_count, rx_count; double tx_sum, rx_sum;
} sofa_window_t;
int sofa_b200_window_reduce(sofa_ctx*, const double* ends, uint32_t n_windows, double delta, sofa_window_t* out,
                            uint32_t* num_gpus);

/* net_profile (sofa_analyze.py:385-493) over the packets of nettrace.csv = rows i % stride == 0 of the parsed
 * net source: net_time (summed duration of the "network:tcp:" names, group by group in sorted-name order) and
 * the payload summed per (pkt_src, pkt_dst) pair (unordered; the host ranks them into netrank.csv). */
typedef struct { int64_t src, dst, payload_sum, count; } sofa_net_pair_t;
int sofa_b200_net_profile(sofa_ctx*, uint32_t stride, double* net_time, int64_t* payload_total, sofa_net_pair_t* pairs,
                          uint32_t cap, uint32_t* n_pairs);
/* net_profile --verbose (:398): df.groupby(['pkt_src','pkt_dst','payload']).size() over the same packets, unordered. */
typedef struct { int64_t src, dst, payload, count; } sofa_net_triple_t;
int sofa_b200_net_triples(sofa_ctx*, uint32_t stride, sofa_net_triple_t* out, uint32_t cap, uint32_t* n_out);
/* --network_filters (sofa_preprocess.py:1214-1231) on the packets i % stride == 0: mask bit k (k < n_others) = payload > 0
 * and pkt_src == base and pkt_dst == others[k] (series "to_<others[k]>"), bit 4 + k = payload > 0 and pkt_dst == base and
 * pkt_src == others[k] ("from_<others[k]>"); counts[8] = rows per bit.  The series are then emitted with
 * sofa_b200_emit_json(SRC_NET, stride, bit).  At most 4 peers per call. */
int sofa_b200_set_net_filters(sofa_ctx*, uint32_t stride, int64_t base, int32_t n_others, const int64_t* others, uint64_t* counts);
/* mpstat_profile (sofa_analyze.py:735-790) over the parsed mpstat.csv body (source 3): n_cores = int(max(deviceId) + 1),
 * sums[core * 5 + q] = sum over the core's rows, in row order with numpy's pairwise tree, of t_USR, t_SYS, t_IDL, t_IOW,
 * t_IRQ (q = 0..4) with dt_all = 0.1 where IDL == 100 else duration / ((100 - IDL) / 100) and t_X = dt_all * X / 100. */
int sofa_b200_mpstat_profile(sofa_ctx*, double* sums, uint32_t cap_cores, uint32_t* n_cores);
/* Column statistics of a parsed side table (source 3..5) over the rows with event == `event` (all rows when use_event
 * is 0): row count, numpy-pairwise sum in row order (mean = sum / count) and, for q = 0.25, 0.5, 0.75, the two order
 * statistics numpy's linear quantile interpolates between.  value_col: 0 timestamp, 2 duration, 3 deviceId, 6 bandwidth.
 * Used by netbandwidth_profile (sofa_analyze.py:531-594). */
int sofa_b200_side_quantiles(sofa_ctx*, int slot, int value_col, int use_event, double event, double* sum, uint64_t* count,
                             double* q_lo, double* q_hi);
/* Sums of the five integers parsed out of the name column of a side table (mpstat.csv: usr sys idl iow irq; vmstat.csv
 * parsed with api_mode = 2: bi bo in cs, vmstat_profile :712-733) and the row count. */
int sofa_b200_side_int_sums(sofa_ctx*, int slot, int64_t* sums, uint64_t* n_rows);
/* AISI per-iteration profile (sofa_aisi.py:21-59 iter_profile over the slices of :429-453): the gpu rows of device
 * `device_id` (df_gpu[df_gpu.deviceId == 1] in the reference) with times[i] <= timestamp < times[i+1] form iteration i
 * (n_times - 1 iterations; the boundaries come from the host-side iteration detection).  Per iteration: row count,
 * timestamps of its first / last row (frame order), numpy-pairwise sums of duration over all its rows and over the
 * rows whose name contains "_fw", "_bw", "gemm", "CUDA_COPY", "nccl", payload sum, number of distinct streams (tid). */
typedef struct {
    uint64_t n_rows;
    double first_ts, last_ts;
    double gpu_time, fw_time, bw_time, gemm_time, copy_name_time, allreduce_time;
    int64_t payload;
    uint32_t streams, pad;
} sofa_iter_t;
int sofa_b200_iter_profile(sofa_ctx*, int32_t device_id, const double* times, uint32_t n_times, sofa_iter_t* out);
/* --spotlight_gpu region of interest (sofa_analyze.py:872-895): the saturating-counter state machine over the
 * event == 0, deviceId == 0 rows of the parsed nvsmi_trace.csv body (source 4), in row order.  roi_begin / roi_end are
 * in/out (cfg.roi_begin / cfg.roi_end); both are reset to 0 when end < begin. */
int sofa_b200_nvsmi_roi(sofa_ctx*, double* roi_begin, double* roi_end);
/* nvsmi_profile (sofa_analyze.py:259-341) over the parsed nvsmi_trace.csv body (source 4), optionally
 * restricted to roi_begin < timestamp <= roi_end (--spotlight_gpu).  One entry per (deviceId, event)
 * present: row count, duration sum (means = sum / count), and for q = 0.25, 0.5, 0.75 the two order
 * statistics numpy's linear quantile interpolates between (q_lo at floor((count-1)q), q_hi at ceil).
 * dev_order: device ids in order of first appearance (the reference iterates a Python set of them);
 * ts_first / ts_last: timestamps of the first / last row used (df.iloc[0] / df.iloc[-1]). */
typedef struct {
    int32_t dev, event;
    uint64_t count; int64_t sum;
    double q_lo[3], q_hi[3];
} sofa_nvsmi_group_t;
int sofa_b200_nvsmi_profile(sofa_ctx*, int use_roi, double roi_begin, double roi_end, sofa_nvsmi_group_t* groups,
                            uint32_t cap, uint32_t* n_groups, int32_t* dev_order, uint32_t* n_dev, double* ts_first,
                            double* ts_last, uint64_t* n_used);
/* Recorder side files -> rows of mpstat.csv + usr_sys.csv (sofa_preprocess.py:482-546), vmstat.csv
 * (:787-1008), nvsmi_trace.csv (nvsmi_query.txt :1013-1084; nvsmi.txt enc/dec rows :1092-1183) and
 * netstat.csv + netbandwidth.csv (:1235-1337), and the records of their report.js series (mpstat_usr;
 * vmstat_*; nvsmi_sm / nvsmi_mem / nvsmi_enc / nvsmi_dec; tx_bandwidth / rx_bandwidth).
 * Upload the file with sofa_b200_upload (nvsmi_query.txt: data_off = after its first line, which
 * the reference skips), then call sofa_b200_side_preprocess once and the emitters as needed.
 *   n_lines : lines parsed;  n_cand : candidate rows per list (mpstat: rows of mpstat[1:], some of
 *   which are dropped or go to usr_sys.csv; nvsmi_query: one sm + one mem row per line; netstat: one tx +
 *   one rx row per line after the first; vmstat: eight rows per accepted line; nvsmi.txt: one enc + one
 *   dec row per accepted line, those of the first seconds dropped);  no_data : nvidia-smi reported a
 *   failure (:1017-1021, :1095-1099).
 * list: 0 = sm / tx / enc (and mpstat), 1 = mem / rx / dec; vmstat 0..7 = bi, bo, in, cs, wa, st, usr, sys.  CSV and aux bodies go to output buffer 0, JSON
 * records to output buffer 1 (fetch as usual).  aux = usr_sys.csv / netbandwidth.csv body. */
typedef struct {
    double time_base;            /* cfg.time_base */
    int32_t absolute_timestamp;  /* cfg.absolute_timestamp */
    int32_t reserved;
    int64_t local_utc_offset;    /* seconds the host's local zone is ahead of UTC (naive datetime.timestamp()) */
    double nvsmi_tz_add;         /* cfg.nvsmi_time_zone * 3600 */
} sofa_side_params_t;
int sofa_b200_side_preprocess(sofa_ctx*, int raw_src, const sofa_side_params_t*, uint64_t* n_lines, uint64_t* n_cand,
                              int32_t* no_data);
int sofa_b200_side_emit_csv(sofa_ctx*, int raw_src, int list, uint64_t* nbytes, uint64_t* n_rows);
int sofa_b200_side_emit_json(sofa_ctx*, int raw_src, int list, uint64_t* nbytes, uint64_t* n_records);
int sofa_b200_side_emit_aux(sofa_ctx*, int raw_src, uint64_t* nbytes, uint64_t* n_rows);

/* Recorder side files, second group: diskstat.txt (table 0; sofa_preprocess.py:548-673), blktrace.txt (1; :688-781),
 * strace.txt (2; :1618-1704), pystacks.txt (3; :1706-1761) -> diskstat.csv / diskstat_vector.csv /
 * diskstat_vector_ui.csv, blktrace.csv, strace.csv, pystacks.csv and the records of the diskstat,
 * blktrace_starting_block, blktrace_time, strace and pystacks series of report.js.  One table at a time:
 * sofa_b200_side2_preprocess (upload + parse + index), for diskstat sofa_b200_side2_dev_table / _set_devs, then
 * sofa_b200_side2_emit per output into output buffer 0 (CSV) or 1 (report.js records); read with sofa_b200_fetch.
 *   category / category_text: the `cpuid` an earlier section of sofa_preprocess left behind (the reference's strace and
 *   pystacks rows carry it in their last column; a string when blktrace.txt set it);
 *   info.total = total_strace_duration; info.last_span = byte range of the cpu token of the last accepted blktrace
 *   line (what `cpuid` becomes), a == b when there is none; info.flags bit 0 = "blktrace maybe record failed!". */
typedef struct {
    double time_base, strace_min_time;
    int64_t pid, category;
    const char* category_text;
    int32_t absolute_timestamp, plot_ratio, category_is_str, category_len;
} sofa_side2_params_t;
typedef struct {
    uint64_t n_lines, n_entries, n_kept, n_devs;
    double total;
    uint32_t last_span_a, last_span_b;
    int32_t flags, pad;
} sofa_side2_info_t;
#define SOFA_SIDE2_DEV_CAP 1024
#define SOFA_SIDE2_CSV0 0
#define SOFA_SIDE2_CSV1 1
#define SOFA_SIDE2_JSON0 2
#define SOFA_SIDE2_JSON1 3
#define SOFA_SIDE2_VEC 4
#define SOFA_SIDE2_VEC_UI 5
int sofa_b200_side2_preprocess(sofa_ctx*, int table, const void* text, uint64_t nbytes, const sofa_side2_params_t* p,
                               sofa_side2_info_t* info);
/* diskstat: span[2 * slot], span[2 * slot + 1] = byte range of the device name held by table slot `slot` (a = 0xffffffff:
 * empty slot); then per slot the sector size the host read from /sys/block/<dev>/queue/hw_sector_size (512 when absent,
 * :563-570) and whether --diskstat_filters selects the device.  set_devs also finds the devices whose traffic never
 * moved (dropped from every output, :650-653) and the dtype of the r_time / w_time / await_time columns. */
int sofa_b200_side2_dev_table(sofa_ctx*, uint32_t* span, uint32_t cap_slots);
int sofa_b200_side2_set_devs(sofa_ctx*, const int32_t* secsize, const uint8_t* ui_keep, uint32_t n_slots);
int sofa_b200_side2_emit(sofa_ctx*, int table, int what, uint64_t* nbytes, uint64_t* n_rows);
/* diskstat_profile (sofa_analyze.py:640-700) over the text of diskstat_vector.csv: one entry per device with the two
 * order statistics pandas' groupby().quantile interpolates between, for q = 0.25, 0.5, 0.75 of d_read, d_write and
 * d_disk_total (q_lo / q_hi[3 * column + q]), and groupby().mean() (Kahan sum in row order / count) of d_read, d_write,
 * d_disk_total, r_iops, w_iops, iops, await_time.  span: as sofa_b200_side2_dev_table (the names, by slot).
 * blktrace_latency (sofa_analyze.py:627-631) over the text of blktrace.csv: the rows whose event is 'C', the order
 * statistics of their duration that numpy's linear quantile interpolates between. */
typedef struct {
    uint32_t slot, pad;
    uint64_t count;
    double q_lo[9], q_hi[9];
    double mean[7];
} sofa_disk_group_t;
int sofa_b200_diskstat_profile(sofa_ctx*, const void* csv_text, uint64_t nbytes, sofa_disk_group_t* groups, uint32_t cap,
                               uint32_t* n_groups, uint32_t* span);
int sofa_b200_blktrace_latency(sofa_ctx*, const void* csv_text, uint64_t nbytes, uint64_t* count, double* q_lo, double* q_hi);
/* CUDA-event stopwatch on the ctx stream: mark(0) ... mark(1); elapsed = ms between them */
int sofa_b200_mark(sofa_ctx*, int which);
int sofa_b200_elapsed_ms(sofa_ctx*, double* ms);

/* Verification helpers (bench.py checks windows of the > 4 GiB outputs against the CPU oracle): a byte range of
 * output buffer 0 / 1 as it is on the device (buffer 1 still carries the leading ", " of the first record), and
 * the inclusive byte offset at which block `block` (128 consecutive selected rows) of the LAST emit ends, read
 * from that emit's look-back state. */
int sofa_b200_fetch_range(sofa_ctx*, int which, uint64_t off, uint64_t nbytes, void* host_out);

/* Streaming I/O of the drop-in entry points (sofa_preprocess reads and writes files through one pinned staging buffer):
 * a source text uploaded in pieces (the first piece sizes the device buffer, the last one makes the source ready for
 * sofa_b200_parse), and outputs parked on the device until report.js is written - the reference's series order
 * (sofa_preprocess.py:1849-2102) is not the order the series are produced in.  hold_output copies output `which`
 * (report.js records without their leading ", ") device to device and returns a handle; hold_fetch reads a piece of
 * it; hold_release(handle) frees it, hold_release(-1) all of them. */
int sofa_b200_upload_chunk(sofa_ctx*, int src, uint64_t off, const void* piece, uint64_t nbytes, uint64_t total, uint64_t data_off);
int sofa_b200_hold_output(sofa_ctx*, int which, int32_t* handle);
int sofa_b200_hold_fetch(sofa_ctx*, int32_t handle, uint64_t off, uint64_t nbytes, void* host_out);
int sofa_b200_hold_release(sofa_ctx*, int32_t handle);
int sofa_b200_emit_block_end(sofa_ctx*, uint64_t block, uint64_t* end_bytes);

/* One logical trace sharded over several contexts / GPUs (newline-aligned byte ranges, one per rank): a column's dtype
 * (int or float text, pandas' per-file inference) is a property of the WHOLE file, so after sofa_b200_parse every rank
 * all-reduces (bitwise OR, NCCL) the value of sofa_b200_row_flags and hands the result back with this call; all later
 * emits of the shard format with it. */
int sofa_b200_set_global_flags(sofa_ctx*, int src, uint32_t stride, uint32_t flags);

/* Gathering the parsed shards of one source on one rank (sofa_analyze runs there): device pointers and sizes of a
 * shard's text and columns (export), fresh buffers for the whole source (reserve), and the switch to them once every
 * shard's bytes are in place (commit; the caller moves the bytes: device copies and NCCL send/recv between ranks).
 * cols[] order - perf: timestamp event duration pid tid line_start pieces; gpu: timestamp duration bandwidth payload
 * deviceId tid pkt_src pkt_dst copyKind bw_is_float line_start name_off name_len; net: timestamp payload pkt_src
 * pkt_dst proto. */
typedef struct {
    uint64_t n_rows, text_bytes, data_off;
    void* text;
    void* cols[16];
    uint32_t col_elem[16];
    int32_t n_cols, pad;
} sofa_shard_t;
int sofa_b200_shard_export(sofa_ctx*, int src, sofa_shard_t* out);
int sofa_b200_shard_reserve(sofa_ctx*, int src, uint64_t total_rows, uint64_t total_text, sofa_shard_t* out);
int sofa_b200_shard_commit(sofa_ctx*, int src, int32_t nshards, const uint64_t* shard_rows, const uint64_t* shard_text);

/* Test hooks.  SOFA_DEBUG_NAME_HASH_MASK: the 64-bit name hashes that key get_top_k_events / nccl_time groups
 * (sofa_analyze.py:34-44, 363-368) are ANDed with `value`; a narrow mask forces collisions, which the group
 * verification pass must then report (SOFA_ERR_ROW) instead of merging two names. */
#define SOFA_DEBUG_NAME_HASH_MASK 1
int sofa_b200_set_debug(sofa_ctx*, int key, uint64_t value);

int sofa_b200_set_profiling(sofa_ctx*, int on);
int sofa_b200_get_stats(sofa_ctx*, sofa_stats_t* out);
int sofa_b200_reset_stats(sofa_ctx*);
const char* sofa_b200_kernel_name(int kind);

/* Synthetic trace text generated on the device (bench / test infrastructure; same spec as
 * sofa_b200/synth.py).  variant: perf 0/1 = without/with [cpu] column; gpu 0 = s/ms/MB,
 * 1 = ms/us/KB + P2P columns; net = percent of non-IP lines. */
int sofa_b200_synth(sofa_ctx*, int src, uint64_t seed, uint64_t n_rows, int variant, uint64_t* nbytes);
/* rows [row0, row0 + n_rows) of the same trace (one shard of it); the nvprof header rows only when with_header */
int sofa_b200_synth_range(sofa_ctx*, int src, uint64_t seed, uint64_t row0, uint64_t n_rows, int variant, int with_header,
                          uint64_t* nbytes);
int sofa_b200_text_fetch(sofa_ctx*, int src, void* host_out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
