"""CudaEngine: the object sofa_preprocess / sofa_analyze drive.  A thin veneer over the
C ABI: host buffers in, numpy byte arrays out.  No record-level work happens in Python."""
import ctypes
import os

import numpy as np

from . import _lib as L
from .hostlogic import ROW_ERRORS


class PinnedBuffer:
    """cudaHostAlloc'd memory exposed as a numpy uint8 array."""
    def __init__(self, lib, nbytes):
        self.lib = lib
        self.ptr = lib.sofa_b200_host_alloc(ctypes.c_uint64(max(1, nbytes)))
        if not self.ptr:
            raise MemoryError("cudaHostAlloc(%d) failed" % nbytes)
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * max(1, nbytes)).from_address(self.ptr))[:nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.sofa_b200_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


class CudaEngine:
    def __init__(self, device=0):
        self.lib = L.load_lib()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.sofa_b200_create(ctypes.byref(self.ctx), int(device))
        if rc != 0:
            self.ctx = None
            raise L.SofaB200Error(rc, self.lib.sofa_b200_last_error().decode())
        self._keep = []

    def close(self):
        if self.ctx:
            if getattr(self, '_stage_buf', None) is not None:
                self._stage_buf = None
                self.lib.sofa_b200_host_free(ctypes.c_void_p(self._stage_ptr))
            self.lib.sofa_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == 0:
            return
        msg = self.lib.sofa_b200_last_error().decode()
        if rc == L.SOFA_ERR_ROW:
            row = ctypes.c_uint64()
            code = ctypes.c_int32()
            self.lib.sofa_b200_error_row(self.ctx, ctypes.byref(row), ctypes.byref(code))
            msg += " (%s)" % ROW_ERRORS.get(code.value, "?")
            raise L.SofaB200Error(rc, msg, row.value, code.value)
        raise L.SofaB200Error(rc, msg)

    # ---- ingest
    @staticmethod
    def _ptr(buf):
        if isinstance(buf, np.ndarray):
            return buf.ctypes.data, buf.nbytes, buf
        if isinstance(buf, (bytes, bytearray)):
            a = np.frombuffer(buf, dtype=np.uint8)
            return a.ctypes.data, a.nbytes, a
        a = np.frombuffer(memoryview(buf), dtype=np.uint8)
        return a.ctypes.data, a.nbytes, a

    def upload(self, kind, text, data_off=0):
        ptr, n, keep = self._ptr(text)
        self._check(self.lib.sofa_b200_upload(self.ctx, kind, ctypes.c_void_p(ptr), n, data_off))

    def upload_async(self, kind, text, data_off=0):
        """H2D on the copy stream (text must stay alive and should be pinned); parse() waits for it."""
        ptr, n, keep = self._ptr(text)
        self._check(self.lib.sofa_b200_upload_async(self.ctx, kind, ctypes.c_void_p(ptr), n, data_off))

    def fetch_async(self, which, n, out):
        """D2H of output buffer `which` (0 csv, 1 json) on the copy stream -> ticket."""
        if out.nbytes < n:
            raise ValueError("output buffer too small")
        t = ctypes.c_int32()
        self._check(self.lib.sofa_b200_fetch_async(self.ctx, which, ctypes.c_void_p(out.ctypes.data), out.nbytes,
                                                   ctypes.byref(t)))
        return t.value

    def wait_fetch(self, ticket):
        self._check(self.lib.sofa_b200_wait_fetch(self.ctx, ticket))

    def set_output_arena(self, nbytes):
        self._check(self.lib.sofa_b200_set_output_arena(self.ctx, int(nbytes)))

    def sync(self):
        self._check(self.lib.sofa_b200_sync(self.ctx))

    def make_params(self, p):
        ip = L.IngestParams()
        ip.time_base, ip.t_offset = p['time_base'], p['t_offset']
        ip.absolute_timestamp, ip.perf_has_cycles = p['absolute_timestamp'], p['perf_has_cycles']
        xp = np.ascontiguousarray(p['mhz_xp'], dtype=np.float64)
        fp = np.ascontiguousarray(p['mhz_fp'], dtype=np.float64)
        self._keep = [xp, fp]
        ip.n_mhz, ip.mhz_xp, ip.mhz_fp = len(xp), xp.ctypes.data, fp.ctypes.data
        ip.n_cudaproc = p['n_cudaproc']
        ip.ts_rescale, ip.dt_rescale, ip.payload_unit = p['ts_rescale'], p['dt_rescale'], p['payload_unit']
        for i in range(9):
            ip.idx[i] = p['idx'][i]
        ip.api_mode = p.get('api_mode', 0)
        return ip

    def parse(self, kind, params):
        ip = params if isinstance(params, L.IngestParams) else self.make_params(params)
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_parse(self.ctx, kind, ctypes.byref(ip), ctypes.byref(n)))
        return n.value

    def load(self, kind, text, params, data_off=0):
        self.upload(kind, text, data_off)
        n = self.parse(kind, params)
        return n, self.row_flags(kind, 1)

    def row_flags(self, kind, stride):
        f = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_row_flags(self.ctx, kind, stride, ctypes.byref(f)))
        return f.value

    # ---- sharded traces: the parsed state of a source as raw device memory
    def shard_export(self, kind):
        d = L.ShardDesc()
        self._check(self.lib.sofa_b200_shard_export(self.ctx, kind, ctypes.byref(d)))
        return d

    def shard_reserve(self, kind, total_rows, total_text):
        d = L.ShardDesc()
        self._check(self.lib.sofa_b200_shard_reserve(self.ctx, kind, int(total_rows), int(total_text), ctypes.byref(d)))
        return d

    def shard_commit(self, kind, rows, text_bytes):
        n = len(rows)
        a = (ctypes.c_uint64 * n)(*rows)
        b = (ctypes.c_uint64 * n)(*text_bytes)
        self._check(self.lib.sofa_b200_shard_commit(self.ctx, kind, n, a, b))

    def set_global_flags(self, kind, stride, flags):
        """sharded traces: the OR over all shards of row_flags (column dtypes are a property of the whole file)"""
        self._check(self.lib.sofa_b200_set_global_flags(self.ctx, kind, stride, int(flags)))

    # ---- streaming I/O of the entry points: one pinned staging buffer between the files and the device
    STAGE_BYTES = 64 << 20

    def _stage(self):
        if getattr(self, '_stage_buf', None) is None:
            self.lib.sofa_b200_host_alloc.restype = ctypes.c_void_p
            p = self.lib.sofa_b200_host_alloc(ctypes.c_uint64(self.STAGE_BYTES))
            if not p:
                raise MemoryError("cudaHostAlloc failed (staging buffer)")
            self._stage_ptr = p
            self._stage_buf = np.ctypeslib.as_array((ctypes.c_uint8 * self.STAGE_BYTES).from_address(p))
        return self._stage_ptr, self._stage_buf

    def load_file(self, kind, path, params, data_off=0):
        """load() of a whole file, read through the pinned staging buffer (readinto + H2D per piece)"""
        ptr, buf = self._stage()
        total = os.path.getsize(path)
        with open(path, 'rb', buffering=0) as f:
            off = 0
            while True:
                m = f.readinto(memoryview(buf)[:min(self.STAGE_BYTES, total - off)]) if off < total else 0
                self._check(self.lib.sofa_b200_upload_chunk(self.ctx, kind, off, ctypes.c_void_p(ptr), m or 0, total, data_off))
                off += m or 0
                if off >= total:
                    break
                if not m:
                    raise IOError("%s shrank while it was read" % path)
        n = self.parse(kind, params)
        return n, self.row_flags(kind, 1)

    def write_output(self, which, n, f):
        """bytes [0, n) of output buffer `which` (report.js records: after their leading ', ') -> file object"""
        ptr, buf = self._stage()
        skip = 2 if which == 1 else 0
        for off in range(0, n, self.STAGE_BYTES):
            m = min(self.STAGE_BYTES, n - off)
            self._check(self.lib.sofa_b200_fetch_range(self.ctx, which, off + skip, m, ctypes.c_void_p(ptr)))
            f.write(memoryview(buf)[:m])

    def hold_output(self, which):
        """park the current output on the device -> handle (see HeldOutput)"""
        h = ctypes.c_int32()
        self._check(self.lib.sofa_b200_hold_output(self.ctx, which, ctypes.byref(h)))
        return h.value

    def hold_write(self, handle, n, f):
        ptr, buf = self._stage()
        for off in range(0, n, self.STAGE_BYTES):
            m = min(self.STAGE_BYTES, n - off)
            self._check(self.lib.sofa_b200_hold_fetch(self.ctx, handle, off, m, ctypes.c_void_p(ptr)))
            f.write(memoryview(buf)[:m])

    def hold_release(self, handle=-1):
        self._check(self.lib.sofa_b200_hold_release(self.ctx, handle))

    # ---- outputs
    def _fetch(self, which, n, out=None):
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        elif out.nbytes < n:
            raise ValueError("output buffer too small")
        self._check(self.lib.sofa_b200_fetch(self.ctx, which, ctypes.c_void_p(out.ctypes.data), out.nbytes))
        return out[:n]

    def csv_device(self, kind, stride):
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_emit_csv(self.ctx, kind, stride, ctypes.byref(n)))
        return n.value

    def csv(self, kind, stride, out=None):
        return self._fetch(0, self.csv_device(kind, stride), out)

    def comm_csv(self, out=None):
        """comm.csv body (sofa_common.py:160-167) of the parsed gpu rows."""
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_emit_comm_csv(self.ctx, ctypes.byref(n)))
        return self._fetch(0, n.value, out)

    def json_device(self, kind, stride, filt):
        n = ctypes.c_uint64()
        nrec = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_emit_json(self.ctx, kind, stride, filt, ctypes.byref(n), ctypes.byref(nrec)))
        return n.value, nrec.value

    def json(self, kind, stride, filt, out=None):
        n, nrec = self.json_device(kind, stride, filt)
        return self._fetch(1, n, out), nrec

    def json_series_device(self, kind, stride, kws):
        """Main series with the keyword filters matched in the same pass -> (bytes, records, [hits per keyword]);
        the filter series are then gathered by json_device(kind, stride, k)."""
        arr = (ctypes.c_char_p * max(1, len(kws)))(*[k.encode() if isinstance(k, str) else k for k in kws])
        counts = (ctypes.c_uint64 * 8)()
        n, nrec = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_emit_json_series(self.ctx, kind, stride, len(kws), arr, counts,
                                                        ctypes.byref(n), ctypes.byref(nrec)))
        return n.value, nrec.value, [counts[i] for i in range(len(kws))]

    def json_series(self, kind, stride, kws, out=None):
        n, nrec, counts = self.json_series_device(kind, stride, kws)
        return self._fetch(1, n, out), nrec, counts

    def fetch_range(self, which, off, n):
        out = np.empty(n, dtype=np.uint8)
        self._check(self.lib.sofa_b200_fetch_range(self.ctx, which, int(off), int(n), ctypes.c_void_p(out.ctypes.data)))
        return out

    def emit_block_end(self, block):
        v = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_emit_block_end(self.ctx, int(block), ctypes.byref(v)))
        return v.value

    def set_filters(self, kind, stride, kws):
        arr = (ctypes.c_char_p * max(1, len(kws)))(*[k.encode() if isinstance(k, str) else k for k in kws])
        counts = (ctypes.c_uint64 * 8)()
        self._check(self.lib.sofa_b200_set_filters(self.ctx, kind, stride, len(kws), arr, counts))
        return [counts[i] for i in range(len(kws))]

    def viz_mangled(self, stride):
        n = ctypes.c_uint64()
        cnt = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_viz_mangled(self.ctx, stride, ctypes.byref(n), ctypes.byref(cnt)))
        if n.value == 0:
            return []
        raw = self._fetch(2, n.value).tobytes()
        names = raw.decode().split('\n')[:-1]
        assert len(names) == cnt.value
        return names

    def set_demangled(self, pairs):
        mb = b"".join(m.encode() for m, _ in pairs)
        db = b"".join(d.encode() for _, d in pairs)
        moff = (ctypes.c_uint64 * (len(pairs) + 1))()
        doff = (ctypes.c_uint64 * (len(pairs) + 1))()
        for i, (m, d) in enumerate(pairs):
            moff[i + 1] = moff[i] + len(m.encode())
            doff[i + 1] = doff[i] + len(d.encode())
        self._check(self.lib.sofa_b200_set_demangled(self.ctx, len(pairs), mb, moff, db, doff))

    _COLTYPES = {0: {0: 'f8', 1: 'f8', 2: 'f8', 9: 'i4', 10: 'i4'},
                 1: {0: 'f8', 2: 'f8', 6: 'f8', 5: 'i8', 3: 'i4', 10: 'i4', 7: 'i4', 8: 'i4', 4: 'u1'},
                 2: {0: 'f8', 5: 'i8', 7: 'i8', 8: 'i8', 4: 'u1'}}

    def column(self, kind, col, n_rows):
        dt = np.dtype(self._COLTYPES[kind][col])
        out = np.empty(n_rows, dtype=dt)
        self._check(self.lib.sofa_b200_get_column(self.ctx, kind, col, ctypes.c_void_p(out.ctypes.data), out.nbytes))
        return out

    # ---- analyze
    def gpu_profile(self, want_matrices=False):
        gp = L.GpuProfile()
        self._check(self.lib.sofa_b200_gpu_profile(self.ctx, ctypes.byref(gp), 1 if want_matrices else 0))
        return gp

    def top_k(self, kind, k):
        """[(name, summed duration)] by descending sum (get_top_k_events)."""
        sums = (ctypes.c_double * max(1, k))()
        n = ctypes.c_uint64()
        cnt = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_top_k(self.ctx, kind, k, sums, ctypes.byref(n), ctypes.byref(cnt)))
        names = self._fetch(2, n.value).tobytes().decode().split('\x00')[:-1] if n.value else []
        return list(zip(names, [sums[i] for i in range(cnt.value)]))

    def cpu_profile(self, k):
        """(total duration, [(row, timestamp, duration, name)] by descending duration)."""
        total = ctypes.c_double()
        rows = (ctypes.c_uint64 * max(1, k))()
        ts = (ctypes.c_double * max(1, k))()
        dur = (ctypes.c_double * max(1, k))()
        nb = ctypes.c_uint64()
        cnt = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_cpu_profile(self.ctx, k, ctypes.byref(total), rows, ts, dur,
                                                   ctypes.byref(nb), ctypes.byref(cnt)))
        names = self._fetch(2, nb.value).tobytes().decode().split('\x00')[:-1] if nb.value else []
        return total.value, [(rows[i], ts[i], dur[i], names[i]) for i in range(cnt.value)]

    def window_reduce(self, ends, delta):
        """Per-window aggregates of the parsed side tables (sources 3,4,5) -> (list of Window, num_gpus)."""
        ends = np.ascontiguousarray(ends, dtype=np.float64)
        out = (L.Window * len(ends))()
        ng = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_window_reduce(self.ctx, ends.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                     len(ends), float(delta), out, ctypes.byref(ng)))
        return out, ng.value

    def net_profile(self, stride):
        """(net_time, total payload, {(pkt_src, pkt_dst): summed payload}) of the packets i % stride == 0."""
        t, total, n = ctypes.c_double(), ctypes.c_int64(), ctypes.c_uint32()
        cap = getattr(self, '_net_pair_cap', 1 << 14)
        while True:
            if getattr(self, '_net_pairs', None) is None or len(self._net_pairs) < cap:
                self._net_pairs = (L.NetPair * cap)()
            rc = self.lib.sofa_b200_net_profile(self.ctx, int(stride), ctypes.byref(t), ctypes.byref(total), self._net_pairs, cap,
                                                ctypes.byref(n))
            if rc == L.SOFA_ERR_CAP and cap < (1 << 24):
                cap *= 8                                        # more address pairs than the buffer holds: grow and repeat
                continue
            self._check(rc)
            break
        self._net_pair_cap = cap
        pairs = self._net_pairs
        return t.value, total.value, {(pairs[i].src, pairs[i].dst): pairs[i].payload_sum for i in range(n.value)}

    def net_triples(self, stride):
        """{(pkt_src, pkt_dst, payload): packets} of the packets i % stride == 0 (net_profile --verbose)."""
        cap = 1 << 16
        n = ctypes.c_uint32()
        while True:
            buf = (L.NetTriple * cap)()
            rc = self.lib.sofa_b200_net_triples(self.ctx, int(stride), buf, cap, ctypes.byref(n))
            if rc == L.SOFA_ERR_CAP and n.value > cap:
                cap = int(n.value) + 16
                continue
            self._check(rc)
            break
        return {(buf[i].src, buf[i].dst, buf[i].payload): buf[i].count for i in range(n.value)}

    def set_net_filters(self, stride, base, others):
        """--network_filters masks for up to 4 peers: bit k = to_<others[k]>, bit 4 + k = from_<others[k]> -> hit counts [8]"""
        arr = (ctypes.c_int64 * 4)(*(list(others) + [0] * (4 - len(others))))
        counts = (ctypes.c_uint64 * 8)()
        self._check(self.lib.sofa_b200_set_net_filters(self.ctx, int(stride), int(base), len(others), arr, counts))
        return [counts[i] for i in range(8)]

    def mpstat_profile(self):
        """[[t_USR, t_SYS, t_IDL, t_IOW, t_IRQ] per core] summed over the parsed mpstat.csv (source 3)"""
        cap = 4096
        sums = (ctypes.c_double * (cap * 5))()
        n = ctypes.c_uint32()
        self._check(self.lib.sofa_b200_mpstat_profile(self.ctx, sums, cap, ctypes.byref(n)))
        return [[sums[c * 5 + q] for q in range(5)] for c in range(n.value)]

    def side_quantiles(self, slot, value_col, event=None):
        """(count, pairwise sum, [(lo, hi)] for q = .25, .5, .75) of one column of a parsed side table"""
        s, n = ctypes.c_double(), ctypes.c_uint64()
        lo, hi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
        self._check(self.lib.sofa_b200_side_quantiles(self.ctx, slot, value_col, 0 if event is None else 1,
                                                      0.0 if event is None else float(event), ctypes.byref(s), ctypes.byref(n), lo, hi))
        return n.value, s.value, [(lo[i], hi[i]) for i in range(3)]

    def side_int_sums(self, slot):
        sums = (ctypes.c_int64 * 5)()
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_side_int_sums(self.ctx, slot, sums, ctypes.byref(n)))
        return [sums[i] for i in range(5)], n.value

    def nvsmi_roi(self, roi_begin, roi_end):
        a, b = ctypes.c_double(float(roi_begin)), ctypes.c_double(float(roi_end))
        self._check(self.lib.sofa_b200_nvsmi_roi(self.ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def iter_profile(self, times, device_id=1):
        """per-iteration sums of the parsed gpu rows of one device between consecutive boundaries -> list of L.IterStats"""
        t = np.ascontiguousarray(times, dtype=np.float64)
        out = (L.IterStats * max(1, len(t) - 1))()
        self._check(self.lib.sofa_b200_iter_profile(self.ctx, int(device_id), t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                    len(t), out))
        return [out[i] for i in range(len(t) - 1)]

    def nvsmi_profile(self, roi=None):
        """Per (deviceId, event) group of the parsed nvsmi_trace.csv (source 4): {(dev, event): (count, sum,
        [(lo, hi)] for q = .25, .5, .75)}, device ids in order of first appearance, first / last timestamp."""
        groups = (L.NvsmiGroup * 256)()
        ng, nd = ctypes.c_uint32(), ctypes.c_uint32()
        order = (ctypes.c_int32 * 64)()
        t0, t1, used = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_nvsmi_profile(self.ctx, 1 if roi else 0, float(roi[0]) if roi else 0.0,
                                                     float(roi[1]) if roi else 0.0, groups, 256, ctypes.byref(ng), order,
                                                     ctypes.byref(nd), ctypes.byref(t0), ctypes.byref(t1), ctypes.byref(used)))
        out = {}
        for i in range(ng.value):
            g = groups[i]
            out[(g.dev, g.event)] = (int(g.count), int(g.sum), [(g.q_lo[q], g.q_hi[q]) for q in range(3)])
        return out, [order[i] for i in range(nd.value)], t0.value, t1.value, used.value

    # ---- recorder side files (mpstat.txt 6, nvsmi_query.txt 7, netstat.txt 8)
    def side_preprocess(self, raw_src, text, time_base, absolute_timestamp, local_utc_offset=0, nvsmi_tz_add=0.0,
                        data_off=0):
        """-> (lines parsed, candidate rows per list, no_data flag); text=None: already uploaded"""
        if text is not None:
            self.upload(raw_src, text, data_off)
        sp = L.SideParams()
        sp.time_base, sp.absolute_timestamp = float(time_base), 1 if absolute_timestamp else 0
        sp.local_utc_offset, sp.nvsmi_tz_add = int(local_utc_offset), float(nvsmi_tz_add)
        nl, nc, nd = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32()
        self._check(self.lib.sofa_b200_side_preprocess(self.ctx, raw_src, ctypes.byref(sp), ctypes.byref(nl),
                                                       ctypes.byref(nc), ctypes.byref(nd)))
        return nl.value, nc.value, bool(nd.value)

    def side_csv(self, raw_src, lst=0):
        """(CSV body of one list, number of rows)"""
        n, nr = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_side_emit_csv(self.ctx, raw_src, lst, ctypes.byref(n), ctypes.byref(nr)))
        return self._fetch(0, n.value), nr.value

    def side_json(self, raw_src, lst=0):
        n, nr = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_side_emit_json(self.ctx, raw_src, lst, ctypes.byref(n), ctypes.byref(nr)))
        return self._fetch(1, n.value), nr.value

    def side_aux(self, raw_src):
        """usr_sys.csv / netbandwidth.csv body (default float repr) and its row count"""
        n, nr = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_side_emit_aux(self.ctx, raw_src, ctypes.byref(n), ctypes.byref(nr)))
        return self._fetch(0, n.value), nr.value

    # ---- recorder side files, second group: diskstat.txt 0, blktrace.txt 1, strace.txt 2, pystacks.txt 3
    def side2_preprocess(self, table, text, time_base, absolute_timestamp, plot_ratio=1, pid=0, strace_min_time=0.0, category=None):
        """upload + parse + index one file; category: the `cpuid` an earlier section left behind (int, or bytes when
        blktrace.txt set it).  -> dict(n_lines, n_entries, n_kept, n_devs, total, last_span, flags)"""
        p = L.Side2Params()
        p.time_base, p.strace_min_time, p.pid = float(time_base), float(strace_min_time), int(pid)
        p.absolute_timestamp, p.plot_ratio = 1 if absolute_timestamp else 0, int(plot_ratio)
        if isinstance(category, (bytes, str)):
            cb = category.encode() if isinstance(category, str) else bytes(category)
            p.category_text, p.category_is_str, p.category_len, p.category = cb, 1, len(cb), 0
        else:
            p.category_text, p.category_is_str, p.category_len, p.category = None, 0, 0, int(category or 0)
        buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(1, dtype=np.uint8)
        info = L.Side2Info()
        self._check(self.lib.sofa_b200_side2_preprocess(self.ctx, table, ctypes.c_void_p(buf.ctypes.data), ctypes.c_uint64(len(text)),
                                                        ctypes.byref(p), ctypes.byref(info)))
        self._s2_text = text
        return dict(n_lines=info.n_lines, n_entries=info.n_entries, n_kept=info.n_kept, n_devs=info.n_devs, total=info.total,
                    last_span=(info.last_span_a, info.last_span_b), flags=info.flags)

    def side2_dev_names(self):
        """[(table slot, device name)] of the diskstat.txt just preprocessed"""
        span = np.zeros(2 * L.SIDE2_DEV_CAP, dtype=np.uint32)
        self._check(self.lib.sofa_b200_side2_dev_table(self.ctx, ctypes.c_void_p(span.ctypes.data), L.SIDE2_DEV_CAP))
        text = bytes(self._s2_text)
        return [(i, text[int(span[2 * i]):int(span[2 * i + 1])]) for i in range(L.SIDE2_DEV_CAP) if span[2 * i] != 0xffffffff]

    def side2_set_devs(self, devs):
        """devs: {slot: (sector size, selected by --diskstat_filters)}"""
        sec = np.full(L.SIDE2_DEV_CAP, 512, dtype=np.int32)
        ui = np.zeros(L.SIDE2_DEV_CAP, dtype=np.uint8)
        for slot, (sz, keep) in devs.items():
            sec[slot], ui[slot] = int(sz), 1 if keep else 0
        self._check(self.lib.sofa_b200_side2_set_devs(self.ctx, ctypes.c_void_p(sec.ctypes.data), ctypes.c_void_p(ui.ctypes.data),
                                                      L.SIDE2_DEV_CAP))

    def side2_emit(self, table, what):
        """what: 0 / 1 CSV rows of list 0 / 1, 2 / 3 report.js records of list 0 / 1, 4 diskstat_vector.csv, 5 diskstat_vector_ui.csv
        -> (bytes, rows)"""
        n, nr = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.sofa_b200_side2_emit(self.ctx, table, what, ctypes.byref(n), ctypes.byref(nr)))
        return self._fetch(1 if what in (2, 3) else 0, n.value), nr.value

    def diskstat_profile(self, csv_text):
        """groups of diskstat_vector.csv by device: [(name, count, q_lo[9], q_hi[9], mean[7])] (see sofa_b200_diskstat_profile)"""
        buf = np.frombuffer(csv_text, dtype=np.uint8) if len(csv_text) else np.zeros(1, dtype=np.uint8)
        groups = (L.DiskGroup * 512)()
        ng = ctypes.c_uint32()
        span = np.zeros(2 * L.SIDE2_DEV_CAP, dtype=np.uint32)
        self._check(self.lib.sofa_b200_diskstat_profile(self.ctx, ctypes.c_void_p(buf.ctypes.data), ctypes.c_uint64(len(csv_text)), groups, 512,
                                                        ctypes.byref(ng), ctypes.c_void_p(span.ctypes.data)))
        text = bytes(csv_text)
        return [(text[int(span[2 * g.slot]):int(span[2 * g.slot + 1])].decode(), g.count, list(g.q_lo), list(g.q_hi), list(g.mean))
                for g in groups[:ng.value]]

    def blktrace_latency(self, csv_text):
        """rows of blktrace.csv whose event is 'C': (count, q_lo[3], q_hi[3]) of their duration"""
        buf = np.frombuffer(csv_text, dtype=np.uint8) if len(csv_text) else np.zeros(1, dtype=np.uint8)
        n = ctypes.c_uint64()
        lo, hi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
        self._check(self.lib.sofa_b200_blktrace_latency(self.ctx, ctypes.c_void_p(buf.ctypes.data), ctypes.c_uint64(len(csv_text)),
                                                        ctypes.byref(n), lo, hi))
        return n.value, list(lo), list(hi)

    # ---- stopwatch (CUDA events on the engine's stream)
    def mark(self, which):
        self._check(self.lib.sofa_b200_mark(self.ctx, which))

    def elapsed_ms(self):
        ms = ctypes.c_double()
        self._check(self.lib.sofa_b200_elapsed_ms(self.ctx, ctypes.byref(ms)))
        return ms.value

    # ---- stats / synth
    def set_profiling(self, on):
        self._check(self.lib.sofa_b200_set_profiling(self.ctx, 1 if on else 0))

    def reset_stats(self):
        self._check(self.lib.sofa_b200_reset_stats(self.ctx))

    def stats(self):
        st = L.Stats()
        self._check(self.lib.sofa_b200_get_stats(self.ctx, ctypes.byref(st)))
        out = {}
        for i in range(32):
            nm = self.lib.sofa_b200_kernel_name(i).decode()
            if nm and st.launches[i]:
                out[nm] = (int(st.launches[i]), float(st.ms[i]))
        return out, int(st.h2d_bytes), int(st.d2h_bytes)

    def synth(self, kind, seed, n_rows, variant=0):
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_synth(self.ctx, kind, seed, n_rows, variant, ctypes.byref(n)))
        return n.value

    def synth_range(self, kind, seed, row0, n_rows, variant=0, with_header=True):
        n = ctypes.c_uint64()
        self._check(self.lib.sofa_b200_synth_range(self.ctx, kind, seed, int(row0), int(n_rows), variant, 1 if with_header else 0,
                                                   ctypes.byref(n)))
        return n.value

    def text_fetch(self, kind, nbytes, out=None):
        if out is None:
            out = np.empty(nbytes, dtype=np.uint8)
        self._check(self.lib.sofa_b200_text_fetch(self.ctx, kind, ctypes.c_void_p(out.ctypes.data), out.nbytes))
        return out[:nbytes]

    def pinned(self, nbytes):
        return PinnedBuffer(self.lib, nbytes)
