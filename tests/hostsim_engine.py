"""Engine with the CudaEngine interface, backed by the CPU build of the row codecs
(tests/hostsim.cpp).  TEST INFRASTRUCTURE: lets the `not gpu` suite drive the product's
host logic (sofa_b200.sofa_preprocess) and check the device row functions' arithmetic and
text output against the reference's golden files without a GPU.  The product never
imports this module."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class HsCfg(ctypes.Structure):
    _fields_ = [("time_base", ctypes.c_double), ("t_offset", ctypes.c_double),
                ("absolute_timestamp", ctypes.c_int), ("perf_has_cycles", ctypes.c_int),
                ("n_mhz", ctypes.c_int), ("mhz_xp", ctypes.c_void_p), ("mhz_fp", ctypes.c_void_p),
                ("ts_rescale", ctypes.c_double), ("dt_rescale", ctypes.c_double),
                ("payload_unit", ctypes.c_double), ("n_cudaproc", ctypes.c_int),
                ("idx", ctypes.c_int * 9), ("api_mode", ctypes.c_int)]


class HsDict(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("hash", ctypes.c_void_p), ("off", ctypes.c_void_p),
                ("blob", ctypes.c_char_p)]


def build_hostsim():
    so = os.path.join(HERE, "_build", "libhostsim.so")
    src = os.path.join(HERE, "hostsim.cpp")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    deps = [src] + [os.path.join(HERE, "..", "sofa_b200", "csrc", f) for f in ("num.cuh", "rows.cuh", "side.cuh", "side2.cuh")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hs_process.restype = ctypes.c_long
    lib.hs_fnv.restype = ctypes.c_ulonglong
    lib.hs_side.restype = ctypes.c_long
    lib.hs_side2.restype = ctypes.c_long
    return lib


class HsSide2(ctypes.Structure):
    _fields_ = [("time_base", ctypes.c_double), ("strace_min_time", ctypes.c_double), ("pid", ctypes.c_longlong),
                ("category", ctypes.c_longlong), ("category_text", ctypes.c_char_p), ("absolute", ctypes.c_int),
                ("plot_ratio", ctypes.c_int), ("category_is_str", ctypes.c_int), ("category_len", ctypes.c_int),
                ("secsize", ctypes.c_void_p), ("ui_keep", ctypes.c_void_p), ("n_dev_given", ctypes.c_int)]


class HsSide2Info(ctypes.Structure):
    _fields_ = [("n_lines", ctypes.c_longlong), ("n_entries", ctypes.c_longlong), ("n_kept", ctypes.c_longlong),
                ("n_devs", ctypes.c_longlong), ("total", ctypes.c_double), ("last_a", ctypes.c_uint), ("last_b", ctypes.c_uint),
                ("flags", ctypes.c_int), ("pad", ctypes.c_int)]


class HostSimError(RuntimeError):
    """row errors carry `row_code` like sofa_b200._lib.SofaB200Error"""
    row_code = None


def _row_error(r):
    e = HostSimError("row %d: error %d" % (r // 1000, r % 1000))
    e.row_code = r % 1000
    return e




class HostSimEngine:
    def __init__(self):
        self.lib = build_hostsim()
        self.src = {}
        self.filters = {}
        self.dict = None
        self._keep = []

    def close(self):
        pass

    def _cfg(self, p):
        c = HsCfg()
        c.time_base, c.t_offset = p['time_base'], p['t_offset']
        c.absolute_timestamp, c.perf_has_cycles = p['absolute_timestamp'], p['perf_has_cycles']
        xp = np.array(p['mhz_xp'], dtype=np.float64)
        fp = np.array(p['mhz_fp'], dtype=np.float64)
        self._keep += [xp, fp]
        c.n_mhz, c.mhz_xp, c.mhz_fp = len(xp), xp.ctypes.data, fp.ctypes.data
        c.ts_rescale, c.dt_rescale, c.payload_unit = p['ts_rescale'], p['dt_rescale'], p['payload_unit']
        c.n_cudaproc = p['n_cudaproc']
        for i in range(9):
            c.idx[i] = p['idx'][i]
        c.api_mode = p.get('api_mode', 0)
        return c

    def _run(self, kind, mode, stride, fmask, kw=None):
        text, cfg = self.src[kind]
        out = ctypes.create_string_buffer(len(text) * 16 + 4096)
        outlen = ctypes.c_long()
        flags = ctypes.c_uint()
        nrec = ctypes.c_long()
        d = ctypes.byref(self.dict) if (self.dict is not None and kind == 0) else None
        r = self.lib.hs_process(kind, text, ctypes.c_long(len(text)), ctypes.byref(cfg), mode, stride,
                                ctypes.c_uint(fmask), kw.encode() if kw else None, d, out,
                                ctypes.byref(outlen), ctypes.byref(flags), ctypes.byref(nrec))
        if r != -1:
            raise _row_error(r)
        return np.frombuffer(out.raw[:outlen.value], dtype=np.uint8), flags.value, nrec.value

    def load(self, kind, text, params, data_off=0):
        text = bytes(text)[data_off:]
        self.src[kind] = (text, self._cfg(params))
        for key in [k for k in getattr(self, 'gflags', {}) if k[0] == kind]:
            del self.gflags[key]
        _, flags, n = self._run(kind, 0, 1, 0x1fff)
        return n, flags

    def set_global_flags(self, kind, stride, flags):
        if not hasattr(self, 'gflags'):
            self.gflags = {}
        self.gflags[(kind, stride)] = flags

    def _fmask(self, kind, stride):
        flags = getattr(self, 'gflags', {}).get((kind, stride))
        if flags is None:
            _, flags, _ = self._run(kind, 0 if kind == 2 else 1, stride, 0x1fff)
        if kind == 0:
            return 0b111
        if kind == 1:
            return 0b101 | (0b1000000 if flags & 1 else 0)
        if flags & 2:
            if not flags & 4:
                raise HostSimError("every down-sampled packet row is empty")
            return 0x17ff            # every numeric column (0..10, 12)
        return 0b1000101

    def csv(self, kind, stride):
        return self._run(kind, 0, stride, self._fmask(kind, stride))[0]

    def set_filters(self, kind, stride, kws):
        self.filters[kind] = (stride, list(kws))

    # ---- the streaming interface of CudaEngine (device-resident outputs, files through a staging buffer), so that the
    # `not gpu` suite drives the same host code path as the product
    def load_file(self, kind, path, params, data_off=0):
        with open(path, 'rb') as f:
            return self.load(kind, f.read(), params, data_off)

    def csv_device(self, kind, stride):
        self._out = getattr(self, '_out', {})
        self._out[0] = self.csv(kind, stride)
        return len(self._out[0])

    def json_device(self, kind, stride, filt):
        self._out = getattr(self, '_out', {})
        self._out[1], nrec = self.json(kind, stride, filt)
        return len(self._out[1]), nrec

    def json_series_device(self, kind, stride, kws):
        self._out = getattr(self, '_out', {})
        self._out[1], nrec, counts = self.json_series(kind, stride, kws)
        return len(self._out[1]), nrec, counts

    def write_output(self, which, n, f):
        f.write(memoryview(self._out[which])[:n])

    def hold_output(self, which):
        self._held = getattr(self, '_held', {})
        h = len(self._held)
        self._held[h] = bytes(self._out[which])
        return h

    def hold_write(self, handle, n, f):
        f.write(self._held[handle][:n])

    def hold_release(self, handle=-1):
        if handle < 0:
            self._held = {}
        else:
            self._held[handle] = None

    def json_series(self, kind, stride, kws):
        """main series + keyword filters set for the following json(kind, stride, k) calls"""
        rec, n = self.json(kind, stride, -1)
        self.set_filters(kind, stride, kws)
        return rec, n, None

    def json(self, kind, stride, filt):
        if filt < 0:
            out, _, n = self._run(kind, 1, stride, self._fmask(kind, stride))
        else:
            fstride, kws = self.filters[kind]
            assert fstride == stride
            out, _, n = self._run(kind, 2, stride, self._fmask(kind, stride), kws[filt])
        return out[2:], n

    # ---- recorder side files: CPU build of side.cuh (hs_side)
    def _side(self, raw_src, mode):
        text, absolute, time_base, local_off, tz_add = self.side[raw_src]
        out = ctypes.create_string_buffer(len(text) * 40 + 4096)
        outlen, nrec, nlines, ncand = ctypes.c_long(), ctypes.c_long(), ctypes.c_long(), ctypes.c_long()
        nodata = ctypes.c_int()
        r = self.lib.hs_side(raw_src - 6, text, ctypes.c_long(len(text)), int(absolute), ctypes.c_double(time_base),
                             ctypes.c_longlong(local_off), ctypes.c_double(tz_add), mode, out, ctypes.byref(outlen),
                             ctypes.byref(nrec), ctypes.byref(nlines), ctypes.byref(ncand), ctypes.byref(nodata))
        if r != -1:
            raise _row_error(r)
        return np.frombuffer(out.raw[:outlen.value], dtype=np.uint8), nrec.value, nlines.value, ncand.value, bool(nodata.value)

    def side_preprocess(self, raw_src, text, time_base, absolute_timestamp, local_utc_offset=0, nvsmi_tz_add=0.0,
                        data_off=0):
        if not hasattr(self, 'side'):
            self.side = {}
        self.side[raw_src] = (bytes(text)[data_off:], 1 if absolute_timestamp else 0, float(time_base),
                              int(local_utc_offset), float(nvsmi_tz_add))
        _, _, nlines, ncand, nodata = self._side(raw_src, 0)
        return nlines, ncand, nodata

    def side_csv(self, raw_src, lst=0):
        return self._side(raw_src, lst)[:2]                 # mode = 16 * fmt + list

    def side_json(self, raw_src, lst=0):
        out, n = self._side(raw_src, 16 + lst)[:2]
        return out[2:], n

    def side_aux(self, raw_src):
        return self._side(raw_src, 32)[:2]

    # ---- diskstat / blktrace / strace / pystacks: CPU build of side2.cuh (hs_side2), same interface as CudaEngine
    def _side2(self, what):
        st = self.s2
        text = st['text']
        p = HsSide2()
        p.time_base, p.strace_min_time, p.pid = st['time_base'], st['min_time'], st['pid']
        cat = st['category']
        if isinstance(cat, (bytes, str)):
            cb = cat.encode() if isinstance(cat, str) else cat
            p.category_text, p.category_is_str, p.category_len, p.category = cb, 1, len(cb), 0
        else:
            p.category_text, p.category_is_str, p.category_len, p.category = None, 0, 0, int(cat or 0)
        p.absolute, p.plot_ratio = st['absolute'], st['ratio']
        sec = np.array(st.get('secsize', []), dtype=np.int32)
        ui = np.array(st.get('ui', []), dtype=np.uint8)
        p.secsize, p.ui_keep, p.n_dev_given = sec.ctypes.data, ui.ctypes.data, len(sec)
        out = ctypes.create_string_buffer(len(text) * 24 + 4096)
        outlen, nrows = ctypes.c_long(), ctypes.c_long()
        info = HsSide2Info()
        r = self.lib.hs_side2(st['table'], text, ctypes.c_long(len(text)), ctypes.byref(p), what, out, ctypes.byref(outlen),
                              ctypes.byref(nrows), ctypes.byref(info))
        if r != -1:
            raise _row_error(r)
        return np.frombuffer(out.raw[:outlen.value], dtype=np.uint8), nrows.value, info

    def side2_preprocess(self, table, text, time_base, absolute_timestamp, plot_ratio=1, pid=0, strace_min_time=0.0, category=None):
        self.s2 = dict(table=table, text=bytes(text), time_base=float(time_base), absolute=1 if absolute_timestamp else 0,
                       ratio=int(plot_ratio), pid=int(pid), min_time=float(strace_min_time), category=category)
        info = self._side2(-1)[2]
        return dict(n_lines=info.n_lines, n_entries=info.n_entries, n_kept=info.n_kept, n_devs=info.n_devs, total=info.total,
                    last_span=(info.last_a, info.last_b), flags=info.flags)

    def side2_dev_names(self):
        """device names of diskstat.txt in table-slot order (here: order of first appearance)"""
        names = []
        for ln in self.s2['text'].split(b'\n'):
            f = ln.split(b',')
            if len(f) > 1 and f[1] not in names:
                names.append(f[1])
        return list(enumerate(names))

    def side2_set_devs(self, devs):
        n = len(self.side2_dev_names())
        self.s2['secsize'] = [int(devs.get(i, (512, 0))[0]) for i in range(n)]
        self.s2['ui'] = [1 if devs.get(i, (512, 0))[1] else 0 for i in range(n)]

    def side2_emit(self, table, what):
        out, n, _ = self._side2(what)
        return (out[2:] if what in (2, 3) else out), n

    def diskstat_profile(self, csv_text):
        """plain-Python stand-in for sofa_b200_diskstat_profile (test infrastructure only)"""
        rows = [ln.split(',') for ln in bytes(csv_text).decode().split('\n')[1:] if ln]
        out = []
        for dev in dict.fromkeys(r[1] for r in rows):
            g = [r for r in rows if r[1] == dev]
            n = len(g)
            q_lo, q_hi, mean = [], [], []
            for col in (8, 9, 10):
                v = sorted(float(r[col]) for r in g)
                for k in (1, 2, 3):
                    num = (n - 1) * k
                    q_lo.append(v[num // 4]); q_hi.append(v[num // 4 + (1 if num % 4 else 0)])
            for col in (8, 9, 10, 2, 3, 4, 7):
                s_, c_ = 0.0, 0.0
                for r in g:                                           # Kahan, row order (pandas group_mean)
                    y = float(r[col]) - c_
                    t = s_ + y
                    c_ = (t - s_) - y
                    s_ = t
                mean.append(s_ / n)
            out.append((dev, n, q_lo, q_hi, mean))
        return out

    def blktrace_latency(self, csv_text):
        """plain-Python stand-in for sofa_b200_blktrace_latency (test infrastructure only)"""
        v = sorted(float(f[2]) for f in (ln.split(',') for ln in bytes(csv_text).decode().split('\n')[1:] if ln) if f[1] == 'C')
        n = len(v)
        lo, hi = [0.0] * 3, [0.0] * 3
        for k in (1, 2, 3):
            if n:
                num = (n - 1) * k
                lo[k - 1], hi[k - 1] = v[num // 4], v[num // 4 + (1 if num % 4 else 0)]
        return n, lo, hi

    # ---- side tables: numpy stand-in for the device window reduction (test infrastructure only)
    def upload(self, kind, text, data_off=0):
        self.src[kind] = (bytes(text)[data_off:], None)

    def parse(self, kind, params):
        return self.src[kind][0].count(b"\n")

    def net_profile(self, stride, cap=0):
        """pandas stand-in for sofa_b200_net_profile (test infrastructure only)."""
        import io
        import pandas as pd
        cols = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload", "bandwidth", "pkt_src",
                "pkt_dst", "pid", "tid", "name", "category"]
        df = pd.read_csv(io.BytesIO(self.csv(2, stride).tobytes()), names=cols)
        g = df.groupby("name")["duration"]
        t = 0
        for key, _ in g:
            if key.find("network:tcp:") != -1:
                t = t + g.get_group(key).sum()
        ok = df.dropna(subset=["payload"])
        pairs = {(int(a), int(b)): int(v) for (a, b), v in ok.groupby(["pkt_src", "pkt_dst"])["payload"].sum().items()}
        return float(t), int(ok.payload.sum()), pairs

    def row_flags(self, kind, stride):
        return self._run(kind, 0 if kind == 2 else 1, stride, 0x1fff)[1]

    def nvsmi_profile(self, roi=None):
        """numpy stand-in for sofa_b200_nvsmi_profile (test infrastructure only)."""
        import io
        import pandas as pd
        cols = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload", "bandwidth", "pkt_src",
                "pkt_dst", "pid", "tid", "name", "category"]
        df = pd.read_csv(io.BytesIO(self.src[4][0]), names=cols)
        if roi:
            df = df[(df.timestamp > roi[0]) & (df.timestamp <= roi[1])]
        groups, order = {}, []
        for d in df.deviceId:
            if int(d) not in order:
                order.append(int(d))
        for (d, e), g in df.groupby(['deviceId', 'event']):
            v = sorted(int(x) for x in g.duration)
            qs = []
            for q in (1, 2, 3):
                num = (len(v) - 1) * q
                lo = num // 4
                qs.append((float(v[lo]), float(v[lo + 1 if num % 4 else lo])))
            groups[(int(d), int(e))] = (len(v), sum(v), qs)
        if len(df) == 0:
            return groups, order, 0.0, 0.0, 0
        return groups, order, float(df.timestamp.iloc[0]), float(df.timestamp.iloc[-1]), len(df)

    def window_reduce(self, ends, delta):
        import io
        import types
        import pandas as pd
        cols = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload", "bandwidth", "pkt_src",
                "pkt_dst", "pid", "tid", "name", "category"]
        df = {k: (pd.read_csv(io.BytesIO(self.src[k][0]), names=cols) if self.src[k][0] else pd.DataFrame([], columns=cols))
              for k in (3, 4, 5)}
        out = []
        for t in ends:
            wb, we = t - delta, t
            mp = df[3][(df[3].timestamp > wb) & (df[3].timestamp <= we)]
            nv = df[4][(df[4].timestamp > wb) & (df[4].timestamp <= we) & (df[4].event == 0)]
            ns = df[5][(df[5].timestamp > wb) & (df[5].timestamp <= we)]
            r = [[int(x) for x in nm.split(':')[1].split('|')[1:6]] for nm in mp.name]
            w = types.SimpleNamespace()
            w.mp_count = len(mp)
            for i, nm in enumerate(('usr', 'sys', 'idl', 'iow')):
                setattr(w, nm + '_max', max([x[i] for x in r]) if r else 0)
            w.usr_sum = sum(x[0] for x in r); w.sys_sum = sum(x[1] for x in r); w.irq_sum = sum(x[4] for x in r)
            cpu = [x[0] + x[1] + x[4] for x in r]
            w.cpu_max = max([0] + cpu); w.cpu_min = min([100] + cpu)
            w.nv_count = len(nv); w.nv_max = float(nv.duration.max()) if len(nv) else 0.0
            w.nv_min = float(nv.duration.min()) if len(nv) else 0.0
            w.nv_sum = int(nv.duration.sum()); w.nv_ndev = len(set(nv.deviceId))
            tx, rx = ns[ns.event == 0], ns[ns.event == 1]
            w.tx_count, w.rx_count = len(tx), len(rx)
            w.tx_sum, w.rx_sum = float(tx.bandwidth.sum()), float(rx.bandwidth.sum())
            out.append(w)
        return out, len(set(df[4].deviceId))

    def viz_mangled(self, stride):
        text, _ = self.src[0]
        names = set()
        for i, line in enumerate(text.decode().split('\n')):
            if not line or i % stride:
                continue
            tok = line.split()
            sh = 1 if (tok[1].startswith('[') and tok[1].endswith(']')) else 0
            x = '[%s]' % tok[3 + sh].replace('-', '_') + tok[5 + sh] + tok[6 + sh]
            v = x[x.find(']') + 1: x.find('+')].split('@')[0]
            if v.startswith('_Z'):
                names.add(v)
        return sorted(names)

    def set_demangled(self, pairs):
        pairs = sorted(pairs, key=lambda p: self.lib.hs_fnv(p[0].encode(), ctypes.c_long(len(p[0].encode()))))
        hashes = np.array([self.lib.hs_fnv(m.encode(), ctypes.c_long(len(m.encode()))) for m, _ in pairs], dtype=np.uint64)
        blob = b"".join(d.encode() for _, d in pairs)
        off = np.zeros(len(pairs) + 1, dtype=np.int64)
        for i, (_, d) in enumerate(pairs):
            off[i + 1] = off[i] + len(d.encode())
        self._keep += [hashes, off, blob]
        self.dict = HsDict(len(pairs), hashes.ctypes.data, off.ctypes.data, blob)
