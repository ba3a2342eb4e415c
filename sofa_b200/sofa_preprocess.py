"""Drop-in `sofa_preprocess(cfg)` (reference: bin/sofa_preprocess.py:377-2106).

Same entry point, same `cfg` attributes, same files in `cfg.logdir`; every per-record
step (line split, tokenise, float/int parse, time-base maths, tagging, name synthesis,
keyword filters, down-sampling, CSV and report.js text) runs in sm_100a CUDA kernels
behind libsofa_b200.so.  The host keeps only what the reference does once per file:
reading side files, the nvprof header rows, demangling the unique mangled symbols, and
writing the series wrappers of report.js.  There is no CPU fallback: without the CUDA
library this module raises.

Sources handled: perf.script (a1), nvprof gpu trace (a2) and API trace (a4), tcpdump text (a3), and every recorder
side file of SURVEY.md §8(f) N1: mpstat.txt, vmstat.txt, nvsmi_query.txt, nvsmi.txt, netstat.txt, diskstat.txt,
blktrace.txt, strace.txt, pystacks.txt.  Files go to and from the device through the engine's pinned staging buffer;
the host never holds a whole trace or output (PlainIO.streaming).
"""
import ctypes
import glob
import os
import subprocess

from . import hostlogic as hl
from .hostlogic import SRC_PERF, SRC_GPU, SRC_NET, PreprocessError
from .sofa_print import print_main_progress, print_progress, print_info, print_warning

sofa_fieldnames = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload",
                   "bandwidth", "pkt_src", "pkt_dst", "pid", "tid", "name", "category"]
CSV_HEADER = (",".join(sofa_fieldnames) + "\n").encode()

S2_DISK, S2_BLK, S2_STRACE, S2_PYSTACKS = 0, 1, 2, 3        # second group of side files (sofa_b200_side2_*)
SYS_BLOCK = '/sys/block/'        # where diskstat's sector sizes are looked up (:563-570); tests point it at an empty directory
VECTOR_HEADER = b"time,dev,r_iops,w_iops,iops,r_time,w_time,await_time,d_read,d_write,d_disk_total\n"
RAW_MPSTAT, RAW_NVSMI, RAW_NETSTAT, RAW_VMSTAT, RAW_DMON = 6, 7, 8, 9, 10
VMSTAT_LISTS = ('bi', 'bo', 'in', 'cs', 'wa', 'st', 'usr', 'sys')

_libstdcxx = None


def demangle(name):
    """cxxfilt.demangle: libstdc++ __cxa_demangle on names starting with _Z (:1811-1816)."""
    global _libstdcxx
    if not name.startswith('_Z'):
        return name
    if _libstdcxx is None:
        _libstdcxx = ctypes.CDLL('libstdc++.so.6')
        _libstdcxx.__cxa_demangle.restype = ctypes.c_void_p
    status = ctypes.c_int(0)
    p = _libstdcxx.__cxa_demangle(name.encode(), None, None, ctypes.byref(status))
    if status.value != 0 or not p:
        raise PreprocessError("cxxfilt.InvalidName: %s" % name)
    s = ctypes.string_at(p).decode()
    libc = ctypes.CDLL('libc.so.6')
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(p)
    return s


class SeriesOut:
    """One report.js series (sofa_models.py:1-7): device-produced record bytes (an array, or a HeldOutput still on the device)."""
    def __init__(self, name, title, color, records=None, nrec=0):
        self.name, self.title, self.color = name, title, color
        self.records, self.nrec = records, nrec


class HeldOutput:
    """Record bytes of a series parked on the device (sofa_b200_hold_output) until report.js is written: the reference's
    series order is not the order the series are produced in, and the host need not hold the whole file in memory."""
    def __init__(self, eng, handle, nbytes):
        self.eng, self.handle, self.nbytes = eng, handle, nbytes

    def write_to(self, f):
        self.eng.hold_write(self.handle, self.nbytes, f)
        self.eng.hold_release(self.handle)


class DeviceOutput:
    """The output buffer of the emit that just ran, written to a file straight from the device (pinned staging pieces)."""
    def __init__(self, eng, which, nbytes):
        self.eng, self.which, self.nbytes = eng, which, nbytes

    def write_to(self, f):
        self.eng.write_output(self.which, self.nbytes, f)


def _put(f, data):
    if hasattr(data, 'write_to'):
        data.write_to(f)
    else:
        f.write(memoryview(data))


def _json_str(s):
    import json
    return json.dumps(s)


def write_report_js(series, path):
    """traces_to_json (:343-374): records come from the device, wrappers from here."""
    if len(series) == 0:
        return
    with open(path, 'wb') as f:
        for s in series:
            if s.nrec > 0:
                f.write(("%s = {\"name\": %s, \"color\": %s, \"data\": [" %
                         (s.name, _json_str(s.title), _json_str(s.color))).encode())
                _put(f, s.records)
                f.write(b"]}")
            f.write(b"\n\n")
        f.write(b"sofa_traces = [ ")
        for s in series:
            if s.nrec > 0:
                f.write((s.name + ",").encode())
        f.write(b" ]")


def _read_bytes(path):
    with open(path, 'rb') as f:
        return f.read()


def _check_keyword(kw):
    if any(ch in kw for ch in '.^$*+?{}[]\\|()'):
        raise PreprocessError("filter keyword %r contains regex metacharacters; the device "
                              "matcher implements literal substring search only" % kw)
    if len(kw.encode()) > 32 or len(kw) == 0:
        raise PreprocessError("filter keyword %r must be 1..32 bytes" % kw)


def gpu_first_timestamp(nvvp):
    """First CUPTI activity start / 1e9 (:1355-1380) via stdlib sqlite3; 0 when the tables
    cannot be read (the reference swallows every exception)."""
    import sqlite3
    bases = []
    for tbl in ('CUPTI_ACTIVITY_KIND_MEMSET', 'CUPTI_ACTIVITY_KIND_MEMCPY',
                'CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL', 'CUPTI_ACTIVITY_KIND_KERNEL'):
        try:
            con = sqlite3.connect('file:%s?mode=ro' % nvvp, uri=True)
            try:
                row = con.execute('SELECT start FROM %s LIMIT 1' % tbl).fetchone()
            finally:
                con.close()
            if row is not None:
                bases.append(row[0])
        except Exception:
            pass
    return sorted(bases)[0] * 1.0 / 1e+9 if bases else 0


def _write(path, *parts):
    with open(path, 'wb') as f:
        for p in parts:
            f.write(memoryview(p))


def _net_filter_series(cfg, eng, ratio):
    """--network_filters (:1214-1231): cfg.net_filters[0] is the node of interest; one series "to_<f>" (payload > 0,
    pkt_src == node, pkt_dst == f) per other address, then one "from_<f>" each.  Equality tests run on the device over
    the packed addresses; the colours are random in the reference too."""
    import random

    def packed(s):
        v = float(s)
        return int(v) if v == int(v) else None
    base = packed(cfg.net_filters[0])
    others = list(cfg.net_filters[1:])
    out = []
    for direction, bit0 in (('to_', 0), ('from_', 4)):
        for c0 in range(0, len(others), 4):
            chunk = others[c0:c0 + 4]
            vals = [packed(f) for f in chunk]
            if base is not None:
                eng.set_net_filters(ratio, base, [v if v is not None else -(1 << 62) for v in vals])
            for k, f in enumerate(chunk):
                color = 'rgba(%s,%s,%s,0.8)' % (random.randint(0, 255), random.randint(0, 255), random.randint(0, 255))
                rec, nrec = (eng.json(SRC_NET, ratio, bit0 + k) if base is not None and vals[k] is not None else (None, 0))
                out.append(SeriesOut('%s%s' % (direction, f), '[keyword]%s%s' % (direction, f), color, rec, nrec))
    return out


def _side_sources(cfg, eng, logdir):
    """mpstat.txt (:482-546), nvsmi_query.txt (:1013-1084), netstat.txt (:1235-1337): rows, CSV text and
    report.js records are produced on the device; this function only names the files."""
    out = {}
    if os.path.isfile(logdir + 'mpstat.txt'):
        eng.side_preprocess(RAW_MPSTAT, _read_bytes(logdir + 'mpstat.txt'), cfg.time_base, cfg.absolute_timestamp)
        aux, _ = eng.side_aux(RAW_MPSTAT)
        _write(logdir + 'usr_sys.csv', b"time,usr,sys\n", aux)
        body, nrows = eng.side_csv(RAW_MPSTAT)
        if nrows > 0:
            _write(logdir + 'mpstat.csv', CSV_HEADER, body)
            out['mpstat_usr'] = eng.side_json(RAW_MPSTAT)
        else:
            print_warning(cfg, 'Empty list cannot be exported to mpstat.csv!')
    # `cpuid`: a local of the reference that leaks from section to section and ends up in the last column of the strace
    # and pystacks rows (blktrace.txt leaves a string, vmstat.txt -1, the nvidia-smi files a gpu index)
    cfg._sofa_b200_cpuid = None
    if os.path.isfile(logdir + 'blktrace.txt'):                          # :688-781
        text = _read_bytes(logdir + 'blktrace.txt')
        print_info(cfg, "Length of blktrace = %d" % (text.count(b'\n') + (1 if text and not text.endswith(b'\n') else 0)))
        if len(text) > 0:
            info = eng.side2_preprocess(S2_BLK, text, cfg.time_base, cfg.absolute_timestamp)
            d_rows, nd = eng.side2_emit(S2_BLK, 0)
            if nd > 0:
                _write(logdir + 'blktrace.csv', CSV_HEADER, d_rows)
                out['blktrace_starting_block'] = eng.side2_emit(S2_BLK, 2)
            else:
                print_warning(cfg, 'Empty list cannot be exported to blktrace.csv!')
            c_rows, nc = eng.side2_emit(S2_BLK, 1)
            if nc > 0:
                with open(logdir + 'blktrace.csv', 'ab') as f:             # appended without a header (mode 'a')
                    f.write(memoryview(c_rows))
                out['blktrace_time'] = eng.side2_emit(S2_BLK, 3)
            else:
                print_warning(cfg, 'Empty list cannot be exported to blktrace.csv!')
            if info['flags'] & 1:
                print_warning(cfg, 'blktrace maybe record failed!')
            a, b = info['last_span']
            if b > a:
                cfg._sofa_b200_cpuid = bytes(text[a:b])
    if os.path.isfile(logdir + 'diskstat.txt'):                          # :548-673
        text = _read_bytes(logdir + 'diskstat.txt')
        eng.side2_preprocess(S2_DISK, text, cfg.time_base, cfg.absolute_timestamp)
        flt = getattr(cfg, 'diskstat_filters', None)
        devs = {}
        for slot, name in eng.side2_dev_names():
            secsize = 512
            try:
                with open(SYS_BLOCK + name.decode() + '/queue/hw_sector_size') as f:
                    import re
                    secsize = int(re.match(r"\d+", f.readline()).group())
            except Exception:
                pass
            devs[slot] = (secsize, (not flt) or name.decode() == flt)
        eng.side2_set_devs(devs)
        if flt:
            print(flt)                                                   # :662
        body, _ = eng.side2_emit(S2_DISK, 0)
        _write(logdir + 'diskstat.csv', CSV_HEADER, body)
        _write(logdir + 'diskstat_vector.csv', VECTOR_HEADER, eng.side2_emit(S2_DISK, 4)[0])
        _write(logdir + 'diskstat_vector_ui.csv', VECTOR_HEADER, eng.side2_emit(S2_DISK, 5)[0])
        out['diskstat'] = eng.side2_emit(S2_DISK, 2)
    if os.path.isfile(logdir + 'vmstat.txt'):
        text = _read_bytes(logdir + 'vmstat.txt')
        print_info(cfg, "Length of vmstat_traces = %d" % text.count(b'\n'))
        if len(text) > 0:
            _, ncand, _ = eng.side_preprocess(RAW_VMSTAT, text, cfg.time_base, cfg.absolute_timestamp)
            if ncand > 0:
                cfg._sofa_b200_cpuid = -1
                _write(logdir + 'vmstat.csv', CSV_HEADER, *[eng.side_csv(RAW_VMSTAT, i)[0] for i in range(8)])
                for i, nm in enumerate(VMSTAT_LISTS):
                    if nm in ('cs', 'bi', 'bo') or (cfg.enable_vmstat and nm in ('usr', 'sys', 'in')):
                        out['vmstat_' + nm] = eng.side_json(RAW_VMSTAT, i)
            else:
                for _ in range(8):
                    print_warning(cfg, 'Empty list cannot be exported to vmstat.csv!')
    if os.path.isfile(logdir + 'nvsmi_query.txt'):
        text = _read_bytes(logdir + 'nvsmi_query.txt')
        if len(text) == 0:
            raise PreprocessError("nvsmi_query.txt is empty (next(f) raises StopIteration in the reference)")
        nl = text.find(b'\n')
        data_off = nl + 1 if nl >= 0 else len(text)                       # next(f): the header line
        n_lines, ncand, no_data = eng.side_preprocess(
            RAW_NVSMI, text, cfg.time_base, cfg.absolute_timestamp, hl.nvsmi_local_offset(text, data_off),
            cfg.nvsmi_time_zone * 3600, data_off)
        if no_data:
            print_warning(cfg, 'No nvsmi query data.')
        else:
            print_info(cfg, "Length of nvsmi_query_traces = %d" % n_lines)
            if ncand > 0:
                cfg._sofa_b200_cpuid = hl.last_nvsmi_gpu(text, data_off)
                sm, _ = eng.side_csv(RAW_NVSMI, 0)
                mem, _ = eng.side_csv(RAW_NVSMI, 1)
                _write(logdir + 'nvsmi_trace.csv', CSV_HEADER, sm, mem)
                out['nvsmi_sm'] = eng.side_json(RAW_NVSMI, 0)
                out['nvsmi_mem'] = eng.side_json(RAW_NVSMI, 1)
    if os.path.isfile(logdir + 'nvsmi.txt'):                              # `nvidia-smi dmon`: enc / dec rows (:1092-1183)
        text = _read_bytes(logdir + 'nvsmi.txt')
        n_lines, ncand, no_data = eng.side_preprocess(RAW_DMON, text, cfg.time_base, cfg.absolute_timestamp)
        if no_data:
            print_warning(cfg, 'No nvsmi data.')
        else:
            print_info(cfg, "Length of nvsmi_traces = %d" % n_lines)
            last = hl.last_dmon_gpu(text)
            if last is not None:
                cfg._sofa_b200_cpuid = last
            enc, nrows = eng.side_csv(RAW_DMON, 0) if ncand > 0 else (b'', 0)
            if nrows > 0:
                cfg.nvsmi_data = True
                dec, _ = eng.side_csv(RAW_DMON, 1)
                with open(logdir + 'nvsmi_trace.csv', 'ab') as f:          # appended below the sm / mem rows, no header
                    f.write(memoryview(enc))
                    f.write(memoryview(dec))
                if cfg.enable_encode_decode:
                    out['nvsmi_enc'] = eng.side_json(RAW_DMON, 0)
                    out['nvsmi_dec'] = eng.side_json(RAW_DMON, 1)
            else:
                print_warning(cfg, "Program exectution time is fewer than 3 seconds, so nvsmi trace analysis will not be displayed.")
    if os.path.isfile(logdir + 'netstat.txt'):
        text = _read_bytes(logdir + 'netstat.txt')
        if len(text) > 0:
            eng.side_preprocess(RAW_NETSTAT, text, cfg.time_base, cfg.absolute_timestamp)
            aux, _ = eng.side_aux(RAW_NETSTAT)
            _write(logdir + 'netbandwidth.csv', b"time,tx_bandwidth,rx_bandwidth\n", aux)
            tx, _ = eng.side_csv(RAW_NETSTAT, 0)
            rx, _ = eng.side_csv(RAW_NETSTAT, 1)
            _write(logdir + 'netstat.csv', CSV_HEADER, tx, rx)
            out['tx_bandwidth'] = eng.side_json(RAW_NETSTAT, 0)
            out['rx_bandwidth'] = eng.side_json(RAW_NETSTAT, 1)
    return out


def _strace_pystacks(cfg, eng, logdir, ratio):
    """strace.txt (:1618-1704) and pystacks.txt (:1706-1761): rows, CSV text and report.js records come from the device;
    `cpuid` is whatever the side-file sections left behind (NameError in the reference when none did)."""
    out = {}
    cpuid = getattr(cfg, '_sofa_b200_cpuid', None)
    total = 0
    if os.path.isfile(logdir + 'strace.txt'):
        text = _read_bytes(logdir + 'strace.txt')
        n_lines = text.count(b'\n') + (1 if text and not text.endswith(b'\n') else 0)
        print_info(cfg, "Length of straces = %d" % n_lines)
        if n_lines > 1:
            info = eng.side2_preprocess(S2_STRACE, text, cfg.time_base, cfg.absolute_timestamp, ratio, getattr(cfg, 'pid', 0),
                                        getattr(cfg, 'strace_min_time', 1e-6), cpuid)
            total = info['total']
            body, nrows = eng.side2_emit(S2_STRACE, 0)
            if nrows > 0 and cpuid is None:
                raise PreprocessError("strace.txt: name 'cpuid' is not defined (the reference raises NameError: no side file "
                                      "left a cpu / gpu id behind)")
            print_info(cfg, 'strace.txt reading is done.')
            if nrows > 0:
                _write(logdir + 'strace.csv', CSV_HEADER, body)
                out['strace'] = eng.side2_emit(S2_STRACE, 2)
    print_info(cfg, 'Total strace duration: %.3lf' % total)
    if os.path.isfile(logdir + 'pystacks.txt'):
        text = _read_bytes(logdir + 'pystacks.txt')
        info = eng.side2_preprocess(S2_PYSTACKS, text, cfg.time_base, cfg.absolute_timestamp, category=cpuid)
        if info['n_kept'] == 0:
            raise PreprocessError("pystacks.txt: no stack outside 'idle' (max() of an empty dict raises ValueError in the reference)")
        if info['n_kept'] > 1:                                       # the reference builds one row per remaining stack
            if cpuid is None:
                raise PreprocessError("pystacks.txt: name 'cpuid' is not defined (the reference raises NameError)")
            body, nrows = eng.side2_emit(S2_PYSTACKS, 0)
            if nrows > 0:
                _write(logdir + 'pystacks.csv', CSV_HEADER, body)
                out['pystacks'] = eng.side2_emit(S2_PYSTACKS, 2)
            else:
                print_warning(cfg, 'Empty list cannot be exported to pystacks.csv!')
    return out


class _Timing(dict):
    """wall-clock seconds per phase of one sofa_preprocess call (cfg._timing; reported by bench.py's e2e_entrypoints)"""
    def add(self, key, t0):
        import time
        self[key] = self.get(key, 0.0) + (time.perf_counter() - t0)


def _timed_engine(eng, tm):
    """wrap the engine's bulk calls with the phase timers"""
    import time

    class T:
        def __getattr__(self, name):
            f = getattr(eng, name)
            phase = {'load': 'upload+parse', 'csv': 'csv emit+fetch', 'json': 'json emit+fetch', 'json_series': 'json emit+fetch',
                     'set_filters': 'filters', 'viz_mangled': 'viz names', 'set_demangled': 'viz names',
                     'side_preprocess': 'side files', 'side_csv': 'side files', 'side_json': 'side files', 'side_aux': 'side files',
                     'side2_preprocess': 'side files', 'side2_emit': 'side files', 'side2_set_devs': 'side files'}.get(name)
            if phase is None or not callable(f):
                return f

            def g(*a, **k):
                t0 = time.perf_counter()
                try:
                    return f(*a, **k)
                finally:
                    tm.add(phase, t0)
            return g
    return T()


class _FileText:
    """A source file that goes to the device piece by piece through the engine's pinned staging buffer (never whole in
    host memory)."""
    def __init__(self, path):
        self.path, self.size = path, os.path.getsize(path)

    def __len__(self):
        return self.size

    def count(self, what):                      # (verbose length prints only)
        n = 0
        with open(self.path, 'rb') as f:
            for piece in iter(lambda: f.read(1 << 24), b''):
                n += piece.count(what)
        return n


def _streaming_engine(eng):
    """csv / json / json_series / load of the engine with device-resident results: CSV bodies are written from the
    output buffer of the emit that just ran, report.js records stay parked on the device until the file is assembled
    in series order, source files are read through the staging buffer."""
    class S:
        def __getattr__(self, name):
            return getattr(eng, name)

        def load(self, kind, text, params, data_off=0):
            if isinstance(text, _FileText):
                return eng.load_file(kind, text.path, params, data_off)
            return eng.load(kind, text, params, data_off)

        def csv(self, kind, stride):
            return DeviceOutput(eng, 0, eng.csv_device(kind, stride))

        def json(self, kind, stride, filt):
            n, nrec = eng.json_device(kind, stride, filt)
            return HeldOutput(eng, eng.hold_output(1), n), nrec

        def json_series(self, kind, stride, kws):
            n, nrec, counts = eng.json_series_device(kind, stride, kws)
            return HeldOutput(eng, eng.hold_output(1), n), nrec, counts
    return S()


class PlainIO:
    """File access of an unsharded run.  sofa_b200.sharded.ShardIO has the same interface for one logical trace split
    over several ranks (byte ranges in, part files out, dtype flags all-reduced)."""
    rank, world = 0, 1
    tm = None
    streaming = True            # files <-> device through the engine's pinned staging buffer when the engine offers it

    def head(self, path, n):
        with open(path, 'rb') as f:
            return f.read(n)

    def read(self, path, data_off=0):
        """(text of this rank's rows incl. the header on rank 0, offset of the first row in it)"""
        import time
        t0 = time.perf_counter()
        try:
            if getattr(self, 'lazy', False):
                return _FileText(path), data_off          # read later, by the engine, into its staging buffer
            return _read_bytes(path), data_off
        finally:
            if self.tm is not None:
                self.tm.add('read files', t0)

    def size(self, path):
        return os.path.getsize(path)

    def sync_flags(self, eng, kind, stride):
        pass

    def write_csv(self, path, header, body):
        import time
        t0 = time.perf_counter()
        with open(path, 'wb') as f:
            f.write(header)
            _put(f, body)
        if self.tm is not None:
            self.tm.add('write csv', t0)

    def series(self, s):
        return s

    def finish(self, series, path):
        write_report_js(series, path)


def sofa_preprocess(cfg, engine=None, io=None):
    print_main_progress('SOFA preprocessing...')
    own_engine = engine is None
    if own_engine:
        from .engine import CudaEngine
        engine = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
    try:
        if io is None:
            # bounded device memory: a source larger than the segment budget is walked in newline-aligned segments
            # (sofa_b200.sharded.sofa_preprocess_segmented); text + columns + outputs of a segment are ~5x its text
            budget = int(os.environ.get('SOFA_B200_SEGMENT_BYTES', 16 << 30))
            big = [fn for fn in ('perf.script', 'gputrace.tmp', 'net.tmp')
                   if os.path.isfile(cfg.logdir + fn) and os.path.getsize(cfg.logdir + fn) > budget]
            if big and not getattr(cfg, 'cuda_api_tracing', False):
                from .sharded import sofa_preprocess_segmented
                return sofa_preprocess_segmented(cfg, engine, budget)
        return _preprocess(cfg, engine, io or PlainIO())
    finally:
        if own_engine:
            if getattr(cfg, 'keep_engine', False):
                cfg._sofa_b200_engine = engine      # `sofa report`: sofa_analyze reuses the resident columns
            else:
                engine.close()


def _preprocess(cfg, eng, io):
    import time
    tm = cfg._timing = _Timing()
    if getattr(io, 'streaming', False) and hasattr(eng, 'hold_output') and os.environ.get('SOFA_B200_STREAM_IO', '1') != '0':
        eng.hold_release(-1)
        eng = _streaming_engine(eng)
        io.lazy = True
    eng = _timed_engine(eng, tm)
    io.tm = tm
    cfg.time_base = 0
    logdir = cfg.logdir
    with open(logdir + 'misc.txt', 'r') as f:
        lines = f.readlines()
        if len(lines) == 4:
            cfg.pid = int(lines[3].split()[1])
        else:
            print_warning(cfg, 'Incorrect misc.txt content. Some profiling information may not be available.')
    if int(os.system('command -v perf 1> /dev/null')) == 0 and os.path.isfile(logdir + 'perf.data'):
        with open(logdir + 'perf.script', 'w') as logfile:
            subprocess.call('perf script -F time,pid,tid,event,ip,sym,dso,symoff,period,brstack,brstacksym '
                            '-i %s/perf.data --kallsym %s/kallsyms' % (logdir, logdir),
                            shell=True, stdout=logfile, stderr=subprocess.DEVNULL)
    cfg.time_base = hl.read_time_base(logdir, cfg.cpu_time_offset)
    print_info(cfg, 'Time offset applied to timestamp (s):' + str(cfg.cpu_time_offset))
    print_info(cfg, 'SOFA global time base (s):' + str(cfg.time_base))
    mhz_xp, mhz_fp = hl.read_cpuinfo(logdir)
    ratio = int(cfg.plot_ratio)
    if ratio < 1:
        raise PreprocessError("plot_ratio must be >= 1")

    if io.world > 1 and ratio != 1:
        if not hasattr(io, 'aligned_range'):
            raise NotImplementedError("this way of splitting a trace needs plot_ratio 1 (the down-sampling phase of a part is not carried)")
        io.ratio = ratio              # shards / segments start at rows whose global index is a multiple of plot_ratio (ShardIO.aligned_range)
    side = _side_sources(cfg, eng, logdir) if io.rank == 0 else {}
    net_series = SeriesOut('net_trace', 'NET', 'blue')
    net_filter_series = []
    gpu_series = SeriesOut('gpu_kernel_trace', 'GPU kernel', 'rgba(0,180,0,0.8)')
    gpu_filter_series = []
    cpu_series = []

    # ---------------- network (:1187-1232): nettrace.csv holds the down-sampled rows
    if os.path.isfile('%s/sofa.pcap' % logdir):
        if not os.path.isfile(logdir + 'net.tmp') or os.path.getsize(logdir + 'sofa.pcap') > 0:
            with open(logdir + 'net.tmp', 'w') as f:
                subprocess.check_call(["tcpdump", "-q", "-n", "-tt", "-r", "%s/sofa.pcap" % logdir],
                                      stdout=f, stderr=subprocess.DEVNULL)
        text, _ = io.read(logdir + 'net.tmp')
        if cfg.verbose:
            print_info(cfg, "Length of net_traces = %d" % text.count(b'\n'))
        if io.size(logdir + 'net.tmp') > 0:
            p = hl.base_params(SRC_NET, cfg.time_base, cfg.absolute_timestamp, 0.0)
            eng.load(SRC_NET, text, p)
            io.sync_flags(eng, SRC_NET, ratio)
            body = eng.csv(SRC_NET, ratio)
            io.write_csv(logdir + 'nettrace.csv', CSV_HEADER, body)
            net_series.records, net_series.nrec = eng.json(SRC_NET, ratio, -1)
            if getattr(cfg, 'net_filters', None):
                net_filter_series = _net_filter_series(cfg, eng, ratio)
    else:
        print_warning(cfg, "no network traces were recorded.")

    # ---------------- GPU (:1343-1457)
    nvvps = glob.glob(logdir + "gputrace*[0-9].nvvp")
    if len(nvvps) > 1:
        raise NotImplementedError("more than one gputrace*.nvvp: the reference overwrites gputrace.csv "
                                  "per file in glob order (SURVEY.md §7); not supported")
    for nvvp in nvvps:
        print_progress(cfg, "Read " + nvvp + " by nvprof -- begin")
        if not os.path.isfile(logdir + 'gputrace.tmp') or os.path.getsize(nvvp) > 0:
            with open(logdir + "gputrace.tmp", "w") as f:
                subprocess.call(["nvprof", "--csv", "--print-gpu-trace", "-i", nvvp], stderr=f)
        t_glb_gpu_base = gpu_first_timestamp(nvvp)
        print_info(cfg, "Timestamp of the first GPU trace = " + str(t_glb_gpu_base))
        print_progress(cfg, "Read " + nvvp + " by nvprof -- end")
        head = io.head(logdir + 'gputrace.tmp', 1 << 16)
        hdr = hl.gpu_header(head) if len(head) > 0 else None
        if hdr is None:
            print_warning(cfg, "gputrace existed, but no kernel traces were recorded.")
            continue
        upd, data_off = hdr
        text, data_off = io.read(logdir + 'gputrace.tmp', data_off)
        p = hl.base_params(SRC_GPU, cfg.time_base, cfg.absolute_timestamp, t_glb_gpu_base - upd.pop('t_base'))
        p.update(upd)
        p['n_cudaproc'] = 1
        n_rows, _ = eng.load(SRC_GPU, text, p, data_off)
        print_info(cfg, "Length of gpu_traces = %d" % n_rows)
        io.sync_flags(eng, SRC_GPU, 1)
        body = eng.csv(SRC_GPU, 1)
        io.write_csv(logdir + 'gputrace.csv', CSV_HEADER, body)
        kws = [flt.keyword for flt in cfg.gpu_filters]
        for kw in kws:
            _check_keyword(kw)
        if kws and ratio == 1:
            # the filters run on the frame of the main series: one formatting pass matches them too, and the
            # filter series below are gathered from its records
            gpu_series.records, gpu_series.nrec, _ = eng.json_series(SRC_GPU, 1, kws)
        else:
            gpu_series.records, gpu_series.nrec = eng.json(SRC_GPU, ratio, -1)
            if kws:
                eng.set_filters(SRC_GPU, 1, kws)               # on the FULL frame (:1446-1452)
        if kws:
            for i, flt in enumerate(cfg.gpu_filters):
                rec, nrec = eng.json(SRC_GPU, 1, i)
                gpu_filter_series.append(SeriesOut(flt.keyword, '[keyword]' + flt.keyword, flt.color, rec, nrec))

    # ---------------- CUDA API trace (:1459-1543), only with --cuda_api_tracing.  The rows go through
    # the gpu source slot of the engine in api_mode (the gpu trace above is finished by now).
    api_series = None
    if getattr(cfg, 'cuda_api_tracing', False):
        api_series = SeriesOut('cuda_api_trace', 'CUDA API', 'DarkSlateGray')
        for nvvp in nvvps:
            if not os.path.isfile(logdir + 'cuda_api_trace.tmp') or os.path.getsize(nvvp) > 0:
                with open(logdir + "cuda_api_trace.tmp", "w") as f:
                    subprocess.call(["nvprof", "--csv", "--print-api-trace", "-i", nvvp], stderr=f)
            text = _read_bytes(logdir + 'cuda_api_trace.tmp')
            hdr = hl.api_header(text[:1 << 16]) if len(text) > 0 else None
            if hdr is None:
                continue
            upd, data_off = hdr
            t_base = hl.api_t_base(text, data_off)            # first_corid = 1 (no sqlite RUNTIME table)
            p = hl.base_params(SRC_GPU, cfg.time_base, cfg.absolute_timestamp, 0 - t_base)
            p.update(upd)
            n_rows, _ = eng.load(SRC_GPU, text, p, data_off)
            print_info(cfg, "Length of cuda_api_traces = %d" % n_rows)
            with open(logdir + 'cuda_api_trace.csv', 'wb') as f:
                f.write(CSV_HEADER)
                _put(f, eng.csv(SRC_GPU, 1))
            api_series.records, api_series.nrec = eng.json(SRC_GPU, ratio, -1)
        cfg._sofa_b200_gpu_slot_is_api = True

    # ---------------- CPU (:1544-1551, 1766-1838)
    cfg.perf_events = hl.read_perf_events(logdir)
    print_info(cfg, 'perf_events_used: %s' % (cfg.perf_events))
    if glob.glob(logdir + "cuhello*[0-9].nvvp"):
        print_warning(cfg, "cuhello nvvp clock sync needs nvprof's sqlite export; using perf_timebase.txt")
    t_offset = hl.read_perf_timebase(logdir)
    try:
        if not os.path.isfile(logdir + 'perf.script'):
            raise IOError('perf.script')
        text, _ = io.read(logdir + 'perf.script')
        if cfg.verbose:
            print_info(cfg, "Length of cpu_traces = %d" % text.count(b'\n'))
        if io.size(logdir + 'perf.script') > 0:
            p = hl.base_params(SRC_PERF, cfg.time_base, cfg.absolute_timestamp, t_offset)
            p['perf_has_cycles'] = 1 if cfg.perf_events.find('cycles') != -1 else 0
            p['mhz_xp'], p['mhz_fp'] = mhz_xp, mhz_fp
            eng.load(SRC_PERF, text, p)
            body = eng.csv(SRC_PERF, 1)
            io.write_csv(logdir + 'cputrace.csv', CSV_HEADER, body)
            mangled = eng.viz_mangled(ratio)                  # unique _Z... viz names
            t0 = time.perf_counter()
            pairs = [(m, demangle(m)) for m in mangled]
            tm.add('demangle (host)', t0)
            eng.set_demangled(pairs)
            kws = [flt.keyword for flt in cfg.cpu_filters]
            for kw in kws:
                _check_keyword(kw)
            if kws:                                           # filters on the viz frame (:1819-1827): matched in the same pass
                rec, nrec, _ = eng.json_series(SRC_PERF, ratio, kws)
            else:
                rec, nrec = eng.json(SRC_PERF, ratio, -1)
            cpu_series.append(SeriesOut('cpu_trace', 'CPU', 'DarkGray', rec, nrec))
            if kws:
                for i, flt in enumerate(cfg.cpu_filters):
                    rec, nrec = eng.json(SRC_PERF, ratio, i)
                    cpu_series.append(SeriesOut(flt.keyword, '[keyword]' + flt.keyword, flt.color, rec, nrec))
    except IOError:
        print_warning(cfg, 'no perf traces.')
    except Exception as e:
        # The reference wraps its whole perf section in a bare `except:` (:1783-1837): a sample it cannot convert (too few
        # fields, float() / int() ValueError) costs the CPU trace - this warning, no cputrace.csv, no cpu series - not the
        # run; the nvprof and tcpdump sections have no such net and raise.  Spellings only this codec refuses stay loud.
        if getattr(e, 'row_code', None) not in (1, 2, 3, 5):
            raise
        print_warning(cfg, 'no perf traces.')
        cpu_series = []

    if io.rank == 0:
        side.update(_strace_pystacks(cfg, eng, logdir, ratio))

    print_progress(cfg, "Export Overhead Dynamics JSON File of CPU, Network and GPU traces -- begin")
    # ---------------- series list (:1849-2102); sources off the hot path are empty here
    series = list(cpu_series)
    series.append(SeriesOut('blktrace_starting_block', 'BLKTRACE_STARTING_BLOCK', 'Green', *side.get('blktrace_starting_block', (None, 0))))
    series.append(SeriesOut('blktrace_time', 'BLKTRACE_TIME', 'DodgerBlue', *side.get('blktrace_time', (None, 0))))
    series.append(SeriesOut('vmstat_cs', 'VMSTAT_CS', 'Pink', *side.get('vmstat_cs', (None, 0))))
    series.append(SeriesOut('vmstat_bi', 'VMSTAT_BI', 'DarkOrange', *side.get('vmstat_bi', (None, 0))))
    series.append(SeriesOut('vmstat_bo', 'VMSTAT_BO', 'DarkOrchid', *side.get('vmstat_bo', (None, 0))))
    if cfg.enable_mpstat:
        series.append(SeriesOut('mpstat_usr', 'MPSTAT_USR', 'Cyan', *side.get('mpstat_usr', (None, 0))))
    if cfg.enable_diskstat:
        series.append(SeriesOut('diskstat', 'DISK_USAGE (MB/s)', 'GreenYellow', *side.get('diskstat', (None, 0))))
    if cfg.enable_vmstat:
        series.append(SeriesOut('vmstat_usr', 'CPU_USAGE_USR', 'Magenta', *side.get('vmstat_usr', (None, 0))))
        series.append(SeriesOut('vmstat_sys', 'CPU_USAGE_SYS', 'LightBlue', *side.get('vmstat_sys', (None, 0))))
        series.append(SeriesOut('vmstat_in', 'VMSTAT_IN', 'DarkMagenta', *side.get('vmstat_in', (None, 0))))
        series.append(SeriesOut('vmstat_cs', 'VMSTAT_CS', 'DarkOliveGreen', *side.get('vmstat_cs', (None, 0))))
    series.append(SeriesOut('strace', 'STRACE.', 'DarkSlateGray', *side.get('strace', (None, 0))))
    series.append(SeriesOut('pystacks', 'Python-stacks.', 'Tomato', *side.get('pystacks', (None, 0))))
    series.append(SeriesOut('nvsmi_mem', 'GPU_MEM_Util.', 'lightblue', *side.get('nvsmi_mem', (None, 0))))
    series.append(SeriesOut('nvsmi_sm', 'GPU_SM_Util.', 'red', *side.get('nvsmi_sm', (None, 0))))
    if getattr(cfg, 'enable_encode_decode', False):
        series.append(SeriesOut('nvsmi_enc', 'GPU_ENC_Util.', 'rgba(255, 215, 0, 0.8)', *side.get('nvsmi_enc', (None, 0))))
        series.append(SeriesOut('nvsmi_dec', 'GPU_DEC_Util.', 'rgba(218, 165, 32, 0.8)', *side.get('nvsmi_dec', (None, 0))))
    series.append(net_series)
    series.extend(net_filter_series)                             # :2046-2055
    series.append(SeriesOut('tx_bandwidth', 'tx Bandwidth', 'rgba(135,206,250,0.8)', *side.get('tx_bandwidth', (None, 0))))
    series.append(SeriesOut('rx_bandwidth', 'rx Bandwidth', 'rgba(25,25,112,0.8)', *side.get('rx_bandwidth', (None, 0))))
    series.append(gpu_series)
    series.extend(gpu_filter_series)
    if api_series is not None:
        series.append(api_series)
    t0 = time.perf_counter()
    io.finish(series, logdir + 'report.js')
    tm.add('write report.js', t0)
    print_progress(cfg, "Export Overhead Dynamics JSON File of CPU, Network and GPU traces -- end")
