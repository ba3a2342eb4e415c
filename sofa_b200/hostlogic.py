"""Host-side scalars that parameterise the device ingest (no per-record work here).

Follows sofa_preprocess.py: time_base :418-422, cpuinfo table :424-436,
perf_events :1545-1551, perf_timebase.txt :1766-1784, nvprof header/unit rows
:1385-1417, t_base :1419 (reference paths relative to /root/reference/bin).
"""
import os

SRC_PERF, SRC_GPU, SRC_NET = 0, 1, 2

GPU_COLS = ('Start', 'Duration', 'Size', 'Throughput', 'Context', 'Stream', 'Name', 'Src Ctx', 'Dst Ctx')


class PreprocessError(RuntimeError):
    pass


def read_time_base(logdir, cpu_time_offset):
    with open(logdir + 'sofa_time.txt') as f:
        return float(f.readlines()[0]) + cpu_time_offset


def read_cpuinfo(logdir):
    xp, fp = [0.0], [3000.0]
    try:
        with open(logdir + 'cpuinfo.txt') as f:
            for line in f.readlines():
                fields = line.split()
                xp.append(float(fields[0]))
                fp.append(float(fields[1]))
    except Exception:
        pass
    return xp, fp


def read_perf_events(logdir):
    with open(logdir + 'perf_events_used.txt', 'r') as f:
        lines = f.readlines()
    return lines[0] if lines else ''


def read_perf_timebase(logdir):
    """-> perf_timebase_unix - perf_timebase_uptime (the cuhello sync path needs nvprof's
    sqlite export and is not supported; perf_timebase.txt is the reference's fallback)."""
    uptime = unix = 0
    with open(logdir + 'perf_timebase.txt') as f:
        lines = f.readlines()
    if len(lines) <= 3 or lines[0].find('WARNING') != -1:
        return 0
    try:
        uptime = float(lines[-2].split()[2].split(':')[0])
        unix = float(lines[-1].split()[0])
    except Exception:
        pass
    return unix - uptime


def base_params(kind, time_base, absolute_timestamp, t_offset=0.0):
    return dict(kind=kind, time_base=float(time_base), t_offset=float(t_offset),
                absolute_timestamp=1 if absolute_timestamp else 0, perf_has_cycles=1,
                mhz_xp=[0.0], mhz_fp=[3000.0], ts_rescale=1.0, dt_rescale=1.0, payload_unit=1.0,
                n_cudaproc=1, idx=[-1] * 9, api_mode=0)


def gpu_header(head):
    """Parse the three header rows from the first bytes of gputrace.tmp.

    Returns (params-update dict, data_offset) or None when the file holds no kernel
    trace (reference prints a warning and skips, :1453-1456)."""
    text = head.decode('utf-8', errors='replace')
    rows = text.split('\n', 3)
    if len(rows) < 4 and not text.endswith('\n'):
        if len(rows) < 3:
            return None
    if len(rows) < 2 or rows[1].split(',')[0] != '"Start"':
        return None
    indices = rows[1].replace('"', '').replace('\n', '').split(',')
    units = rows[2].split(',')
    pu = {'GB': 1024.0 ** 3, 'MB': 1024.0 ** 2, 'KB': 1024.0, 'B': 1.0}
    if len(units) <= 11 or units[11] not in pu:
        raise PreprocessError("The payload unit in gputrace.tmp was not recognized!")
    scale = {'ms': 1.0e3, 'us': 1.0e6}
    idx = [indices.index(c) if c in indices else -1 for c in GPU_COLS]
    if idx[0] < 0 or idx[1] < 0 or idx[6] < 0:
        raise PreprocessError("gputrace.tmp header lacks Start/Duration/Name")
    data_off = len(rows[0].encode()) + len(rows[1].encode()) + len(rows[2].encode()) + 3
    first = rows[3] if len(rows) > 3 else ''
    if not first:
        raise PreprocessError("gputrace.tmp has a header but no records")   # records[0] IndexError
    t_base = float(first.split(',')[0])                      # :1419, not rescaled
    return dict(ts_rescale=scale.get(units[0], 1.0), dt_rescale=scale.get(units[1], 1.0),
                payload_unit=pu[units[11]], idx=idx, t_base=t_base), data_off


def nvsmi_local_offset(text, data_off):
    """Seconds the host's local zone is ahead of UTC for the stamps of nvsmi_query.txt.  The reference
    calls datetime.strptime(...).timestamp() on the naive stamps (sofa_preprocess.py:1032), i.e. reads
    them as local time; the device computes the civil-to-epoch part as UTC and subtracts this.  Looked
    up for the first and the last record; a change in between (DST switch) is refused."""
    import calendar
    import datetime
    body = bytes(memoryview(text)[data_off:])
    first = body[:body.find(b'\n')] if b'\n' in body else body
    tail = body.rstrip(b'\n')
    last = tail[tail.rfind(b'\n') + 1:]
    offs = set()
    for ln in (first, last):
        try:
            dt = datetime.datetime.strptime(ln.split(b',')[0].decode(), '%Y/%m/%d %H:%M:%S.%f')
        except (ValueError, UnicodeDecodeError):
            continue                                    # the device reports the malformed record
        offs.add(int(calendar.timegm(dt.timetuple()) - dt.replace(microsecond=0).timestamp()))
    if len(offs) > 1:
        raise PreprocessError("the local UTC offset changes within nvsmi_query.txt (DST switch): not supported")
    return offs.pop() if offs else 0


ROW_ERRORS = {1: "too few fields (IndexError in the reference)", 2: "float() ValueError",
              3: "int() ValueError", 4: "number spelling not supported by the device codec",
              5: "int('0x01'+ip, 16) ValueError", 6: "line longer than 65535 bytes / tile overlap",
              7: "integer out of range", 8: "carriage return in trace text",
              9: "too many short lines in one tile", 10: "value outside the text codec's range",
              11: "invalid UTF-8 in a name"}


def api_header(head):
    """Header rows of `nvprof --csv --print-api-trace` (sofa_preprocess.py:1487-1506)."""
    text = head.decode('utf-8', errors='replace')
    rows = text.split('\n', 3)
    if len(rows) < 3 or rows[1].split(',')[0] != '"Start"':
        return None
    indices = rows[1].replace('"', '').replace('\n', '').split(',')
    units = rows[2].split(',')
    scale = {'ms': 1.0e3, 'us': 1.0e6}
    idx = [-1] * 9
    for slot, col in ((0, 'Start'), (1, 'Duration'), (6, 'Name')):
        if col not in indices:
            raise PreprocessError("cuda_api_trace.tmp header lacks %s" % col)
        idx[slot] = indices.index(col)
    data_off = len(rows[0].encode()) + len(rows[1].encode()) + len(rows[2].encode()) + 3
    return dict(ts_rescale=scale.get(units[0], 1.0), dt_rescale=scale.get(units[1], 1.0), idx=idx, api_mode=1), data_off


def api_t_base(text, data_off, first_corid=1):
    """t_base of the API trace (sofa_preprocess.py:1510-1518): Start of the first record, replaced
    by the Start of the record whose 4th field equals first_corid when records have 4 fields."""
    body = memoryview(text)[data_off:]
    end = bytes(body[:4096]).find(b'\n')
    first = bytes(body[:end if end >= 0 else len(body)]).decode()
    if not first:
        raise PreprocessError("cuda_api_trace.tmp has a header but no records")
    t_base = float(first.split(',')[0])
    if len(first.split(',')) == 4:
        pos = 0
        raw = bytes(body)
        while pos < len(raw):
            e = raw.find(b'\n', pos)
            if e < 0:
                e = len(raw)
            rec = raw[pos:e + 1].decode()
            if int(rec.split(',')[3]) == first_corid:
                t_base = float(rec.split(',')[0])
                break
            pos = e + 1
    return t_base


def last_nvsmi_gpu(text, data_off):
    """gpu index of the last nvsmi_query.txt row: the value the reference's `cpuid` local is left with after the
    nvidia-smi query section (sofa_preprocess.py:1036, 1044) and that later lands in the strace / pystacks rows"""
    for ln in reversed(bytes(text[data_off:]).split(b'\n')):
        f = ln.split(b',')
        if len(f) >= 5:
            return int(f[2])
    return None


def last_dmon_gpu(text):
    """gpu index of the last accepted `nvidia-smi dmon` line (sofa_preprocess.py:1108-1120): what `cpuid` keeps"""
    for ln in reversed(bytes(text).split(b'\n')):
        if ln.find(b'gpu') != -1 or ln.find(b'Idx') != -1:
            continue
        f = ln.split()
        if len(f) >= 5:
            return int(f[0])
    return None
