"""CPU ORACLE for the SOFA preprocess/analyze hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import it.  The product
path (sofa_b200.sofa_preprocess / sofa_analyze) calls libsofa_b200.so and fails
loudly when the CUDA library is missing; it never falls back to this file.

It restates the reference's per-record algorithms in plain Python/numpy/pandas,
using the same library calls where the reference's *output bytes* are defined by a
library (pandas `to_csv(float_format='%.6f')`, `to_json`, `json.dump`, Python
`float()`, `np.log10`, `np.interp`).  Every function cites the reference lines it
follows (paths relative to /root/reference/bin).

Pinning: the reference has no golden vectors of its own.  This
oracle is pinned against outputs of the UNMODIFIED reference run in the build
container on synthetic logdirs: tests/golden/make_golden.py generates
tests/golden/<case>/{cputrace,gputrace,nettrace}.csv, report.js, comm.csv and the
gpu_profile/comm_profile feature table; tests/test_oracle_golden.py requires this
oracle to reproduce them byte for byte.
"""
import json
import math
import multiprocessing as mp
import os
import re
from functools import partial

import numpy as np
import pandas as pd

FIELDNAMES = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload",
              "bandwidth", "pkt_src", "pkt_dst", "pid", "tid", "name", "category"]

_CPU_COL = re.compile(r'\[\d+\]')


# --------------------------------------------------------------------------- a1
def parse_perf_sample(sample, absolute_timestamp, time_base, t_offset, mhz_xp, mhz_fp,
                      perf_events):
    """One `perf script` line -> 13-col row.  sofa_preprocess.py:110-154."""
    tok = sample.split()
    shift = 1 if _CPU_COL.match(tok[1]) is not None else 0      # :114 all-CPU mode
    t = float(tok[1 + shift].split(':')[0])                      # :115 / :120
    name = '[%s]' % tok[3 + shift].replace('-', '_') + tok[5 + shift] + tok[6 + shift]
    counts = float(tok[2 + shift])                               # :117 / :122
    event_raw = 1.0 * int("0x01" + tok[4 + shift], 16)           # :118 / :123
    if not absolute_timestamp:
        t = t - time_base                                        # :125-126
    t_begin = t + t_offset                                       # :128
    if len(mhz_xp) > 1:
        duration = counts / (np.interp(t_begin, mhz_xp, mhz_fp) * 1e6)   # :131-132
    else:
        duration = counts / (3000.0 * 1e6)                       # :134
    event = np.log10(event_raw)                                  # :136
    if perf_events.find('cycles') == -1:
        duration = np.log2(event_raw / 1e14)                     # :138-139
    ids = tok[0].split('/')
    return [t_begin, event, duration, -1, -1, 0, 0, -1, -1, int(ids[0]), int(ids[1]), name, 0]


# --------------------------------------------------------------------------- a3
def parse_net_packet(packet, absolute_timestamp, time_base, t_offset):
    """One `tcpdump -q -n -tt` line -> row, or [] (non-IP / unknown protocol).
    sofa_preprocess.py:156-201."""
    tok = packet.split()
    t = float(tok[0])
    if not absolute_timestamp:
        t = t - time_base
    t_begin = t + t_offset
    if tok[1] != 'IP':
        return []                                                # :167-168
    proto = tok[5]
    if proto == 'UDP,':
        payload = int(tok[7])
    elif proto == 'tcp':
        payload = int(tok[6])
    else:
        return []                                                # :175-176
    duration = float(payload / 128.0e6)
    src = 0
    dst = 0
    a = tok[2].split('.')
    b = tok[4].split('.')
    for i in range(4):                                           # :182-186 (np.int64 sums)
        src = src + int(a[i]) * np.power(1000, 3 - i)
        dst = dst + int(b[i]) * np.power(1000, 3 - i)
    return [t_begin, payload * 100 + 17, duration, -1, -1, payload, 128.0e6, src, dst, -1, -1,
            "network:%s:%d_to_%d_with_%d" % (proto, src, dst, payload), 0]


# --------------------------------------------------------------------------- a2
def parse_gpu_record(record, absolute_timestamp, time_base, indices, n_cudaproc, ts_rescale,
                     dt_rescale, payload_unit, t_offset):
    """One `nvprof --csv --print-gpu-trace` row -> row.  sofa_preprocess.py:249-341."""
    v = record.replace('"', '').split(',')                       # :258 (no strip: Name keeps '\n')
    name = v[indices.index('Name')]
    t = float(v[indices.index('Start')]) / ts_rescale + t_offset
    if not absolute_timestamp:
        t = t - time_base
    duration = float(v[indices.index('Duration')]) / dt_rescale
    try:
        payload = int(float(v[indices.index('Size')]) * payload_unit)
    except BaseException:
        payload = 0
    try:
        bandwidth = float(v[indices.index('Throughput')])
    except BaseException:
        bandwidth = 0
    try:
        dev = int(float(v[indices.index('Context')]))
    except BaseException:
        dev = -1
    try:
        tid = int(float(v[indices.index('Stream')]))
    except BaseException:
        tid = -1
    src = dst = kind = 0
    if name.find('HtoD') != -1:                                  # :293-323 tagging order
        kind, src, dst = 1, 0, dev
        name = "CUDA_COPY_H2D_%dB" % payload
    elif name.find('DtoH') != -1:
        kind, src, dst = 2, dev, 0
        name = "CUDA_COPY_D2H_%dB" % payload
    elif name.find('DtoD') != -1:
        kind, src, dst = 8, dev, dev
        name = "CUDA_COPY_D2D_%dB" % payload
    elif name.find('PtoP') != -1:
        kind = 10
        try:
            src = int(v[indices.index('Src Ctx')])
        except BaseException:
            src = 0
        try:
            dst = int(v[indices.index('Dst Ctx')])
        except BaseException:
            dst = 0
        name = "[CUDA_COPY_P2P]from_gpu%d_to_gpu%d_%dB" % (src, dst, payload)
    if dev != -1:
        name = '[gpu%d]' % dev + name                            # :325-326
    return [t, payload * 100 + 17, duration, dev, kind, payload, bandwidth, src, dst,
            n_cudaproc, tid, name, 0]


def parse_api_record(record, absolute_timestamp, time_base, indices, n_cudaproc, ts_rescale, dt_rescale,
                     t_offset):
    """One `nvprof --csv --print-api-trace` row -> row.  sofa_preprocess.py:203-247."""
    v = record.replace('"', '').split(',')
    name = '[CUDA_API]' + v[indices.index('Name')]
    t = float(v[indices.index('Start')]) / ts_rescale + t_offset
    if not absolute_timestamp:
        t = t - time_base
    duration = float(v[indices.index('Duration')]) / dt_rescale
    return [t, 0 * 100 + 17, duration, -1, 0, 0, 0, 0, 0, n_cudaproc, -1, name, 0]


def gpu_units(unit_row):
    """Unit row -> (ts_rescale, dt_rescale, payload_unit).  sofa_preprocess.py:1392-1415."""
    u = unit_row.split(',')
    pu = {'GB': np.power(1024, 3), 'MB': np.power(1024, 2), 'KB': np.power(1024, 1), 'B': 1}
    if u[11] not in pu:
        raise ValueError("The payload unit in gputrace.tmp was not recognized!")
    scale = {'ms': 1.0e3, 'us': 1.0e6}
    return scale.get(u[0], 1.0), scale.get(u[1], 1.0), pu[u[11]]


# --------------------------------------------------------------------------- a6 / a7 / a9
def downsample(rows, ratio):
    """sofa_preprocess.py:51-57."""
    return [rows[i] for i in range(len(rows)) if i % ratio == 0]


def frame(rows):
    df = pd.DataFrame(rows)
    df.columns = FIELDNAMES
    return df


def write_trace_csv(df, path):
    """sofa_preprocess.py:1206-1211 / 1439-1444 / 1802-1807."""
    df.to_csv(path, mode='w', header=True, index=False, float_format='%.6f')


def viz_name(x, demangle):
    """sofa_preprocess.py:1811-1816: text between the first ']' and the first '+',
    cut at '@', demangled."""
    return demangle(str(x[x.find(']') + 1: x.find('+')].split('@')[0]))


# --------------------------------------------------------------------------- a10
class Series:
    """One report.js series (sofa_models.py:1-7)."""
    def __init__(self, name, title, color, data, x_field='timestamp', y_field='duration'):
        self.name, self.title, self.color, self.data = name, title, color, data
        self.x_field, self.y_field = x_field, y_field


def write_report_js(series, path):
    """sofa_preprocess.py:343-374."""
    if len(series) == 0:
        return
    with open(path, 'w') as f:
        for s in series:
            if len(s.data) > 0:
                f.write(s.name + " = ")
                d = s.data.rename(columns={s.x_field: 'x', s.y_field: 'y'})
                json.dump({"name": s.title, "color": s.color,
                           "data": json.loads(d.to_json(orient='records'))}, f)
            f.write("\n\n")
        f.write("sofa_traces = [ ")
        for s in series:
            if len(s.data) > 0:
                f.write(s.name + ",")
        f.write(" ]")


# --------------------------------------------------------------------------- N1: recorder side files
def _to_csv_list(rows, path, mode):
    """list_to_csv_and_traces (sofa_preprocess.py:89-102) on an already sliced list."""
    if len(rows) == 0:
        return []
    df = frame(rows)
    df.to_csv(path, mode=mode, header=(mode == 'w'), index=False, float_format='%.6f')
    return df


def preprocess_mpstat(logdir, absolute_timestamp, time_base):
    """mpstat.txt -> usr_sys.csv + mpstat.csv.  sofa_preprocess.py:482-546."""
    mpstat = np.genfromtxt(logdir + 'mpstat.txt', delimiter=',', invalid_raise=False)      # :484
    mpstat = mpstat[1:]                                                                    # header row (all nan)
    n_cores = int(mpstat[:, 1].max() + 1)                                                  # :489
    stride = n_cores + 1
    rows, usr_sys = [], []
    for i in range(len(mpstat)):
        if i <= stride:                                                                    # :495
            continue
        core = mpstat[i, 1]
        d = mpstat[i, :] - mpstat[i - stride, :]
        total = d[2] + d[4] + d[5] + d[6] + d[7]
        if total == 0:
            continue
        usr = d[2] * 100 / float(total)
        sys_ = d[4] * 100 / float(total)
        idl = d[5] * 100 / float(total)
        iow = d[6] * 100 / float(total)
        irq = d[7] * 100 / float(total)
        cpu_time = (total - d[5]) * 0.01
        t_begin = mpstat[i, 0]
        if not absolute_timestamp:
            t_begin = t_begin - time_base
        if mpstat[i, 1] != -1:
            name = 'mpstat_core%d (usr|sys|idl|iow|irq): |%3d|%3d|%3d|%3d|%3d|' % (core, usr, sys_, idl, iow, irq)
            rows.append([t_begin, -1, cpu_time, core, -1, -1, -1, -1, -1, -1, -1, name, 0])
        else:
            usr_sys.append((t_begin, usr, sys_))
    pd.DataFrame(usr_sys, columns=['time', 'usr', 'sys']).to_csv(logdir + 'usr_sys.csv', index=False)   # :544-545
    return _to_csv_list(rows, logdir + 'mpstat.csv', 'w')


def preprocess_nvsmi_query(logdir, absolute_timestamp, time_base, nvsmi_time_zone=0):
    """nvsmi_query.txt -> nvsmi_trace.csv (sm rows, then mem rows).  sofa_preprocess.py:1013-1084."""
    import datetime
    with open(logdir + 'nvsmi_query.txt') as f:
        next(f)
        lines = f.readlines()
    for line in lines:
        if line.find('failed') != -1 or line.find('Failed') != -1:
            return [], []
    sm, mem = [], []
    for line in lines:
        fields = line.split(',')
        nv_time = datetime.datetime.strptime(fields[0], '%Y/%m/%d %H:%M:%S.%f').timestamp() + nvsmi_time_zone * 3600
        gid, v_sm, v_mem = int(fields[2]), int(fields[3][:-2]), int(fields[4][:-2])
        t_begin = nv_time if absolute_timestamp else nv_time - time_base
        sm.append([t_begin, 0, v_sm, gid, -1, -1, -1, -1, -1, -1, -1, "GPUID_sm=%d_%d" % (gid, v_sm), gid])
        mem.append([t_begin, 1, v_mem, gid, -1, -1, -1, -1, -1, -1, -1, "GPUID_mem=%d_%d" % (gid, v_mem), gid])
    if not sm:
        return [], []
    return _to_csv_list(sm, logdir + 'nvsmi_trace.csv', 'w'), _to_csv_list(mem, logdir + 'nvsmi_trace.csv', 'a')


def preprocess_netstat(logdir, absolute_timestamp, time_base):
    """netstat.txt -> netbandwidth.csv + netstat.csv (tx rows, then rx rows).  sofa_preprocess.py:1235-1337."""
    with open(logdir + 'netstat.txt') as f:
        lines = f.readlines()
    if not lines:
        return [], []
    tmp_time = float(lines[0].split(',')[0])
    tmp_tx, tmp_rx = int(lines[0].split(',')[1]), int(lines[0].split(',')[2])
    tx_list, rx_list, bw = [], [], []
    for line in lines[1:]:
        time = float(line.split(',')[0])
        tx, rx = int(line.split(',')[1]), int(line.split(',')[2])
        tx_bw = (tx - tmp_tx) / (time - tmp_time)
        rx_bw = (rx - tmp_rx) / (time - tmp_time)
        t_begin = time if absolute_timestamp else time - time_base
        tx_list.append([t_begin, 0, -1, -1, -1, -1, tx_bw, -1, -1, -1, -1, "network_bandwidth_tx(bytes):%d" % tx_bw, 0])
        rx_list.append([t_begin, 1, -1, -1, -1, -1, rx_bw, -1, -1, -1, -1, "network_bandwidth_rx(bytes):%d" % rx_bw, 0])
        bw.append([t_begin, tx_bw, rx_bw])
        tmp_time, tmp_tx, tmp_rx = time, tx, rx
    # the reference grows this frame by pd.concat one row at a time (:1315-1316); the bytes are those of one frame
    pd.DataFrame(bw, columns=['time', 'tx_bandwidth', 'rx_bandwidth']).to_csv(logdir + 'netbandwidth.csv', header=True, index=False)
    tx_traces = pd.DataFrame(tx_list, columns=FIELDNAMES)
    tx_traces.to_csv(logdir + 'netstat.csv', mode='w', header=True, index=False, float_format='%.6f')
    rx_traces = pd.DataFrame(rx_list, columns=FIELDNAMES)
    rx_traces.to_csv(logdir + 'netstat.csv', mode='a', header=False, index=False, float_format='%.6f')
    return tx_traces, rx_traces


def preprocess_diskstat(logdir, absolute_timestamp, time_base, diskstat_filters=None):
    """diskstat.txt -> diskstat.csv, diskstat_vector.csv, diskstat_vector_ui.csv.  sofa_preprocess.py:548-673.
    Pinned against the reference (tests/golden/disk_small); the device version (side2.cuh) is checked against it."""
    with open(logdir + 'diskstat.txt') as f:
        diskstats = f.readlines()
    tmp_list = [d[:-1].split(',') for d in diskstats]
    devs = list(set(map(lambda x: x[1], tmp_list)))
    n_dev = len(devs)
    rows, vectors = [], []
    for i in range(len(diskstats)):
        if i < n_dev:
            continue
        m = diskstats[i][:-1].split(',')
        dev = m[1]
        m_last = diskstats[i - n_dev][:-1].split(',')
        secsize = 512
        try:
            with open('/sys/block/' + dev + '/queue/hw_sector_size') as f:
                secsize = int(re.match(r"\d+", f.readline()).group())
        except Exception:
            pass
        d_read = (int(m[2]) - int(m_last[2])) * secsize
        d_write = (int(m[3]) - int(m_last[3])) * secsize
        d_disk_total = d_read + d_write
        t_begin = float(m[0])
        if not absolute_timestamp:
            t_begin = t_begin - time_base
        d_duration = float(m[0]) - float(m_last[0])
        d_throughput = round((d_disk_total / d_duration) / float(1024 ** 2), 2)
        read_bandwidth = d_read / d_duration
        write_bandwidth = d_write / d_duration
        total_bandwidth = d_disk_total / d_duration
        r_io = int(m[4]) - int(m_last[4])
        w_io = int(m[5]) - int(m_last[5])
        r_iops = int(r_io / d_duration)
        w_iops = int(w_io / d_duration)
        iops = int((r_io + w_io) / d_duration)
        r_time, w_time, await_time = 0, 0, 0
        if r_io > 0:
            r_time = (int(m[6]) - int(m_last[6])) / r_io
        if w_io > 0:
            w_time = (int(m[7]) - int(m_last[7])) / w_io
        if (r_io + w_io) > 0:
            await_time = ((int(m[6]) - int(m_last[6])) + (int(m[7]) - int(m_last[7]))) / (r_io + w_io)
        vectors.append((t_begin, m[1], r_iops, w_iops, iops, r_time, w_time, await_time, read_bandwidth, write_bandwidth,
                        total_bandwidth))
        info = 'diskstat_dev:%s (read|write): |%3d|%3d| bytes' % (m[1], d_read, d_write)
        rows.append([t_begin, -1, d_duration, m[1], -1, d_disk_total, d_throughput, -1, -1, -1, -1, info, 0])
    traces = pd.DataFrame(rows, columns=FIELDNAMES)
    table = pd.DataFrame(vectors, columns=['time', 'dev', 'r_iops', 'w_iops', 'iops', 'r_time', 'w_time', 'await_time',
                                           'd_read', 'd_write', 'd_disk_total'])
    for dev in devs:
        if table[table['dev'] == dev]['d_disk_total'].max() == 0:
            table = table[table.dev != dev]
            traces = traces[traces.deviceId != dev]
    traces.to_csv(logdir + 'diskstat.csv', index=False)
    table.to_csv(logdir + 'diskstat_vector.csv', index=False)
    if diskstat_filters:
        table = table[table['dev'] == diskstat_filters]
    table.to_csv(logdir + 'diskstat_vector_ui.csv', index=False)
    return traces


def preprocess_blktrace(logdir):
    """blktrace.txt -> blktrace.csv (D rows, then completion latencies).  sofa_preprocess.py:688-781.
    Returns (starting-block frame, latency frame, cpuid left behind for later sections)."""
    with open(logdir + 'blktrace.txt') as f:
        lines = f.readlines()
    d_dict, d_list, c_list = {}, [], []
    cpuid = None
    for line in lines:
        if len(line) > 50 and "Read" not in line and "CPU" not in line and "IO unplugs" not in line:
            fields = line.split()
            cpu = fields[1]
            ts = float(fields[3])
            pid = fields[4]
            event = fields[5]
            try:
                start_block = int(fields[7])
            except Exception:
                start_block = 0
            cpuid = cpu
            trace = [ts, event, start_block, cpu, -1, -1, -1, -1, -1, pid, pid, 'starting_block=' + str(start_block), cpu]
            if event == 'D':                       # the reference writes `'D' is event` (interned one-character strings)
                d_list.append(trace)
                d_dict[start_block] = trace
            if event == 'C':
                if start_block not in d_dict:
                    continue
                d_event = d_dict[start_block]
                lat = float(ts - d_event[0])
                c_list.append([d_event[0], event, float(lat), cpu, -1, -1, -1, -1, -1, pid, pid, 'latency=%0.6f' % float(lat), cpu])
    return _to_csv_list(d_list, logdir + 'blktrace.csv', 'w'), _to_csv_list(c_list, logdir + 'blktrace.csv', 'a'), cpuid


def preprocess_pystacks(logdir, absolute_timestamp, time_base, cpuid):
    """pystacks.txt -> pystacks.csv.  sofa_preprocess.py:1706-1761.  list_to_csv_and_traces drops the first element
    of the list it is given; this list has no placeholder row, so the first stack is lost - as in the reference."""
    import itertools
    ret = {}
    with open(logdir + 'pystacks.txt', 'r') as f:
        for ts, fs in itertools.zip_longest(*[f] * 2):
            fs = fs.replace('\n', '').replace(';', '<br>')
            if fs.find('idle') != -1:
                continue
            ret[int(ts) / 10 ** 6] = fs
    duration = {}
    prev = None
    for k in ret:
        if prev is None:
            prev = k
            continue
        duration[prev] = k - prev
        prev = k
    del ret[max(ret.keys())]
    rows = []
    for key, info in ret.items():
        if cpuid is None:
            raise NameError("name 'cpuid' is not defined")
        t_begin = key if absolute_timestamp else key - time_base
        rows.append([t_begin, -1, float(duration[key]), -1, -1, -1, -1, -1, -1, -1, -1, info, cpuid])
    return _to_csv_list(rows[1:], logdir + 'pystacks.csv', 'w') if rows else []


STRACE_FILTER_KEYS = ('resume', 'nanosleep', 'clock_gettime', 'brk', 'stat', 'close', 'exited', 'access', 'lseek', 'getrusage', 'exited')


def preprocess_strace(logdir, absolute_timestamp, time_base, plot_ratio, pid, strace_min_time, cpuid):
    """strace.txt -> strace.csv.  sofa_preprocess.py:1618-1704.  `cpuid` is whatever an earlier section left in
    that variable (the reference never sets it here)."""
    with open(logdir + 'strace.txt') as f:
        lines = f.readlines()
    rows = []
    if len(lines) > 1:
        for i in range(len(lines)):
            if i % plot_ratio > 0:
                continue
            if any(lines[i].find(key) != -1 for key in STRACE_FILTER_KEYS):
                continue
            fields = lines[i].split()
            if fields[0].find('pid') != -1:
                tid = int(fields[1].split(']')[0])
                t_begin = float(fields[2])
                info = ''.join(fields[3:-3])
            else:
                tid = pid
                t_begin = float(fields[1])
                info = ''.join(fields[1:-3])
            if not absolute_timestamp:
                t_begin = t_begin - time_base
            try:
                duration = float(fields[-1].split('<')[1].split('>')[0])
            except Exception:
                duration = 0
            if duration < strace_min_time:
                continue
            if cpuid is None:
                raise NameError("name 'cpuid' is not defined")
            rows.append([t_begin, -1, duration, -1, -1, -1, -1, -1, -1, pid, tid, info, cpuid])
    return _to_csv_list(rows, logdir + 'strace.csv', 'w')


def preprocess_vmstat(logdir, absolute_timestamp, time_base):
    """vmstat.txt -> vmstat.csv (bi, bo, in, cs, wa, st, usr, sys lists).  sofa_preprocess.py:787-1008.
    Returns the eight frames in that order."""
    with open(logdir + 'vmstat.txt') as f:
        lines = f.readlines()
    lists = [[] for _ in range(8)]
    cols = (8, 9, 10, 11, 15, 16, 12, 13)
    keys = ('r', 'b', 'sw', 'fr', 'bu', 'ca', 'si', 'so', 'bi', 'bo', 'in', 'cs', 'usr', 'sys', 'idl', 'wa', 'st')
    t = 0
    for line in lines:
        if line.find('procs') == -1 and line.find('swpd') == -1:
            fields = line.split()
            if len(fields) < 17:
                continue
            vm = [float(fields[k]) + 1e-5 for k in range(17)]
            t_begin = t + time_base if absolute_timestamp else t
            info = '|'.join('%s=%s' % (k, str(int(v))) for k, v in zip(keys, vm))
            for lst, c in zip(lists, cols):
                lst.append([t_begin, -1, vm[c], -1, -1, -1, -1, -1, -1, -1, -1, info, -1])
            t = t + 1
    return [_to_csv_list(lst, logdir + 'vmstat.csv', 'w' if i == 0 else 'a') for i, lst in enumerate(lists)]


def preprocess_nvsmi_dmon(logdir, absolute_timestamp, time_base):
    """nvsmi.txt (`nvidia-smi dmon -s u`) -> enc / dec rows appended to nvsmi_trace.csv.
    sofa_preprocess.py:1092-1183.  Returns (enc frame, dec frame, nvsmi_data flag)."""
    with open(logdir + 'nvsmi.txt') as f:
        lines = f.readlines()
    for line in lines:
        if line.find('failed') != -1 or line.find('Failed') != -1:
            return [], [], False
    enc_list, dec_list = [], []
    t = 0
    for line in lines:
        if line.find('gpu') == -1 and line.find('Idx') == -1:
            fields = line.split()
            if len(fields) < 5:
                continue
            gid = int(fields[0])
            enc = int(0) if fields[3] == '-' else int(fields[3])
            dec = int(0) if fields[4] == '-' else int(fields[4])
            t_begin = t + time_base if absolute_timestamp else t
            if t > 3:
                enc_list.append([t_begin, 2, enc, gid, -1, -1, -1, -1, -1, -1, -1, "GPUID_enc=%d_%d" % (gid, enc), gid])
                dec_list.append([t_begin, 3, dec, gid, -1, -1, -1, -1, -1, -1, -1, "GPUID_dec=%d_%d" % (gid, dec), gid])
            if gid == 0:
                t = t + 1
    if not enc_list:
        return [], [], False
    return _to_csv_list(enc_list, logdir + 'nvsmi_trace.csv', 'a'), _to_csv_list(dec_list, logdir + 'nvsmi_trace.csv', 'a'), True


# --------------------------------------------------------------------------- driver
class OracleConfig:
    """The cfg attributes the hot path reads (bin/sofa:159-312 defaults)."""
    def __init__(self, logdir, **kw):
        self.logdir = logdir if logdir.endswith('/') else logdir + '/'
        self.absolute_timestamp = False
        self.cpu_time_offset = 0.0
        self.plot_ratio = 1
        self.cpu_filters = [("idle", "black")]
        self.gpu_filters = [("CUDA_COPY_H2D", "Red"), ("CUDA_COPY_D2H", "Peru"),
                            ("CUDA_COPY_P2P", "Purple"), ("_fw_", "royalblue"),
                            ("_bw_", "darkviolet"), ("AllReduceKernel", "indigo")]
        self.enable_mpstat = True
        self.enable_diskstat = True
        self.enable_vmstat = False
        self.enable_encode_decode = False
        self.cuda_api_tracing = False
        self.verbose = False
        self.nprocs = None
        self.nvsmi_time_zone = 0
        for k, v in kw.items():
            if k == 'cpu_filters' and isinstance(v, str):
                v = [tuple(p.split(':')) for p in v.split(',')]
            setattr(self, k, v)
        if self.cuda_api_tracing and ("CUDA_API", "DarkSlateGray") not in self.gpu_filters:
            self.gpu_filters.insert(3, ("CUDA_API", "DarkSlateGray"))            # bin/sofa:276-277


def _pmap(fn, items, nprocs):
    if nprocs == 1 or len(items) < 64:
        return [fn(x) for x in items]
    with mp.Pool(processes=nprocs or mp.cpu_count()) as pool:    # :1196, :1420, :1791
        return pool.map(fn, items)


def read_time_bases(cfg):
    """time_base (:418-422), cpuinfo table (:424-436), perf_events (:1545-1551),
    perf_timebase.txt (:1766-1784)."""
    logdir = cfg.logdir
    with open(logdir + 'sofa_time.txt') as f:
        time_base = float(f.readlines()[0]) + cfg.cpu_time_offset
    xp, fp = [0.0], [3000.0]
    try:
        with open(logdir + 'cpuinfo.txt') as f:
            for line in f.readlines():
                t = line.split()
                xp.append(float(t[0]))
                fp.append(float(t[1]))
    except Exception:
        pass
    with open(logdir + 'perf_events_used.txt') as f:
        lines = f.readlines()
        perf_events = lines[0] if lines else ''
    uptime = unix = 0
    with open(logdir + 'perf_timebase.txt') as f:
        lines = f.readlines()
        if len(lines) <= 3 or lines[0].find('WARNING') != -1:
            uptime = unix = 0
        else:
            try:
                uptime = float(lines[-2].split()[2].split(':')[0])
                unix = float(lines[-1].split()[0])
            except Exception:
                pass
    return time_base, xp, fp, perf_events, unix - uptime


def preprocess(cfg, demangle=lambda s: s, emit_json=True):
    """perf / nvprof-gpu / tcpdump sources of sofa_preprocess (sofa_preprocess.py:377-2106),
    reading the pre-dumped text (`gputrace.tmp`, `net.tmp`) instead of shelling out.
    Returns the list of Series written to report.js."""
    logdir = cfg.logdir
    time_base, xp, fp, perf_events, cpu_t_offset = read_time_bases(cfg)
    net_traces = []
    filtered_net_groups = []
    cpu_viz = []
    filtered_cpu = []
    gpu_viz = []
    filtered_gpu = []

    # ---- recorder side files  :482-546, :1013-1084, :1235-1337
    mpstat_traces, nvsmi_sm, nvsmi_mem, tx_traces, rx_traces = [], [], [], [], []
    if os.path.isfile(logdir + 'mpstat.txt'):
        mpstat_traces = preprocess_mpstat(logdir, cfg.absolute_timestamp, time_base)
    vm = [[] for _ in range(8)]
    nvsmi_enc, nvsmi_dec = [], []
    diskstat_traces = []
    blk_d_traces, blk_traces, strace_traces = [], [], []
    cpuid = None                                    # a local of the reference that leaks from section to section
    if os.path.isfile(logdir + 'blktrace.txt'):
        with open(logdir + 'blktrace.txt') as f:
            if len(f.readlines()) > 0:
                blk_d_traces, blk_traces, c_ = preprocess_blktrace(logdir)
                cpuid = c_ if c_ is not None else cpuid
    if os.path.isfile(logdir + 'diskstat.txt'):
        diskstat_traces = preprocess_diskstat(logdir, cfg.absolute_timestamp, time_base, getattr(cfg, 'diskstat_filters', None))
    if os.path.isfile(logdir + 'vmstat.txt'):
        with open(logdir + 'vmstat.txt') as f:
            if len(f.readlines()) > 0:                                    # :791
                vm = preprocess_vmstat(logdir, cfg.absolute_timestamp, time_base)
                if len(vm[0]) > 0:
                    cpuid = -1
    if os.path.isfile(logdir + 'nvsmi_query.txt'):
        nvsmi_sm, nvsmi_mem = preprocess_nvsmi_query(logdir, cfg.absolute_timestamp, time_base, cfg.nvsmi_time_zone)
        if len(nvsmi_sm) > 0:
            cpuid = int(nvsmi_sm['deviceId'].iloc[-1])
    if os.path.isfile(logdir + 'nvsmi.txt'):
        nvsmi_enc, nvsmi_dec, cfg.nvsmi_data = preprocess_nvsmi_dmon(logdir, cfg.absolute_timestamp, time_base)
        with open(logdir + 'nvsmi.txt') as f:
            dm = [ln.split() for ln in f.readlines() if ln.find('gpu') == -1 and ln.find('Idx') == -1]
        dm = [x for x in dm if len(x) >= 5]
        if dm and not any(('failed' in ' '.join(x) or 'Failed' in ' '.join(x)) for x in dm):
            cpuid = int(dm[-1][0])
    if os.path.isfile(logdir + 'netstat.txt'):
        tx_traces, rx_traces = preprocess_netstat(logdir, cfg.absolute_timestamp, time_base)

    # ---- network  :1187-1232  (nettrace.csv holds the DOWN-SAMPLED rows)
    if os.path.isfile(logdir + 'sofa.pcap') and os.path.isfile(logdir + 'net.tmp'):
        with open(logdir + 'net.tmp') as f:
            packets = f.readlines()
        if packets:
            res = _pmap(partial(parse_net_packet, absolute_timestamp=cfg.absolute_timestamp,
                                time_base=time_base, t_offset=0), packets, cfg.nprocs)
            net_traces = frame(downsample(res, cfg.plot_ratio))
            write_trace_csv(net_traces, logdir + 'nettrace.csv')

    # ---- GPU  :1343-1457 (single nvvp; t_glb_gpu_base = 0 when sqlite tables are unreadable)
    if os.path.isfile(logdir + 'gputrace.tmp'):
        with open(logdir + 'gputrace.tmp') as f:
            records = f.readlines()
        if len(records) > 0 and records[1].split(',')[0] == '"Start"':
            indices = records[1].replace('"', '').replace('\n', '').split(',')
            ts_rescale, dt_rescale, payload_unit = gpu_units(records[2])
            records = records[3:]
            t_base = float(records[0].split(',')[0])             # :1419 (un-rescaled)
            res = _pmap(partial(parse_gpu_record, absolute_timestamp=cfg.absolute_timestamp,
                                time_base=time_base, indices=indices, n_cudaproc=1,
                                ts_rescale=ts_rescale, dt_rescale=dt_rescale,
                                payload_unit=payload_unit, t_offset=0 - t_base),
                        records, cfg.nprocs)
            gpu_traces = frame(res)
            gpu_viz = frame(downsample(res, cfg.plot_ratio))
            write_trace_csv(gpu_traces, logdir + 'gputrace.csv')
            for kw, color in cfg.gpu_filters:                    # :1446-1452
                filtered_gpu.append((kw, color, gpu_traces[gpu_traces['name'].str.contains(kw)]))

    # ---- CUDA API trace  :1459-1543 (only with --cuda_api_tracing)
    api_viz = []
    if cfg.cuda_api_tracing and os.path.isfile(logdir + 'cuda_api_trace.tmp'):
        with open(logdir + 'cuda_api_trace.tmp') as f:
            records = f.readlines()
        if len(records) > 0 and records[1].split(',')[0] == '"Start"':
            indices = records[1].replace('"', '').replace('\n', '').split(',')
            scale = {'ms': 1.0e3, 'us': 1.0e6}
            ts_rescale = scale.get(records[2].split(',')[0], 1.0)
            dt_rescale = scale.get(records[2].split(',')[1], 1.0)
            records = records[3:]
            t_base = float(records[0].split(',')[0])
            if len(records[0].split(',')) == 4:                  # :1511-1517, first_corid = 1
                for record in records:
                    if int(record.split(',')[3]) == 1:
                        t_base = float(record.split(',')[0])
                        break
            res = _pmap(partial(parse_api_record, absolute_timestamp=cfg.absolute_timestamp,
                                time_base=time_base, indices=indices, n_cudaproc=1, ts_rescale=ts_rescale,
                                dt_rescale=dt_rescale, t_offset=0 - t_base), records, cfg.nprocs)
            write_trace_csv(frame(res), logdir + 'cuda_api_trace.csv')
            api_viz = frame(downsample(res, cfg.plot_ratio))

    # ---- CPU  :1783-1838.  The reference wraps this whole section in a bare `except:` (:1836-1837): a sample that cannot be
    # converted costs the CPU trace ("no perf traces."), not the run - unlike the nvprof and tcpdump sections, which raise.
    if os.path.isfile(logdir + 'perf.script'):
        viz_assigned = False
        try:
            with open(logdir + 'perf.script') as f:
                samples = f.readlines()
            if len(samples) > 0:
                res = _pmap(partial(parse_perf_sample, absolute_timestamp=cfg.absolute_timestamp,
                                    time_base=time_base, t_offset=cpu_t_offset, mhz_xp=xp, mhz_fp=fp,
                                    perf_events=perf_events), samples, cfg.nprocs)
                cpu_traces = frame(res)
                write_trace_csv(cpu_traces, logdir + 'cputrace.csv')
                cpu_viz = frame(downsample(res, cfg.plot_ratio))
                viz_assigned = True
                cpu_viz['name'] = cpu_viz['name'].apply(lambda x: viz_name(x, demangle))
                for kw, color in cfg.cpu_filters:                # :1819-1827
                    filtered_cpu.append((kw, color, cpu_viz[cpu_viz['name'].str.contains(kw)]))
        except Exception:
            print('[WARNING] no perf traces.')
            if viz_assigned:
                # past the point where cpu_traces_viz holds rows (:1808), e.g. cxxfilt.InvalidName: the swallowed exception
                # leaves `filtered_groups` unassigned and the series assembly (:1861) dies on it
                raise UnboundLocalError("cannot access local variable 'filtered_groups' where it is not associated with a value")

    # ---- strace  :1618-1704 (after the trace sources in the reference; it only needs `cpuid` from above)
    if os.path.isfile(logdir + 'strace.txt'):
        pid = 0
        with open(logdir + 'misc.txt') as f:
            ml = f.readlines()
            if len(ml) == 4:
                pid = int(ml[3].split()[1])
        strace_traces = preprocess_strace(logdir, cfg.absolute_timestamp, time_base, cfg.plot_ratio, pid,
                                          getattr(cfg, 'strace_min_time', 1e-6), cpuid)

    pystacks_traces = []
    if os.path.isfile(logdir + 'pystacks.txt'):                              # :1730-1761
        pystacks_traces = preprocess_pystacks(logdir, cfg.absolute_timestamp, time_base, cpuid)

    # ---- series list  :1849-2102 (sources outside the hot path are empty placeholders)
    series = []
    if len(cpu_viz) > 0:
        series.append(Series('cpu_trace', 'CPU', 'DarkGray', cpu_viz))
        for kw, color, grp in filtered_cpu:
            series.append(Series(kw, '[keyword]' + kw, color, grp.copy()))
    vm_bi, vm_bo, vm_in, vm_cs, vm_wa, vm_st, vm_usr, vm_sys = vm
    for nm, title, color, data in (('blktrace_starting_block', 'BLKTRACE_STARTING_BLOCK', 'Green', blk_d_traces),
                                   ('blktrace_time', 'BLKTRACE_TIME', 'DodgerBlue', blk_traces),
                                   ('vmstat_cs', 'VMSTAT_CS', 'Pink', vm_cs),
                                   ('vmstat_bi', 'VMSTAT_BI', 'DarkOrange', vm_bi),
                                   ('vmstat_bo', 'VMSTAT_BO', 'DarkOrchid', vm_bo)):
        series.append(Series(nm, title, color, data))
    if cfg.enable_mpstat:
        series.append(Series('mpstat_usr', 'MPSTAT_USR', 'Cyan', mpstat_traces))
    if cfg.enable_diskstat:
        series.append(Series('diskstat', 'DISK_USAGE (MB/s)', 'GreenYellow', diskstat_traces, y_field='bandwidth'))
    if cfg.enable_vmstat:                                                   # :1964-1999
        series.append(Series('vmstat_usr', 'CPU_USAGE_USR', 'Magenta', vm_usr))
        series.append(Series('vmstat_sys', 'CPU_USAGE_SYS', 'LightBlue', vm_sys))
        series.append(Series('vmstat_in', 'VMSTAT_IN', 'DarkMagenta', vm_in))
        series.append(Series('vmstat_cs', 'VMSTAT_CS', 'DarkOliveGreen', vm_cs))
    series.append(Series('strace', 'STRACE.', 'DarkSlateGray', strace_traces))
    series.append(Series('pystacks', 'Python-stacks.', 'Tomato', pystacks_traces))
    series.append(Series('nvsmi_mem', 'GPU_MEM_Util.', 'lightblue', nvsmi_mem))
    series.append(Series('nvsmi_sm', 'GPU_SM_Util.', 'red', nvsmi_sm))
    if cfg.enable_encode_decode:                                            # :2033-2050
        series.append(Series('nvsmi_enc', 'GPU_ENC_Util.', 'rgba(255, 215, 0, 0.8)', nvsmi_enc))
        series.append(Series('nvsmi_dec', 'GPU_DEC_Util.', 'rgba(218, 165, 32, 0.8)', nvsmi_dec))
    series.append(Series('net_trace', 'NET', 'blue', net_traces))
    series.append(Series('tx_bandwidth', 'tx Bandwidth', 'rgba(135,206,250,0.8)', tx_traces, y_field='bandwidth'))
    series.append(Series('rx_bandwidth', 'rx Bandwidth', 'rgba(25,25,112,0.8)', rx_traces, y_field='bandwidth'))
    series.append(Series('gpu_kernel_trace', 'GPU kernel', 'rgba(0,180,0,0.8)', gpu_viz))
    for kw, color, grp in filtered_gpu:
        series.append(Series(kw, '[keyword]' + kw, color, grp.copy()))
    if cfg.cuda_api_tracing:
        series.append(Series('cuda_api_trace', 'CUDA API', 'DarkSlateGray', api_viz))    # :2094-2102
    if emit_json:
        write_report_js(series, logdir + 'report.js')
    return series


# --------------------------------------------------------------------------- a11-a13
def top_k_events(df, topk):
    """sofa_analyze.py:34-44: names ordered by summed duration, descending."""
    agg = df.groupby(['name']).aggregate(np.sum)               # :37 (np.sum per group = pairwise)
    return agg.sort_values(by=['duration'], ascending=False)['duration'].head(topk)


def comm_profile(df_gpu, logdir=None):
    """sofa_common.py:23-177 -> (features dict, payload matrix, count matrix, bw matrix,
    bw-count matrix).  Writes comm.csv when logdir is given (:160-167)."""
    n_gpus = 0
    dev = df_gpu['deviceId'].to_numpy()
    for i in range(len(dev)):                                    # :54-56
        if dev[i] > n_gpus:
            n_gpus = int(dev[i])
    if n_gpus == 0:
        return None
    kinds = (1, 2, 8, 10)
    payload_table = {k: 0 for k in kinds}
    for key, grp in df_gpu.groupby(by='copyKind')['payload']:    # :65-71
        if key in payload_table:
            payload_table[key] = grp.sum()
    bw_table = {k: 0 for k in kinds}
    for key, grp in df_gpu.query('payload > 64000').groupby(by='copyKind')['bandwidth']:   # :76-82
        if key in bw_table:
            bw_table[key] = grp.mean()
    duration_table = {k: 0 for k in kinds}
    for key, grp in df_gpu.groupby(by='copyKind')['duration']:   # :87-93
        if key in duration_table:
            duration_table[key] = grp.sum()
    accum = np.zeros((1 + n_gpus, 1 + n_gpus))
    accum_bw = np.zeros((1 + n_gpus, 1 + n_gpus))
    accum_count = np.zeros((1 + n_gpus, 1 + n_gpus))
    accum_bw_count = np.zeros((1 + n_gpus, 1 + n_gpus))
    ck = df_gpu['copyKind'].to_numpy()
    srcs = df_gpu['pkt_src'].to_numpy()
    dsts = df_gpu['pkt_dst'].to_numpy()
    pls = df_gpu['payload'].to_numpy()
    bws = df_gpu['bandwidth'].to_numpy()
    for i in range(len(df_gpu)):                                 # :108-119 (row order)
        if ck[i] == 0 or ck[i] == 8:
            continue
        s, d = srcs[i], dsts[i]
        accum[s][d] = float(accum[s][d] + pls[i])
        accum_count[s][d] = int(accum_count[s][d] + 1)
        if pls[i] > 64000:
            accum_bw[s][d] = float(accum_bw[s][d] + bws[i])
            accum_bw_count[s][d] = int(accum_bw_count[s][d] + 1)
    if logdir is not None:
        df_gpu.to_csv(logdir + '/' + 'comm.csv',
                      columns=["timestamp", "pkt_src", "pkt_dst", "payload", "bandwidth"])
    feats = {}
    for nm, k in (('h2d', 1), ('d2h', 2), ('d2d', 8), ('p2p', 10)):
        feats[nm + '_payload'] = payload_table[k]
        feats[nm + '_bandwidth'] = bw_table[k]
        feats[nm + '_time'] = duration_table[k]
    return feats, accum, accum_count, accum_bw, accum_bw_count


def gpu_profile(df_gpu, logdir=None):
    """sofa_analyze.py:343-377 -> ordered list of (feature, value)."""
    gpu_time = 0
    groups = df_gpu.groupby("deviceId")["duration"]
    for key, grp in groups:                                      # :347-355
        gpu_time = gpu_time + grp.sum()
    num_gpus = len(groups)
    kernel_time = 0
    for key, grp in df_gpu.groupby("copyKind")["duration"]:      # :357-361
        if key == 0:
            kernel_time = grp.sum()
    nccl_time = 0
    for key, grp in df_gpu.groupby("name")["duration"]:          # :363-368
        if key.find("nccl") != -1:
            nccl_time = nccl_time + grp.sum()
    out = []
    cp = comm_profile(df_gpu, logdir)
    if cp is not None:
        f = cp[0]
        for nm in ('h2d_payload', 'd2h_payload', 'd2d_payload', 'p2p_payload',
                   'h2d_bandwidth', 'd2h_bandwidth', 'd2d_bandwidth', 'p2p_bandwidth',
                   'h2d_time', 'd2h_time', 'd2d_time', 'p2p_time'):
            out.append((nm, f[nm]))
    out += [('gpu_time', gpu_time), ('num_gpus', num_gpus), ('kernel_time', kernel_time),
            ('nccl_time', nccl_time)]
    return out


def cpu_profile(df):
    """sofa_analyze.py:694-710 -> (total_exec_time, top-20 frame by duration)."""
    total = 0
    for key, grp in df.groupby("deviceId")["duration"]:
        total = total + grp.sum()
    d = df[['timestamp', 'duration', 'name']].sort_values(by=['duration'], ascending=False)
    return total, d[:20]


# --------------------------------------------------------------------------- ujson double
def _convertbyte(B):
    """sofa_analyze.py:495-511."""
    B = int(B)
    KB = float(1024)
    MB = float(KB ** 2)
    GB = float(KB ** 3)
    TB = float(KB ** 4)
    if B < KB:
        return '{} Bytes'.format(B)
    elif KB <= B < MB:
        return '{0:.2f} KB'.format(B / KB)
    elif MB <= B < GB:
        return '{0:.2f} MB'.format(B / MB)
    elif GB <= B < TB:
        return '{0:.2f} GB'.format(B / GB)
    return '{0:.2f} TB'.format(B / TB)


def net_profile(df, logdir):
    """sofa_analyze.py:385-493 without the verbose tables: net_time feature and netrank.csv."""
    grouped_df = df.groupby("name")["duration"]
    net_time = 0
    for key, item in grouped_df:
        if key.find("network:tcp:") != -1:
            net_time = net_time + grouped_df.get_group(key).sum()
    packet_sum_matrix = df.groupby(['pkt_src', 'pkt_dst'])["payload"].sum().unstack(level=1, fill_value=0)

    def zero(s):
        if s[0:2] == '00':
            s = s[2]
        elif (s[0] == '0') and (s[1] != '0'):
            s = s[1:3]
        return s

    def ip_text(j):
        j = str(int(j))
        return j[:-9] + '.' + zero(j[-9:-6]) + '.' + zero(j[-6:-3]) + '.' + zero(j[-3:])
    record = []
    for index in packet_sum_matrix.index:
        for column in packet_sum_matrix.columns:
            record.append((ip_text(index), ip_text(column), packet_sum_matrix[column][index]))
    record.sort(key=lambda tup: tup[2], reverse=True)
    final = [[s_, d_, _convertbyte(v), round(v / df['payload'].sum(), 2)] for s_, d_, v in record if v != 0]
    pd.DataFrame(final, columns=['Source', 'Destination', 'Amount', 'Percentage of a Node']).to_csv(
        logdir + 'netrank.csv', mode='w', header=True, index=False)
    return [('net_time', net_time)]


def nvsmi_profile(df_nvsmi, nvsmi_data=False, spotlight_gpu=False, roi_begin=0, roi_end=0):
    """sofa_analyze.py:259-341 -> the six feature rows (verbose tables are print-only)."""
    if spotlight_gpu and roi_end != 0:
        df_nvsmi = df_nvsmi[(df_nvsmi['timestamp'] > roi_begin) & (df_nvsmi['timestamp'] <= roi_end)]
    df_nvsmi.iloc[0].timestamp                                           # :271 raises on an empty frame
    gpu_sm_util = df_nvsmi.groupby(['event'])['duration'].mean()[0]
    gpu_mem_util = df_nvsmi.groupby(['event'])['duration'].mean()[1]
    sm = df_nvsmi['event'] == int(0)
    mem = df_nvsmi['event'] == int(1)
    gpunum = list(set(df_nvsmi['deviceId']))
    gpuid = df_nvsmi['deviceId'] == int(gpunum[-1])                      # the loop variable leaks (:295, :330)
    return [('gpu_sm_util_q2', df_nvsmi[sm & gpuid]['duration'].quantile(0.5)),
            ('gpu_sm_util_q3', df_nvsmi[sm & gpuid]['duration'].quantile(0.75)),
            ('gpu_sm_util', int(gpu_sm_util)),
            ('gpu_mem_util_q2', df_nvsmi[mem & gpuid]['duration'].quantile(0.5)),
            ('gpu_mem_util_q3', df_nvsmi[mem & gpuid]['duration'].quantile(0.75)),
            ('gpu_mem_util', int(gpu_mem_util))]


def ujson_double_str(value, precision=10):
    """pandas' vendored ultrajson Buffer_AppendDoubleUnchecked (what `to_json` prints for a
    float64, double_precision=10) restated for the fixed-notation range; used to pin the
    device JSON float codec.  Returns None outside 1e-15 <= |v| <= 1e16-1 (printf path)."""
    if value != value or value in (math.inf, -math.inf):
        return 'null'
    neg = value < 0
    if neg:
        value = -value
    if value > 1e16 - 1 or (value != 0.0 and value < 1e-15):
        return None
    p10 = float(10 ** precision)
    whole = int(value)
    tmp = (value - whole) * p10
    frac = int(tmp)
    diff = tmp - frac
    if diff > 0.5:
        frac += 1
    elif diff == 0.5 and (frac == 0 or (frac & 1)):
        frac += 1
    if frac >= p10:
        frac = 0
        whole += 1
    fs = ('%0*d' % (precision, frac)).rstrip('0') or '0'
    return ('-' if neg else '') + '%d.%s' % (whole, fs)


# --------------------------------------------------------------------------- a14
def window_ends(t_begin, t_end, sample_time):
    """Window end times of concurrency_breakdown: repeated t = t + sample_time (sofa_analyze.py:89-99)."""
    ends = []
    t = t_begin
    while t < t_end:
        t = t + sample_time
        ends.append(t)
    return ends


def concurrency_breakdown(df_mpstat, df_cpu, df_nvsmi, df_bandwidth, sys_mon_rate=10, is_idle_threshold=10,
                          roi_begin=0, roi_end=0, spotlight_gpu=False):
    """sofa_analyze.py:75-243 -> (feature list, performance table or None).  Windows (t-d, t];
    per window: max usr/sys/idl/iow share of the mpstat rows, SM utilisation max/sum/min of the
    nvsmi rows with event 0, tx/rx bandwidth sums; dominator classes; Pearson table."""
    feats = []
    if len(df_mpstat) == 0:
        return feats, None
    total = {'usr': 0, 'sys': 0, 'gpu': 0, 'iow': 0, 'idl': 0}
    interval_vecs, perf_vecs = [], []
    sample_time = (1 / float(sys_mon_rate))
    t_begin = df_mpstat.iloc[0]['timestamp']
    t_end = df_mpstat.iloc[-1]['timestamp']
    num_gpus = len(list(set(df_nvsmi['deviceId'])))
    for t in window_ends(t_begin, t_end, sample_time):
        if roi_end > 0 and (t < roi_begin or t > roi_end):
            continue
        wb, we = t - sample_time, t
        if len(df_cpu) > 0 and df_cpu.iloc[0].timestamp > we:
            continue
        nv = df_nvsmi[(df_nvsmi['timestamp'] > wb) & (df_nvsmi['timestamp'] <= we) & (df_nvsmi['event'] == int(0))]
        mpi = df_mpstat[(df_mpstat['timestamp'] > wb) & (df_mpstat['timestamp'] <= we)]
        inw = (df_bandwidth['timestamp'] > wb) & (df_bandwidth['timestamp'] <= we)
        tx = df_bandwidth[inw & (df_bandwidth['event'] == float(0))]
        rx = df_bandwidth[inw & (df_bandwidth['event'] == float(1))]
        usr, sysv, idl, iow, irq = [], [], [], [], []
        cpu_max, cpu_min = 0, 100
        for i in range(len(mpi)):
            r = mpi.iloc[i]['name'].split(':')[1].split('|')
            usr.append(int(r[1])); sysv.append(int(r[2])); idl.append(int(r[3])); iow.append(int(r[4])); irq.append(int(r[5]))
            c = int(r[1]) + int(r[2]) + int(r[5])
            cpu_max = max(cpu_max, c)
            cpu_min = min(cpu_min, c)
        if len(mpi) == 0:
            continue
        mp = {k: np.asarray([sample_time * v / 100.0 for v in arr]) for k, arr in
              (('usr', usr), ('sys', sysv), ('idl', idl), ('iow', iow))}
        elapsed = {'usr': mp['usr'].max(), 'sys': mp['sys'].max(),
                   'gpu': nv['duration'].max() * 0.01 * sample_time, 'iow': mp['iow'].max(), 'idl': 0}
        dominator = max(elapsed, key=elapsed.get)
        if elapsed[dominator] > sample_time * int(is_idle_threshold) / 100:
            total[dominator] = total[dominator] + sample_time
        else:
            total['idl'] += sample_time
        time_gpu_avg = nv['duration'].sum() * 0.01 * sample_time / num_gpus if num_gpus > 0 else 0
        interval_vecs.append((mp['usr'].max(), mp['sys'].max(), mp['iow'].max(), mp['idl'].max(), time_gpu_avg,
                              tx['bandwidth'].sum(), rx['bandwidth'].sum()))
        sm_avg = nv['duration'].sum() / int(len(list(set(nv['deviceId'])))) if num_gpus > 0 else 0
        perf_vecs.append((we, nv['duration'].max(), sm_avg, nv['duration'].min(),
                          round((np.asarray(usr).mean() + np.asarray(sysv).mean() + np.asarray(irq).mean()), 0),
                          cpu_max, cpu_min))
    total_all = sum(total.values())
    if total_all > 0:
        ratio = {k: 100 * total[k] / total_all for k in total}
        hotspot = roi_end - roi_begin if spotlight_gpu else 0
        feats += [('elapsed_usr_time_ratio', ratio['usr']), ('elapsed_sys_time_ratio', ratio['sys']),
                  ('elapsed_gpu_time_ratio', ratio['gpu']), ('elapsed_iow_time_ratio', ratio['iow']),
                  ('elapsed_hotspot_time', hotspot)]
    perf = None
    if len(perf_vecs) > 0:
        perf = pd.DataFrame(perf_vecs, columns=['time', 'max_gpu_util', 'avg_gpu_util', 'min_gpu_util', 'cpu_util',
                                                'cpu_max', 'cpu_min'])
        vt = pd.DataFrame(interval_vecs, columns=['usr', 'sys', 'iow', 'idl', 'gpu', 'net_tx', 'net_rx'])
        pearson = vt.corr(method='pearson').round(2)
        feats += [('corr_gpu_usr', pearson['gpu'].usr), ('corr_gpu_sys', pearson['gpu'].sys),
                  ('corr_gpu_iow', pearson['gpu'].iow), ('corr_gpu_ntx', pearson['gpu'].net_tx),
                  ('corr_gpu_nrx', pearson['gpu'].net_rx)]
    return feats, perf
