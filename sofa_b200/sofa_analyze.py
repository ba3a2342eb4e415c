"""Drop-in `sofa_analyze(cfg)` (reference: bin/sofa_analyze.py:793-1055).

Same driver order, same stdout (feature table, verbose tables) and the same files (comm.csv, netrank.csv,
performance.csv) as the reference; pinned byte for byte against the unmodified reference driver on
tests/golden/driver_* (tests/golden/make_golden_driver.py).  Every per-row reduction runs in device kernels
behind libsofa_b200.so: group sums in numpy's pairwise order, radix sorts for top-k / quartiles, hash tables
for the (pkt_src, pkt_dst[, payload]) groups, window reductions, the --spotlight_gpu state machine.  The host
keeps the scalar logic and the printing of the small result tables (a few hundred cells) with pandas, as the
reference does.

Profile functions: cpu_profile, strace_profile (a title), net_profile (+ --verbose matrices), netbandwidth_profile,
blktrace_latency_profile, diskstat_profile, vmstat_profile, mpstat_profile, nvsmi_profile, gpu_profile (+ comm_profile,
get_top_k_events), concurrency_breakdown.  Not built: the PDF charts (no matplotlib), nvlink ring search, POTATO, the AISI
driver hook (sofa_b200.sofa_aisi.iter_profile is the device part of it).

Like the reference, a stand-alone `sofa analyze` (or `sofa report --skip_preprocess`) reads only the CSV files of the
logdir: cputrace.csv / gputrace.csv / nettrace.csv are parsed on the device as pd.read_csv would see them (quoted,
multi-line name fields included), and so are the small side tables (mpstat.csv, nvsmi_trace.csv, netstat.csv,
vmstat.csv).  After `sofa report` the rows sofa_preprocess left resident on the device are analysed directly
(`cfg.keep_engine`): same values, since every reduction reads them through the "%.6f" round trip.
"""
import glob
import os
import subprocess

from . import hostlogic as hl
from .hostlogic import SRC_GPU
from .sofa_print import print_main_progress, print_title, print_warning

CKTABLE = {-1: "NON", 0: "KER", 1: "H2D", 2: "D2H", 8: "D2D", 10: "P2P"}


def comm_profile_report(cfg, gp):
    """Verbose tables of sofa_common.py:62-158 and the 12 features of :169-177."""
    kinds = (1, 2, 8, 10)
    if cfg.verbose:
        print("\n=========== Payload report for data-copy kinds =================")
        for i, k in enumerate(kinds):
            if gp.payload[i] or gp.duration[i]:
                print('%s : %.2lf (MB)' % (CKTABLE[k], gp.payload[i] / 1024.0 ** 2))
        print("\n=========== Bandwidth report for large data-copy (64kb+) =======")
        for i, k in enumerate(kinds):
            if gp.bandwidth[i]:
                print('%s : %.2lf (GB/s)' % (CKTABLE[k], gp.bandwidth[i]))
        print("\n=========== Duration report for large data-copy (64kb+) =======")
        for i, k in enumerate(kinds):
            if gp.payload[i] or gp.duration[i]:
                print('%s : %.2lf (s)' % (CKTABLE[k], gp.duration[i]))
        n = int(gp.n_gpus_max)
        if n <= 15:
            dim = 1 + n
            print("\n=========== Data-copy payload matrix (MB) ======================")
            hdr = "\tHOST\t" + "".join("GPU%d\t" % i for i in range(1, dim))
            print(hdr)
            for i in range(dim):
                row = "HOST\t" if i == 0 else "GPU%d\t" % i
                for j in range(dim):
                    row += "%d" % (gp.accum[i * dim + j] / (1024 * 1024)) + "\t"
                print(row)
            print("\n=========== Data-Copy Bandwidth Matrix (GB/s) =================")
            print(hdr)
            for i in range(dim):
                row = "HOST\t" if i == 0 else "GPU%d\t" % i
                for j in range(dim):
                    if gp.accum_bw_count[i * dim + j] > 0:
                        row += "%.2lf" % (gp.accum_bw[i * dim + j] / gp.accum_bw_count[i * dim + j]) + "\t"
                    else:
                        row += "0\t"
                print(row)
    feats = []
    for j, nm in enumerate(('payload', 'bandwidth', 'time')):
        arr = (gp.payload, gp.bandwidth, gp.duration)[j]
        for i, k in enumerate(('h2d', 'd2h', 'd2d', 'p2p')):
            feats.append(('%s_%s' % (k, nm), arr[i]))
    return feats


def gpu_profile(cfg, eng):
    """sofa_analyze.py:343-377 on the engine's parsed gpu rows -> ordered feature list."""
    gp = eng.gpu_profile(want_matrices=bool(cfg.verbose))
    if cfg.verbose:
        print_title('GPU Profiling')
        print('Per-GPU time (s):')
        for i in range(min(int(gp.num_gpus), 64)):                # :349-354
            print("[%d]: %lf" % (gp.per_gpu_id[i], gp.per_gpu_time[i]))
    feats = []
    if gp.n_gpus_max == 0:
        # sofa_common.py:58-60 calls print_warning with one argument: the reference dies with
        # TypeError here; fail as loudly.
        raise TypeError("comm_profile: no GPU communication traces (deviceId never > 0); the "
                        "reference raises TypeError at sofa_common.py:59")
    feats += comm_profile_report(cfg, gp)
    with open(cfg.logdir + '/' + 'comm.csv', 'wb') as f:          # sofa_common.py:160-167
        f.write(b",timestamp,pkt_src,pkt_dst,payload,bandwidth\n")
        f.write(memoryview(eng.comm_csv()))
    if cfg.verbose:                                   # get_top_k_events (sofa_analyze.py:34-44)
        import pandas as pd
        top = eng.top_k(SRC_GPU, 10)
        print("Top %d Events: " % 10)
        print(pd.DataFrame({'duration': [t for _, t in top]}, index=pd.Index([n for n, _ in top], name='name')))
    feats += [('gpu_time', gp.gpu_time), ('num_gpus', gp.num_gpus), ('kernel_time', gp.kernel_time),
              ('nccl_time', gp.nccl_time)]
    return feats


def cpu_profile(cfg, eng):
    """sofa_analyze.py:694-710 (verbose only): total execution time and the 20 longest samples."""
    total, rows = eng.cpu_profile(cfg.cpu_top_k if hasattr(cfg, 'cpu_top_k') else 20)
    print_title('CPU Profiling:')
    print('elapsed_time (s) = %.6lf' % cfg.elapsed_time)
    print("[%d]: %lf" % (-1, total))
    print("total execution time (s) = %.3lf" % total)
    import numpy as np
    import pandas as pd
    df = pd.DataFrame({'timestamp': [r[1] for r in rows], 'duration': [r[2] for r in rows], 'name': [r[3] for r in rows]},
                      columns=['timestamp', 'duration', 'name'])
    with np.errstate(all='ignore'):
        df['ratio(%)'] = df['duration'] / total * 100
    print(df[['timestamp', 'ratio(%)', 'duration', 'name']].to_string(index=False))
    return total, rows


SRC_MPSTAT_CSV, SRC_NVSMI_CSV, SRC_NETSTAT_CSV = 3, 4, 5


def _csv_body(path):
    with open(path, 'rb') as f:
        data = f.read()
    nl = data.find(b'\n')
    return data[nl + 1:] if nl >= 0 else b''


def _first_last_timestamp(body):
    first = body[:body.find(b',')]
    tail = body.rstrip(b'\n')
    last_line = tail[tail.rfind(b'\n') + 1:]
    return float(first), float(last_line[:last_line.find(b',')])


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


def _ip_text(j):
    """check_str / zero of net_profile (:406-426): the 3-digit groups of the packed address, as the reference prints them."""
    def zero(s):
        if s[0:2] == '00':
            s = s[2]
        elif (s[0] == '0') and (s[1] != '0'):
            s = s[1:3]
        return s
    j = str(int(j))
    return j[:-9] + '.' + zero(j[-9:-6]) + '.' + zero(j[-6:-3]) + '.' + zero(j[-3:])


def _load_trace_csv(eng, logdir, kind, fn):
    """<src>trace.csv -> the engine's parsed rows of source `kind`, as pd.read_csv would give them (the reference's
    sofa_analyze works on these files, :898-977; they are the resume format of `sofa report --skip_preprocess`)."""
    with open(logdir + fn, 'rb') as f:
        data = f.read()
    nl = data.find(b'\n')
    p = hl.base_params(kind, 0.0, True)
    p['api_mode'] = 3
    eng.load(kind, data, p, nl + 1 if nl >= 0 else len(data))
    return True


def load_net_source(cfg, eng):
    """The packets for a stand-alone `sofa analyze`: nettrace.csv (already down-sampled by plot_ratio at preprocess
    time), else the raw net.tmp; returns False when neither is there."""
    logdir = cfg.logdir
    eng._net_rows_are_csv = False
    if os.path.isfile(logdir + 'nettrace.csv'):
        eng._net_rows_are_csv = True
        return _load_trace_csv(eng, logdir, hl.SRC_NET, 'nettrace.csv')
    if not os.path.isfile(logdir + 'net.tmp'):
        return False
    with open(logdir + 'net.tmp', 'rb') as f:
        text = f.read()
    cfg.time_base = hl.read_time_base(logdir, cfg.cpu_time_offset)
    eng.load(hl.SRC_NET, text, hl.base_params(hl.SRC_NET, cfg.time_base, cfg.absolute_timestamp, 0.0))
    return True


def net_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:385-493.  The group-by-name duration sums and the payload sums per (pkt_src, pkt_dst) come
    from the device (rows i % plot_ratio == 0 of the parsed packets = the rows of nettrace.csv); ranking and the
    text of netrank.csv are a few hundred cells on the host."""
    import numpy as np
    if not cfg.cluster_ip:
        print_title("Network Profiling:")
    ratio = 1 if getattr(eng, '_net_rows_are_csv', False) else int(cfg.plot_ratio)     # nettrace.csv rows are the down-sampled ones
    net_time, total, pairs = eng.net_profile(ratio)
    as_float = bool(eng.row_flags(hl.SRC_NET, ratio) & 2)            # a [] packet row makes the payload column float64
    cast = (lambda x: np.float64(x)) if as_float else (lambda x: np.int64(x))
    total = cast(total)
    if cfg.verbose:
        _net_verbose(eng, ratio, pairs, total, as_float)
    record = [(s, d, cast(v)) for (s, d), v in sorted(pairs.items())]          # index-major, ascending, as the unstacked frame
    record.sort(key=lambda tup: tup[2], reverse=True)
    with open(logdir + 'netrank.csv', 'w') as f:
        f.write('Source,Destination,Amount,Percentage of a Node\n')
        for s, d, value in record:
            if value == 0:
                continue
            f.write('%s,%s,%s,%s\n' % (_ip_text(s), _ip_text(d), _convertbyte(value), repr(float(round(value / total, 2)))))
    return features + [('net_time', net_time)]


def _net_verbose(eng, ratio, pairs, total, as_float):
    """The two matrices of sofa_analyze.py:398-461: payload summed per (pkt_src, pkt_dst) and packets per
    (pkt_src, pkt_dst, payload), both from device hash tables, laid out and printed with pandas as the reference does."""
    import numpy as np
    import pandas as pd
    kt = np.float64 if as_float else np.int64

    def frame(d, names, dtype):
        keys = sorted(d)
        idx = pd.MultiIndex.from_arrays([np.asarray([k[i] for k in keys], dtype=kt) for i in range(len(names))], names=names)
        return pd.Series(np.asarray([d[k] for k in keys], dtype=dtype), index=idx).unstack(level=1, fill_value=0)
    psum = frame(pairs, ['pkt_src', 'pkt_dst'], kt)
    pnum = frame(eng.net_triples(ratio), ['pkt_src', 'pkt_dst', 'payload'], np.int64)
    psum = psum.rename(columns={c: _ip_text(c) for c in psum.columns.tolist()})
    pnum = pnum.rename(columns={c: _ip_text(c) for c in pnum.columns.tolist()})
    psum = psum.rename(index={i: _ip_text(i) for i in psum.index.tolist()})
    # (:458 renames level 0 of the packet-count index with set_levels(inplace=True), which current pandas cannot do in
    #  place: the golden stdout of the reference on this stack keeps the numeric level, and so does this)
    print("total amount of network traffic : ", _convertbyte(total), '\n', psum.to_string(), "\n")
    print("total amount of network packets = %d\n" % pnum.sum().sum(), pnum.to_string(), "\n")


def _convertbytes(B):
    """sofa_analyze.py:513-529."""
    B = float(B)
    KB = float(1024)
    MB = float(KB ** 2)
    GB = float(KB ** 3)
    TB = float(KB ** 4)
    if B < KB:
        return '{0:.2f} B/s'.format(B)
    elif KB <= B < MB:
        return '{0:.2f} KB/s'.format(B / KB)
    elif MB <= B < GB:
        return '{0:.2f} MB/s'.format(B / MB)
    elif GB <= B < TB:
        return '{0:.2f} GB/s'.format(B / GB)
    return '{0:.2f} TB/s'.format(B / TB)


def _np_quantile(lo, hi, count, q):
    """numpy's linear quantile between the order statistics at floor / ceil of (count - 1) * q (its _lerp, both branches)."""
    import numpy as np
    vi = np.float64(q) * (count - 1)
    t = vi - np.floor(vi)
    a, b = np.float64(lo), np.float64(hi)
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def _upload_side_csv(eng, logdir, slot, fn, api_mode=0):
    """Body of a 13-column side CSV (mpstat / nvsmi_trace / netstat / vmstat) -> parsed columns of a side-table slot."""
    p = hl.base_params(slot, 0.0, True)
    p['api_mode'] = api_mode
    body = _csv_body(logdir + fn)
    eng.upload(slot, body)
    eng.parse(slot, p)
    return body


def netbandwidth_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:531-594 (the PDF chart needs matplotlib and is not drawn)."""
    if not cfg.cluster_ip and cfg.verbose:
        print_title('Network Bandwidth Profiling:')
    _upload_side_csv(eng, logdir, SRC_NETSTAT_CSV, 'netstat.csv')
    st = {}
    for key, ev in (('tx', 0.0), ('rx', 1.0)):
        cnt, total, qs = eng.side_quantiles(SRC_NETSTAT_CSV, 6, ev)
        if cnt == 0:
            raise ValueError("cannot convert float NaN to integer")          # int(nan mean) in the reference
        st[key] = [_np_quantile(lo, hi, cnt, 0.25 * (i + 1)) for i, (lo, hi) in enumerate(qs)] + [int(total / cnt)]
    with open('%s/netstat.txt' % logdir) as f:
        lines = f.readlines()
    first_line, last_line = lines[0], lines[-1]
    tx_amount = int(last_line.split(',')[1]) - int(first_line.split(',')[1])
    rx_amount = int(last_line.split(',')[2]) - int(first_line.split(',')[2])
    if not cfg.cluster_ip and cfg.verbose:
        print('Amount of Network Traffic : %s' % (_convertbyte(tx_amount + rx_amount)))
        print('Amount of tx : %s' % _convertbyte(tx_amount))
        print('Amount of rx : %s' % _convertbyte(rx_amount))
        print('Bandwidth Quartile :')
        for i, nm in enumerate(('Q1 ', 'Q2 ', 'Q3 ')):
            print('%s tx : %s, rx : %s' % (nm, _convertbytes(st['tx'][i]), _convertbytes(st['rx'][i])))
        print('Avg tx : %s, rx : %s' % (_convertbytes(st['tx'][3]), _convertbytes(st['rx'][3])))
        print('Network Bandwidth Chart is saved at %s/network_report.pdf' % logdir)
        features = features + [('bw_tx_q2', st['tx'][1]), ('bw_tx_q3', st['tx'][2]), ('bw_rx_q2', st['rx'][1]), ('bw_rx_q3', st['rx'][2])]
    return features


def blktrace_latency_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:596-638.  btt.txt / offset_all.txt are written by the recorder's `btt` run (sofa_preprocess.py:684-686);
    when one is missing open() raises inside the driver's try block and the reference reports "blktrace.csv is not found".
    The quartiles of the 'C' rows' duration come from the device; the block-offset chart needs matplotlib and is not drawn."""
    with open('%s/btt.txt' % logdir) as f:
        lines = f.readlines()
    start = end = None
    for i, line in enumerate(lines):
        if '==================== All Devices ====================' in line:
            start = i
        if '==================== Device Merge Information ====================' in line:
            end = i
            break
    if start is None or end is None:
        raise NameError("btt.txt: the 'All Devices' / 'Device Merge Information' sections are missing (UnboundLocalError in the reference)")
    bttoutput_result = lines[start:end]
    open('%s/offset_all.txt' % logdir).close()                       # read for the chart only; must exist
    print('Offset of Device Report is saved at %s/offset_of_device_report.pdf' % logdir)
    if cfg.verbose:
        print_title('Storage Profiling:')
        print('Blktracae Latency (s):')
        for btt in bttoutput_result:
            print(btt[:-1])
    with open(logdir + 'blktrace.csv', 'rb') as f:
        n, lo, hi = eng.blktrace_latency(f.read())
    qs = [(_np_quantile(lo[i], hi[i], n, 0.25 * (i + 1)) if n else float('nan')) for i in range(3)]
    return features + [('blktrace_latency_q1', qs[0]), ('blktrace_latency_q2', qs[1]), ('blktrace_latency_q3', qs[2])]


def diskstat_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:640-700.  Per device (groupby sorts the names): quartiles and means of diskstat_vector.csv's
    columns from the device (order statistics + Kahan sums, sofa_b200_diskstat_profile); the string table is pandas'."""
    import numpy as np
    import pandas as pd
    with open(logdir + 'diskstat_vector.csv', 'rb') as f:
        groups = sorted(eng.diskstat_profile(f.read()), key=lambda g: g[0])

    def gq(g, col, k):                                               # group_quantile, linear (pandas/_libs/groupby.pyx)
        lo, hi, n = np.float64(g[2][col * 3 + k]), np.float64(g[3][col * 3 + k]), g[1]
        q_idx = np.float64(0.25 * (k + 1)) * (n - 1)
        frac = q_idx % 1
        return lo if frac == 0.0 else lo + (hi - lo) * frac
    columns = ['Q1 throughput(Read)', 'Q2 throughput(Read)', 'Q3 throughput(Read)', 'Avg throughput(Read)',
               'Q1 throughput(Write)', 'Q2 throughput(Write)', 'Q3 throughput(Write)', 'Avg throughput(Write)',
               'Q1 throughput(R+W)', 'Q2 throughput(R+W)', 'Q3 throughput(R+W)', 'Avg throughput(R+W)',
               'Avg IOPS(Read)', 'Avg IOPS(Write)', 'Avg IOPS(R+W)', 'Avg Await time(ms)']
    final_table = pd.DataFrame(columns=columns)
    for g in groups:
        mean = [np.float64(x) for x in g[4]]
        cells = []
        for col in range(3):                                         # read, write, read + write
            cells += [_convertbytes(gq(g, col, 0)), _convertbytes(gq(g, col, 1)), _convertbytes(gq(g, col, 2)), _convertbytes(mean[col])]
        cells += ['%d' % int(mean[3]), '%d' % int(mean[4]), '%d' % int(mean[5]), '%.3lf ms' % mean[6]]
        final_table = pd.concat([final_table, pd.DataFrame([tuple(cells)], columns=columns, index=[g[0]])])
    if cfg.verbose:
        print_title('DISKSTAT Profiling:')
        print('Disk Throughput Quartile :')
        print(final_table.T)
    feats = []
    for k, nm in enumerate(('diskstat_q1', 'diskstat_q2', 'diskstat_q3')):
        per_dev = np.array([gq(g, 2, k) for g in groups], dtype=np.float64)
        feats.append((nm, per_dev.sum() / len(per_dev) if len(per_dev) else float('nan')))     # Series.mean()
    return features + feats


def vmstat_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:712-733: means of bi / bo / in / cs parsed out of the name column of vmstat.csv."""
    _upload_side_csv(eng, logdir, SRC_MPSTAT_CSV, 'vmstat.csv', api_mode=2)
    (bi, bo, in_, cs, _), n = eng.side_int_sums(SRC_MPSTAT_CSV)
    import numpy as np
    vm_bi, vm_bo, vm_in, vm_cs = (np.float64(v) / n for v in (bi, bo, in_, cs))
    if cfg.verbose:
        print_title('VMSTAT Profiling:')
        print('average bi/s: %d' % int(vm_cs))            # (labels as swapped as in the reference)
        print('average bo/s: %d' % int(vm_in))
        print('average cs/s: %d' % int(vm_bi))
        print('average in/s: %d' % int(vm_bo))
    return features + [('vm_bi', vm_bi), ('vm_bo', vm_bo), ('vm_cs', vm_cs), ('vm_in', vm_in)]


def mpstat_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:735-790.  Per-core sums of t_USR .. t_IRQ on the device (stable sort by core + numpy-pairwise
    segment sums); the per-core table is a few dozen cells and is reduced and printed with pandas as in the reference
    (with :752-756's chained assignments taking effect, as on the pandas versions the reference was written for)."""
    import pandas as pd
    if not cfg.cluster_ip and cfg.verbose:
        print_title('MPSTAT Profiling:')
    _upload_side_csv(eng, logdir, SRC_MPSTAT_CSV, 'mpstat.csv')
    sums = eng.mpstat_profile()
    num_cores = len(sums)
    df_summary = pd.DataFrame(sums, columns=['USR', 'SYS', 'IDL', 'IOW', 'IRQ'], dtype='float64')
    verbose = not cfg.cluster_ip and cfg.verbose
    if verbose:
        print('CPU Utilization (%):')
        print('core\tUSR\tSYS\tIDL\tIOW\tIRQ')
    for i in range(len(df_summary)):
        t_sum = df_summary.iloc[i].sum()
        if verbose:
            print('%3d\t%3d\t%3d\t%3d\t%3d\t%3d' % (i, int(100.0 * df_summary.iloc[i]['USR'] / t_sum),
                                                      int(100.0 * df_summary.iloc[i]['SYS'] / t_sum),
                                                      int(100.0 * df_summary.iloc[i]['IDL'] / t_sum),
                                                      int(100.0 * df_summary.iloc[i]['IOW'] / t_sum),
                                                      int(100.0 * df_summary.iloc[i]['IRQ'] / t_sum)))
    if verbose:
        print('CPU Time (s):')
        print('core\tUSR\tSYS\tIDL\tIOW\tIRQ')
        for i in range(len(df_summary)):
            print('%3d\t%.2lf\t%.2lf\t%.2lf\t%.2lf\t%.2lf' % (i, df_summary.iloc[i]['USR'], df_summary.iloc[i]['SYS'],
                                                                df_summary.iloc[i]['IDL'], df_summary.iloc[i]['IOW'],
                                                                df_summary.iloc[i]['IRQ']))
    total_cpu_time = df_summary[['USR', 'SYS', 'IRQ']].sum().sum()
    cpu_util = int(100 * total_cpu_time / (num_cores * cfg.elapsed_time))
    if verbose:
        print('Active CPU Time (s): %.3lf' % total_cpu_time)
        print('Active CPU ratio (%%): %3d' % cpu_util)
    return features + [('num_cores', num_cores), ('cpu_util', cpu_util)]


def spotlight_roi(logdir, cfg, eng):
    """sofa_analyze.py:872-895: cfg.roi_begin / cfg.roi_end from the SM utilisation of GPU 0 (device state machine)."""
    _upload_side_csv(eng, logdir, SRC_NVSMI_CSV, 'nvsmi_trace.csv')
    cfg.roi_begin, cfg.roi_end = eng.nvsmi_roi(cfg.roi_begin, cfg.roi_end)


def _lerp(lo, hi, count, q):
    """numpy's linear quantile between the two order statistics around (count-1)*q."""
    vi = (count - 1) * q
    gamma = vi - int(vi)
    return lo + (hi - lo) * gamma


def nvsmi_profile(logdir, cfg, eng, features):
    """sofa_analyze.py:259-341.  Group counts, duration sums and the order statistics behind the
    quartiles come from one device pass over nvsmi_trace.csv (key build + radix sort); the handful of
    divisions and the verbose tables stay on the host."""
    from . import hostlogic as hl2
    body = _csv_body(logdir + 'nvsmi_trace.csv')
    eng.upload(SRC_NVSMI_CSV, body)
    eng.parse(SRC_NVSMI_CSV, hl2.base_params(SRC_NVSMI_CSV, 0.0, True))
    if not cfg.cluster_ip and cfg.verbose:
        print_title('SM & MEM & ENCODE/DECODE Profiling')
    roi = None
    if cfg.spotlight_gpu:
        if cfg.roi_end == 0:
            print_warning(cfg, 'spotlight_gpu has no effects.')
        else:
            roi = (cfg.roi_begin, cfg.roi_end)
    groups, dev_order, ts_first, ts_last, used = eng.nvsmi_profile(roi)
    if used == 0:
        raise IndexError("single positional indexer is out-of-bounds")          # df_nvsmi.iloc[0] on an empty frame

    def ev_mean(ev):                                             # groupby(['event'])['duration'].mean()[ev]
        cnt = sum(g[0] for (d, e), g in groups.items() if e == ev)
        if cnt == 0:
            raise KeyError(ev)
        return sum(g[1] for (d, e), g in groups.items() if e == ev) / cnt
    gpu_sm_util, gpu_mem_util = ev_mean(0), ev_mean(1)
    gpu_enc_util, gpu_dec_util = (ev_mean(2), ev_mean(3)) if cfg.nvsmi_data else (0, 0)
    nan = float('nan')

    import numpy as np

    def mean(dev, ev):
        g = groups.get((dev, ev))
        return np.float64(g[1] / g[0] if g else nan)

    def quant(dev, ev, qi):
        g = groups.get((dev, ev))
        return np.float64(_lerp(g[2][qi][0], g[2][qi][1], g[0], 0.25 * (qi + 1)) if g else nan)
    gpunum = list(set(dev_order))                                # iteration order of the reference's set
    if not cfg.cluster_ip and cfg.verbose:
        import pandas as pd
        res = pd.DataFrame([], columns=['sm', 'mem', 'enc', 'dec'])
        sm_q = pd.DataFrame([], columns=['Q1', 'Q2', 'Q3', 'Avg'])
        mem_q = pd.DataFrame([], columns=['Q1', 'Q2', 'Q3', 'Avg'])
        for i in gpunum:                                         # frames grown row by row as in :293-318 (their dtypes shape the print)
            res = pd.concat([res, pd.DataFrame([[round(mean(i, e), 2) for e in range(4)]], columns=['sm', 'mem', 'enc', 'dec'], index=[i])])
            sm_q = pd.concat([sm_q, pd.DataFrame([[round(quant(i, 0, 0), 2), round(quant(i, 0, 1), 2), round(quant(i, 0, 2), 2),
                                                   round(mean(i, 0), 2)]], columns=['Q1', 'Q2', 'Q3', 'Avg'], index=[i])])
            mem_q = pd.concat([mem_q, pd.DataFrame([[round(quant(i, 1, 0), 2), round(quant(i, 1, 1), 2), round(quant(i, 1, 2), 2),
                                                     round(mean(i, 1), 2)]], columns=['Q1', 'Q2', 'Q3', 'Avg'], index=[i])])
        for d in (res, sm_q, mem_q):
            d.index.name = 'gpu_id'
        print('GPU Utilization (%):')
        print(res)
        print('\nGPU SM Quartile (%):')
        print(sm_q)
        print('\nGPU MEM Quartile (%):')
        print(mem_q)
        print('Overall Average SM Utilization (%): ', int(gpu_sm_util))
        print('Overall Average MEM Utilization (%): ', int(gpu_mem_util))
        print('Overall Average ENC Utilization (%): ', int(gpu_enc_util))
        print('Overall Average DEC Utilization (%): ', int(gpu_dec_util))
        print('Overall Active GPU Time (s): %.3lf' % ((ts_last - ts_first) * gpu_sm_util / 100.0))
    last = gpunum[-1]                                            # `gpuid` leaks out of the reference's loop (:330-336)
    return features + [('gpu_sm_util_q2', quant(last, 0, 1)), ('gpu_sm_util_q3', quant(last, 0, 2)),
                       ('gpu_sm_util', int(gpu_sm_util)), ('gpu_mem_util_q2', quant(last, 1, 1)),
                       ('gpu_mem_util_q3', quant(last, 1, 2)), ('gpu_mem_util', int(gpu_mem_util))]


def concurrency_breakdown(logdir, cfg, eng, features):
    """sofa_analyze.py:75-243.  The per-window masks over mpstat.csv / nvsmi_trace.csv / netstat.csv
    (O(windows x rows) in the reference) are one device pass (`window_reduce`); the scalar logic per
    window and the small tables (performance.csv, Pearson) stay on the host as in the reference."""
    import numpy as np
    import pandas as pd
    from . import hostlogic as hl2
    if cfg.verbose:
        print_title('Concurrency Breakdown Analysis')
    bodies = {}
    for k, fn in ((SRC_MPSTAT_CSV, 'mpstat.csv'), (SRC_NVSMI_CSV, 'nvsmi_trace.csv'), (SRC_NETSTAT_CSV, 'netstat.csv')):
        bodies[k] = _csv_body(logdir + fn) if os.path.isfile(logdir + fn) else b''
    if len(bodies[SRC_MPSTAT_CSV]) == 0:
        print_warning(cfg, 'no mpstat and perf traces!')
        return features
    for k in (SRC_MPSTAT_CSV, SRC_NVSMI_CSV, SRC_NETSTAT_CSV):
        eng.upload(k, bodies[k])
        eng.parse(k, hl2.base_params(k, 0.0, True))
    t_begin, t_end = _first_last_timestamp(bodies[SRC_MPSTAT_CSV])
    sample_time = (1 / float(cfg.sys_mon_rate))
    first_cpu = None
    if os.path.isfile(logdir + 'cputrace.csv'):
        with open(logdir + 'cputrace.csv', 'rb') as f:           # only the first row's timestamp is needed (the file can be GBs)
            head = f.read(1 << 12)
        cb = head[head.find(b'\n') + 1:] if b'\n' in head else b''
        if cb:
            first_cpu = float(cb[:cb.find(b',')])
    ends = []
    t = t_begin
    while t < t_end:                                             # :89-99, repeated addition
        t = t + sample_time
        ends.append(t)
    if not ends:
        return features
    wins, num_gpus = eng.window_reduce(ends, sample_time)
    total = {'usr': 0, 'sys': 0, 'gpu': 0, 'iow': 0, 'idl': 0}
    interval_vecs, perf_vecs = [], []
    nan = float('nan')
    for t, w in zip(ends, wins):
        if cfg.roi_end > 0 and (t < cfg.roi_begin or t > cfg.roi_end):
            continue
        if first_cpu is not None and first_cpu > t:
            continue
        if w.mp_count == 0:
            continue
        usr_max = np.float64(sample_time * w.usr_max / 100.0)
        sys_max = np.float64(sample_time * w.sys_max / 100.0)
        idl_max = np.float64(sample_time * w.idl_max / 100.0)
        iow_max = np.float64(sample_time * w.iow_max / 100.0)
        nv_max = np.int64(int(w.nv_max)) if w.nv_count else nan
        nv_min = np.int64(int(w.nv_min)) if w.nv_count else nan
        nv_sum = np.int64(w.nv_sum)
        elapsed = {'usr': usr_max, 'sys': sys_max, 'gpu': nv_max * 0.01 * sample_time, 'iow': iow_max, 'idl': 0}
        dominator = max(elapsed, key=elapsed.get)
        if elapsed[dominator] > sample_time * int(cfg.is_idle_threshold) / 100:
            total[dominator] = total[dominator] + sample_time
        else:
            total['idl'] += sample_time
        time_gpu_avg = nv_sum * 0.01 * sample_time / num_gpus if num_gpus > 0 else 0
        interval_vecs.append((usr_max, sys_max, iow_max, idl_max, time_gpu_avg,
                              w.tx_sum if w.tx_count else 0.0, w.rx_sum if w.rx_count else 0.0))
        with np.errstate(all='ignore'):
            sm_avg = nv_sum / int(w.nv_ndev) if num_gpus > 0 else 0
        cpu_util = round((np.float64(w.usr_sum) / w.mp_count + np.float64(w.sys_sum) / w.mp_count +
                          np.float64(w.irq_sum) / w.mp_count), 0)
        perf_vecs.append((t, nv_max, sm_avg, nv_min, cpu_util, w.cpu_max, w.cpu_min))
    total_all = sum(total.values())
    if total_all > 0:
        ratio = {k: 100 * total[k] / total_all for k in total}
        if cfg.verbose:
            print('Elapsed Time = %.1lf ' % total_all)
            print('USR = %.1lf %%' % ratio['usr'])
            print('SYS = %.1lf %%' % ratio['sys'])
            if num_gpus > 0:
                print('GPU = %.1lf %%' % ratio['gpu'])
            print('IDL = %.1lf %%' % ratio['idl'])
            print('IOW = %.1lf %%' % ratio['iow'])
        hotspot = cfg.roi_end - cfg.roi_begin if cfg.spotlight_gpu else 0
        features += [('elapsed_usr_time_ratio', ratio['usr']), ('elapsed_sys_time_ratio', ratio['sys']),
                     ('elapsed_gpu_time_ratio', ratio['gpu']), ('elapsed_iow_time_ratio', ratio['iow']),
                     ('elapsed_hotspot_time', hotspot)]
    if len(perf_vecs) > 0:
        perf = pd.DataFrame(perf_vecs, columns=['time', 'max_gpu_util', 'avg_gpu_util', 'min_gpu_util', 'cpu_util',
                                                'cpu_max', 'cpu_min'])
        perf.to_csv('%s/performance.csv' % logdir)
        vt = pd.DataFrame(interval_vecs, columns=['usr', 'sys', 'iow', 'idl', 'gpu', 'net_tx', 'net_rx'])
        pearson = vt.corr(method='pearson').round(2)
        if cfg.verbose:
            print('Correlation Table :')
            print(pearson)
        features += [('corr_gpu_usr', pearson['gpu'].usr), ('corr_gpu_sys', pearson['gpu'].sys),
                     ('corr_gpu_iow', pearson['gpu'].iow), ('corr_gpu_ntx', pearson['gpu'].net_tx),
                     ('corr_gpu_nrx', pearson['gpu'].net_rx)]
    return features


def load_perf_source(cfg, eng):
    """(Re-)ingest perf.script exactly as sofa_preprocess does (sofa_preprocess.py:1786-1807)."""
    logdir = cfg.logdir
    if os.path.isfile(logdir + 'cputrace.csv'):
        return _load_trace_csv(eng, logdir, hl.SRC_PERF, 'cputrace.csv')
    if not os.path.isfile(logdir + 'perf.script'):
        return False
    with open(logdir + 'perf.script', 'rb') as f:
        text = f.read()
    if not text:
        return False
    from .hostlogic import SRC_PERF
    time_base = hl.read_time_base(logdir, cfg.cpu_time_offset)
    p = hl.base_params(SRC_PERF, time_base, cfg.absolute_timestamp, hl.read_perf_timebase(logdir))
    p['perf_has_cycles'] = 1 if hl.read_perf_events(logdir).find('cycles') != -1 else 0
    p['mhz_xp'], p['mhz_fp'] = hl.read_cpuinfo(logdir)
    eng.load(SRC_PERF, text, p)
    return True


def load_gpu_source(cfg, eng):
    """(Re-)ingest gputrace.tmp exactly as sofa_preprocess does (sofa_preprocess.py:1343-1444)."""
    logdir = cfg.logdir
    if os.path.isfile(logdir + 'gputrace.csv'):
        return _load_trace_csv(eng, logdir, SRC_GPU, 'gputrace.csv')
    if not os.path.isfile(logdir + 'gputrace.tmp'):
        return False
    with open(logdir + 'gputrace.tmp', 'rb') as f:
        text = f.read()
    hdr = hl.gpu_header(text[:1 << 16]) if text else None
    if hdr is None:
        return False
    from .sofa_preprocess import gpu_first_timestamp
    nvvps = glob.glob(logdir + "gputrace*[0-9].nvvp")
    base = gpu_first_timestamp(nvvps[0]) if nvvps else 0
    upd, data_off = hdr
    time_base = hl.read_time_base(logdir, cfg.cpu_time_offset)
    p = hl.base_params(SRC_GPU, time_base, cfg.absolute_timestamp, base - upd.pop('t_base'))
    p.update(upd)
    eng.load(SRC_GPU, text, p, data_off)
    return True


def _has_rows(path):
    """the file exists and pd.read_csv would give a non-empty frame (more than the header line)"""
    if not os.path.isfile(path):
        return False
    with open(path, 'rb') as f:
        head = f.read(1 << 16)
    nl = head.find(b'\n')
    return nl >= 0 and (len(head) > nl + 1 or os.path.getsize(path) > nl + 1)


class _Phase(object):
    """Wall time of one analyze section, added to cfg._timing (read by bench.py's e2e_entrypoints)."""
    def __init__(self, cfg, name):
        self.cfg, self.name = cfg, 'analyze: ' + name

    def __enter__(self):
        import time
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        import time
        tm = getattr(self.cfg, '_timing', None)
        if tm is None:
            tm = self.cfg._timing = {}
        tm[self.name] = tm.get(self.name, 0.0) + time.perf_counter() - self.t0
        return False


def sofa_analyze(cfg, engine=None):
    print_main_progress('SOFA analyzing...')
    logdir = cfg.logdir
    with open(logdir + '/misc.txt') as f:
        lines = f.readlines()
        cfg.elapsed_time = float(lines[0].split()[1])
    features = [('elapsed_time', cfg.elapsed_time)]
    resident = engine is None and getattr(cfg, '_sofa_b200_engine', None) is not None      # `sofa report`: parsed trace rows still on the device
    eng = engine or getattr(cfg, '_sofa_b200_engine', None)
    own = False
    if eng is None:
        from .engine import CudaEngine
        eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
        own = True
    try:
        # ---- --spotlight_gpu region of interest (:872-897)
        if os.path.isfile(logdir + "nvsmi_trace.csv"):
            if _has_rows(logdir + "nvsmi_trace.csv") and cfg.spotlight_gpu:
                spotlight_roi(logdir, cfg, eng)
        else:
            print_warning(cfg, "nvsmi_trace.csv is not found")
        # ---- cpu (:899-907)
        if os.path.isfile(logdir + "cputrace.csv"):
            if cfg.verbose and _has_rows(logdir + "cputrace.csv"):
                if not resident:
                    load_perf_source(cfg, eng)
                cpu_profile(cfg, eng)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "cputrace.csv"))
        # ---- strace (:909-915)
        if os.path.isfile(logdir + "strace.csv"):
            if _has_rows(logdir + "strace.csv"):
                print_title('STRACE Profiling:')                         # strace_profile (:379-382)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "strace.csv"))
        # ---- net (:917-923)
        if os.path.isfile(logdir + "nettrace.csv"):
            if _has_rows(logdir + "nettrace.csv"):
                if not resident and not load_net_source(cfg, eng):
                    raise IOError("nettrace.csv could not be loaded")
                with _Phase(cfg, 'net_profile'):
                    features = net_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "nettrace.csv"))
        # ---- netstat.csv (:925-931)
        if os.path.isfile(logdir + "netstat.csv"):
            if _has_rows(logdir + "netstat.csv"):
                with _Phase(cfg, 'netbandwidth_profile'):
                    features = netbandwidth_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "netstat.csv"))
        # ---- blktrace (:933-939), diskstat (:941-947)
        if os.path.isfile(logdir + "blktrace.csv"):
            if _has_rows(logdir + "blktrace.csv"):
                try:
                    with _Phase(cfg, 'blktrace_latency_profile'):
                        features = blktrace_latency_profile(logdir, cfg, eng, features)
                except IOError:                                  # btt.txt / offset_all.txt missing: the reference's message
                    print_warning(cfg, "%s is not found" % (logdir + "blktrace.csv"))
        else:
            print_warning(cfg, "%s is not found" % (logdir + "blktrace.csv"))
        if os.path.isfile(logdir + "diskstat_vector.csv"):
            if _has_rows(logdir + "diskstat_vector.csv"):
                with _Phase(cfg, 'diskstat_profile'):
                    features = diskstat_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "diskstat_vector.csv"))
        # ---- vmstat (:949-955), mpstat (:957-963)
        if os.path.isfile(logdir + "vmstat.csv"):
            if _has_rows(logdir + "vmstat.csv"):
                with _Phase(cfg, 'vmstat_profile'):
                    features = vmstat_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "vmstat.csv"))
        have_mpstat = False
        if os.path.isfile(logdir + "mpstat.csv"):
            if _has_rows(logdir + "mpstat.csv"):
                have_mpstat = True
                with _Phase(cfg, 'mpstat_profile'):
                    features = mpstat_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "%s is not found" % (logdir + "mpstat.csv"))
        # ---- nvsmi (:965-969), gpu (:971-977)
        if os.path.isfile(logdir + "nvsmi_trace.csv"):
            with _Phase(cfg, 'nvsmi_profile'):
                features = nvsmi_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "nvsmi_trace.csv is not found")
        if os.path.isfile(logdir + "gputrace.csv"):
            if _has_rows(logdir + "gputrace.csv"):
                if (not resident or getattr(cfg, '_sofa_b200_gpu_slot_is_api', False)) and not load_gpu_source(cfg, eng):
                    raise IOError("gputrace.csv could not be loaded")
                with _Phase(cfg, 'gpu_profile'):
                    features += gpu_profile(cfg, eng)
        else:
            print_warning(cfg, "%sgputrace.csv is not found. If there is no need to profile GPU, just ignore it." % logdir)
        # ---- concurrency_breakdown (:979-984)
        if have_mpstat:
            with _Phase(cfg, 'concurrency_breakdown'):
                features = concurrency_breakdown(logdir, cfg, eng, features)
    finally:
        if own:
            eng.close()
        if getattr(cfg, '_sofa_b200_engine', None) is not None and engine is None:
            cfg._sofa_b200_engine.close()
            cfg._sofa_b200_engine = None
    if 'IS_SOFA_ON_HAIHUB' not in os.environ or os.environ['IS_SOFA_ON_HAIHUB'] == 'no':
        print_title('Final Performance Features')
        print('%s%s%s%s' % ('ID'.ljust(10), 'Feature'.ljust(30), 'Value'.ljust(20), 'Unit'.ljust(20)))
        for i, (name, value) in enumerate(features):
            print('%s%s%s' % (str(i).ljust(10), name.ljust(30), ('%.3lf' % value).ljust(20)))
    if cfg.spotlight_gpu:
        # :1001-1005 formats a one-row Series with '%.3lf': a TypeError on current pandas, caught by the reference's
        # bare except - the pinned output (tests/golden/driver_spotlight) is this warning
        print_warning(cfg, 'elpased_hostspot_time is not defined.')
    cfg.features = features
    sofa_home = os.path.dirname(os.path.realpath(__file__))
    board = os.path.join(sofa_home, '..', 'sofaboard')
    if os.path.isdir(board):
        subprocess.Popen(['bash', '-c', 'cp %s/* %s;' % (board, cfg.logdir)])
    print('\n\n')
    print('Complete!!')
    return features
