"""Drop-in `sofa_analyze(cfg)` (reference: bin/sofa_analyze.py:793-1055).

Round-1 scope: gpu_profile + comm_profile features and the
"Final Performance Features" table, computed by device kernels on the parsed columns
(values read through the "%.6f" round trip the reference's pd.read_csv sees; float sums
follow numpy's pairwise tree exactly).  Profiles of sources that are not on the hot path
yet (mpstat / nvsmi / netstat / vmstat / diskstat / strace, concurrency_breakdown, aisi,
potato) are absent; their input files make sofa_preprocess raise, so no silent gaps.

The reference re-reads <src>trace.csv; here the rows are re-ingested from the raw logs in
`cfg.logdir` (same bytes -> same columns), or taken from the engine sofa_preprocess left
resident when `cfg.keep_engine` is set (`sofa report`).
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
        print("Top %d Events: " % 10)
        for name, total in eng.top_k(SRC_GPU, 10):
            print("%-60s %.6f" % (name.rstrip('\n')[:60], total))
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
    print("%12s %10s %10s  %s" % ('timestamp', 'ratio(%)', 'duration', 'name'))
    for _, ts, dur, name in rows:
        print("%12.6f %10.6f %10.6f  %s" % (ts, dur / total * 100 if total else 0.0, dur, name))
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


def load_net_source(cfg, eng):
    """Re-ingest net.tmp (for a stand-alone `sofa analyze`); returns False when it is absent."""
    logdir = cfg.logdir
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
    ratio = int(cfg.plot_ratio)
    net_time, total, pairs = eng.net_profile(ratio)
    as_float = bool(eng.row_flags(hl.SRC_NET, ratio) & 2)            # a [] packet row makes the payload column float64
    cast = (lambda x: np.float64(x)) if as_float else (lambda x: np.int64(x))
    total = cast(total)
    if cfg.verbose:
        raise NotImplementedError("net_profile --verbose tables are not implemented")
    record = [(s, d, cast(v)) for (s, d), v in sorted(pairs.items())]          # index-major, ascending, as the unstacked frame
    record.sort(key=lambda tup: tup[2], reverse=True)
    with open(logdir + 'netrank.csv', 'w') as f:
        f.write('Source,Destination,Amount,Percentage of a Node\n')
        for s, d, value in record:
            if value == 0:
                continue
            f.write('%s,%s,%s,%s\n' % (_ip_text(s), _ip_text(d), _convertbyte(value), repr(float(round(value / total, 2)))))
    return features + [('net_time', net_time)]


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
        cb = _csv_body(logdir + 'cputrace.csv')
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


def sofa_analyze(cfg, engine=None):
    print_main_progress('SOFA analyzing...')
    logdir = cfg.logdir
    with open(logdir + '/misc.txt') as f:
        lines = f.readlines()
        cfg.elapsed_time = float(lines[0].split()[1])
    features = [('elapsed_time', cfg.elapsed_time)]
    own = False
    eng = engine or getattr(cfg, '_sofa_b200_engine', None)
    try:
        if cfg.verbose and os.path.isfile(logdir + "cputrace.csv"):
            if eng is None:
                from .engine import CudaEngine
                eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
                own = True
            if own or engine is not None:
                load_perf_source(cfg, eng)
            cpu_profile(cfg, eng)
        if os.path.isfile(logdir + "nettrace.csv") and os.path.getsize(logdir + "nettrace.csv") > len(",".join(['x'] * 13)):   # :917-923
            if eng is None:
                from .engine import CudaEngine
                eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
                own = True
            if own or engine is not None:
                if not load_net_source(cfg, eng):
                    raise IOError("nettrace.csv exists but net.tmp (raw packets) is missing: CSV re-ingest is not implemented")
            features = net_profile(logdir, cfg, eng, features)
        if os.path.isfile(logdir + "nvsmi_trace.csv"):                     # :957-961
            if eng is None:
                from .engine import CudaEngine
                eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
                own = True
            features = nvsmi_profile(logdir, cfg, eng, features)
        else:
            print_warning(cfg, "nvsmi_trace.csv is not found")
        if os.path.isfile(logdir + "gputrace.csv"):
            if eng is None:
                from .engine import CudaEngine
                eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
                own = True
            if own or engine is not None or getattr(cfg, '_sofa_b200_gpu_slot_is_api', False):
                if not load_gpu_source(cfg, eng):
                    raise IOError("gputrace.csv exists but gputrace.tmp (raw rows) is missing: "
                                  "CSV re-ingest is not implemented")
            features += gpu_profile(cfg, eng)
        else:
            print_warning(cfg, "%sgputrace.csv is not found. If there is no need to profile GPU, just ignore it." % logdir)
        if os.path.isfile(logdir + "mpstat.csv") and os.path.getsize(logdir + "mpstat.csv") > 0:     # :979-984
            if eng is None:
                from .engine import CudaEngine
                eng = CudaEngine(int(os.environ.get('SOFA_B200_DEVICE', os.environ.get('LOCAL_RANK', 0))))
                own = True
            features = concurrency_breakdown(logdir, cfg, eng, features)
    finally:
        if own and eng is not None:
            eng.close()
        if getattr(cfg, '_sofa_b200_engine', None) is not None and engine is None:
            cfg._sofa_b200_engine.close()
            cfg._sofa_b200_engine = None
    if 'IS_SOFA_ON_HAIHUB' not in os.environ or os.environ['IS_SOFA_ON_HAIHUB'] == 'no':
        print_title('Final Performance Features')
        print('%s%s%s%s' % ('ID'.ljust(10), 'Feature'.ljust(30), 'Value'.ljust(20), 'Unit'.ljust(20)))
        for i, (name, value) in enumerate(features):
            print('%s%s%s' % (str(i).ljust(10), name.ljust(30), ('%.3lf' % value).ljust(20)))
    cfg.features = features
    sofa_home = os.path.dirname(os.path.realpath(__file__))
    board = os.path.join(sofa_home, '..', 'sofaboard')
    if os.path.isdir(board):
        subprocess.Popen(['bash', '-c', 'cp %s/* %s;' % (board, cfg.logdir)])
    print('\n\n')
    print('Complete!!')
    return features
