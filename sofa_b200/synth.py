"""Deterministic synthetic SOFA logdirs (test / bench infrastructure, not product).

The reference ships no sample logdir, so every parity input is
generated.  The generator is a pure function of (seed, stream, row, field)
through a splitmix64 finaliser, so the same spec can be restated on the device
(`csrc/synth.cu`) and the two can be compared byte for byte.

File formats follow what the reference's parsers consume:
  perf.script   -> sofa_preprocess.py:110-154  (cpu_trace_read)
  gputrace.tmp  -> sofa_preprocess.py:249-341  (gpu_trace_read), header/unit
                   rows :1385-1417
  net.tmp       -> sofa_preprocess.py:156-201  (net_trace_read)
side files: misc.txt (:387-392), sofa_time.txt (:418-422),
perf_events_used.txt (:1545-1551), perf_timebase.txt (:1766-1784).
"""
import os

M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15

STREAM_PERF = 1
STREAM_GPU = 2
STREAM_NET = 3

TIME_BASE = 1600000000.0          # sofa_time.txt
PERF_T0_US = 5000 * 1000000       # perf clock origin (uptime seconds)
PERF_GAP_US = 17
GPU_GAP_US = 23
NET_T0_US = 1600000000 * 1000000 + 250000
NET_GAP_US = 211

WORDS = ["do_syscall_64", "schedule", "page_fault", "tcp_sendmsg",
         "PyEval_EvalFrameDefault", "cudaLaunchKernel", "ioctl", "futex_wait",
         "copy_user_enhanced_fast_string", "__libc_start_main",
         "pthread_mutex_lock", "vfs_read"]
LIBS = ["c", "pthread", "cuda", "cudart", "python3.6m", "tensorflow_framework",
        "stdc++", "nccl"]
KERNELS = ["volta_sgemm_128x64_nn", "volta_scudnn_128x64_relu_interior_nn_v1",
           "maxwell_gcgemm_32x32_nt", "ncclAllReduceKernel_sum_f32",
           "void tensorflow::functor::FillPhiloxRandomKernelLaunch<float, int=4>(float*, long)",
           "void cudnn::detail::implicit_convolve_sgemm<float, float, int=128, int=6>(int, int)",
           "cudnn_fw_conv_kernel", "cudnn_bw_data_kernel",
           "void Eigen::internal::EigenMetaKernel<Eigen::TensorEvaluator>(float)",
           "AllReduceKernel_ring_ll", "gemv2T_kernel_val", "relu_fw_elementwise"]
NVPROF_HEADER = ('"Start","Duration","Grid X","Grid Y","Grid Z","Block X","Block Y",'
                 '"Block Z","Registers Per Thread","Static SMem","Dynamic SMem","Size",'
                 '"Throughput","SrcMemType","DstMemType","Device","Context","Stream","Name"')
NVPROF_HEADER_P2P = ('"Start","Duration","Grid X","Grid Y","Grid Z","Block X","Block Y",'
                     '"Block Z","Registers Per Thread","Static SMem","Dynamic SMem","Size",'
                     '"Throughput","SrcMemType","DstMemType","Device","Context","Stream",'
                     '"Src Dev","Src Ctx","Dst Dev","Dst Ctx","Name"')


def mix64(x):
    x &= M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def rnd(seed, stream, i, f):
    """64 random bits for (seed, stream, row i, field f)."""
    return mix64(seed + ((stream << 58) ^ ((i * 16 + f + 1) * GOLDEN)))


def perf_symbol(k):
    """Symbol text for dictionary slot k in [0, 4096). Returns (sym, has_off)."""
    t = k % 16
    if t == 0:
        return "[unknown]", False
    if t == 1:
        return "cpu_idle_poll_%d" % k, True
    if t == 2:
        nm = "Op%dKernel" % k
        return "_ZN10tensorflow%d%s7ComputeEv" % (len(nm), nm), True
    if t == 3:
        nm = "gemm_pack_%d" % k
        return "_ZN5Eigen8internal%d%s3runEv" % (len(nm), nm), True
    if t == 4:
        return "memcpy_%d@plt" % k, True
    return "%s_%d" % (WORDS[k % 12], k), True


def perf_line(seed, i, cpu_col=False):
    r0 = rnd(seed, STREAM_PERF, i, 0)
    r1 = rnd(seed, STREAM_PERF, i, 1)
    r2 = rnd(seed, STREAM_PERF, i, 2)
    r3 = rnd(seed, STREAM_PERF, i, 3)
    pid = 4242
    tid = 4242 + (((r0 & 63) * ((r0 >> 6) & 63)) >> 6)
    us = PERF_T0_US + i * PERF_GAP_US + (r0 >> 12) % PERF_GAP_US
    period = 1 + (r1 % 3000000)
    kernel_ip = (r2 % 10) < 3
    if kernel_ip:
        ip = 0xffffffff80000000 | ((r2 >> 8) & 0x0fffffff)
    else:
        ip = 0x7f0000000000 | ((r2 >> 8) & 0xffffffffff)
    r20 = r3 & 0xfffff
    k = (r20 * r20 * r20) >> 48
    sym, has_off = perf_symbol(k)
    if has_off:
        sym = "%s+0x%x" % (sym, (r3 >> 20) & 0xfff)
    if kernel_ip:
        dso = "([kernel.kallsyms])"
    else:
        j = (r3 >> 32) & 63
        dso = "(/usr/lib/x86_64-linux-gnu/lib%s.so.%d)" % (LIBS[j & 7], j)
    cpu = " [%03d]" % ((r1 >> 32) & 63) if cpu_col else ""
    return "%5d/%-5d%s %7d.%06d: %10d cycles:  %16x %s %s\n" % (
        pid, tid, cpu, us // 1000000, us % 1000000, period, ip, sym, dso)


def gpu_header(units="s_ms", p2p_cols=False):
    """Three header rows of `nvprof --csv --print-gpu-trace` (banner, names, units)."""
    hdr = NVPROF_HEADER_P2P if p2p_cols else NVPROF_HEADER
    ncol = hdr.count(',') + 1
    if units == "s_ms":
        u = ["s", "ms", "", "", "", "", "", "", "", "B", "B", "MB", "GB/s"]
    elif units == "ms_us_KB":
        u = ["ms", "us", "", "", "", "", "", "", "", "B", "B", "KB", "GB/s"]
    elif units == "us_us_B":
        u = ["us", "us", "", "", "", "", "", "", "", "B", "B", "B", "GB/s"]
    else:
        raise ValueError(units)
    u = u + [""] * (ncol - len(u))
    return "==4242== Profiling result:\n" + hdr + "\n" + ",".join(u) + "\n"


def gpu_line(seed, i, units="s_ms", p2p_cols=False):
    r0 = rnd(seed, STREAM_GPU, i, 0)
    r1 = rnd(seed, STREAM_GPU, i, 1)
    r2 = rnd(seed, STREAM_GPU, i, 2)
    us = 350000 + i * GPU_GAP_US + (r0 >> 12) % GPU_GAP_US   # since CUPTI start
    dur_ns = 800 + (r1 % 4000) * (1 + ((r1 >> 16) & 255))
    if units == "s_ms":
        start = "%d.%06d" % (us // 1000000, us % 1000000)
        dur = "%d.%06d" % (dur_ns // 1000000, dur_ns % 1000000)
    elif units == "ms_us_KB":
        start = "%d.%03d" % (us // 1000, us % 1000)
        dur = "%d.%03d" % (dur_ns // 1000, dur_ns % 1000)
    else:
        start = "%d.000000" % us
        dur = "%d.%03d" % (dur_ns // 1000, dur_ns % 1000)
    ctx = 1 + ((r0 >> 4) & 7)
    stream = 7 + ((r0 >> 7) & 31)
    dev = '"Tesla V100-SXM2-16GB (%d)"' % (ctx - 1)
    kind = r2 % 100
    p2p = ",,,," if p2p_cols else ""
    if kind < 85:
        k = (r2 >> 8) % 512
        base = KERNELS[k % 12]
        name = base if k < 12 else "%s_v%d" % (base, k)
        gx = 1 + ((r2 >> 20) & 1023)
        bx = 32 << ((r2 >> 30) & 3)
        regs = 16 + ((r2 >> 32) & 63)
        return '%s,%s,%d,1,1,%d,1,1,%d,0,0,,,,,%s,"%d","%d",%s"%s"\n' % (
            start, dur, gx, bx, regs, dev, ctx, stream, p2p, name)
    nbytes = 4 << ((r2 >> 8) % 24)
    if units == "s_ms":
        size = "%d.%06d" % (nbytes >> 20, ((nbytes & 0xfffff) * 1000000) >> 20)
    elif units == "ms_us_KB":
        size = "%d.%06d" % (nbytes >> 10, ((nbytes & 0x3ff) * 1000000) >> 10)
    else:
        size = "%d" % nbytes
    thr = 1000 + (r2 >> 40) % 24000000
    thr_s = "%d.%06d" % (thr // 1000000, thr % 1000000)
    if kind < 91:
        name, sm, dm = "[CUDA memcpy HtoD]", "Pageable", "Device"
    elif kind < 96:
        name, sm, dm = "[CUDA memcpy DtoH]", "Device", "Pageable"
    elif kind < 98:
        name, sm, dm = "[CUDA memcpy DtoD]", "Device", "Device"
    else:
        name, sm, dm = "[CUDA memcpy PtoP]", "Device", "Device"
        if p2p_cols:
            sc = 1 + ((r2 >> 16) & 7)
            dc = 1 + ((r2 >> 19) & 7)
            p2p = '"dev (%d)","%d","dev (%d)","%d",' % (sc - 1, sc, dc - 1, dc)
    return '%s,%s,,,,,,,,,,%s,%s,"%s","%s",%s,"%d","%d",%s"%s"\n' % (
        start, dur, size, thr_s, sm, dm, dev, ctx, stream, p2p, name)


def net_line(seed, i, non_ip_pct=1):
    r0 = rnd(seed, STREAM_NET, i, 0)
    r1 = rnd(seed, STREAM_NET, i, 1)
    us = NET_T0_US + i * NET_GAP_US + (r0 >> 12) % NET_GAP_US
    ts = "%d.%06d" % (us // 1000000, us % 1000000)
    sel = r1 % 100
    if sel < non_ip_pct:
        if (r1 >> 8) & 1:
            return "%s ARP, Request who-has 10.0.0.%d tell 10.0.0.%d, length 28\n" % (
                ts, 1 + ((r1 >> 16) & 15), 1 + ((r1 >> 20) & 15))
        return "%s IP 10.0.0.%d > 10.0.0.%d: ICMP echo request, id 7, seq %d, length 64\n" % (
            ts, 1 + ((r1 >> 16) & 15), 1 + ((r1 >> 20) & 15), (r1 >> 24) & 1023)
    src = "10.0.%d.%d" % ((r0 >> 4) & 1, 1 + ((r0 >> 5) & 15))
    dst = "10.0.%d.%d" % ((r0 >> 9) & 1, 1 + ((r0 >> 10) & 15))
    sport = 1024 + ((r1 >> 8) & 0x7fff)
    dport = 1024 + ((r1 >> 24) & 0x7fff)
    if (r1 >> 40) % 4 == 0:
        return "%s IP %s.%d > %s.%d: UDP, length %d\n" % (
            ts, src, sport, dst, dport, (r1 >> 44) % 1473)
    return "%s IP %s.%d > %s.%d: tcp %d\n" % (
        ts, src, sport, dst, dport, (r1 >> 44) % 65536)


API_NAMES = ["cudaLaunchKernel", "cudaMemcpyAsync", "cudaStreamSynchronize", "cudaMalloc", "cudaFree",
             "cudaEventRecord", "cuLaunchKernel", "cudaMemsetAsync"]
STREAM_API = 4


def api_header(with_corid=True):
    cols = '"Start","Duration","Name"' + (',"Correlation_ID"' if with_corid else '')
    units = "ms,us," + ("," if with_corid else "")
    return "==4242== Profiling result:\n" + cols + "\n" + units + "\n"


def api_line(seed, i, with_corid=True):
    r0 = rnd(seed, STREAM_API, i, 0)
    us = 349000 + i * 19 + (r0 >> 12) % 19
    dur_ns = 500 + (r0 >> 20) % 90000
    name = API_NAMES[(r0 >> 4) & 7]
    row = '%d.%03d,%d.%03d,"%s"' % (us // 1000, us % 1000, dur_ns // 1000, dur_ns % 1000, name)
    return row + (",%d\n" % (i + 1) if with_corid else "\n")


def api_text(seed, n, with_corid=True):
    return api_header(with_corid) + "".join(api_line(seed, i, with_corid) for i in range(n))


def perf_text(seed, n, cpu_col=False):
    return "".join(perf_line(seed, i, cpu_col) for i in range(n))


def gpu_text(seed, n, units="s_ms", p2p_cols=False):
    return gpu_header(units, p2p_cols) + "".join(
        gpu_line(seed, i, units, p2p_cols) for i in range(n))


def net_text(seed, n, non_ip_pct=1):
    return "".join(net_line(seed, i, non_ip_pct) for i in range(n))


def write_side_files(logdir, elapsed=12.5, perf_events="cycles", cpuinfo=None):
    """misc.txt / sofa_time.txt / perf_events_used.txt / perf_timebase.txt as the
    recorder writes them (sofa_record.py:236-247, 356-357, 416-431)."""
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "misc.txt"), "w") as f:
        f.write("elapsed_time %.6f\ncores 64\nvcores 64\npid 4242\n" % elapsed)
    with open(os.path.join(logdir, "sofa_time.txt"), "w") as f:
        f.write("%.1f\n" % TIME_BASE)
    with open(os.path.join(logdir, "perf_events_used.txt"), "w") as f:
        f.write(perf_events)
    with open(os.path.join(logdir, "perf_timebase.txt"), "w") as f:
        # last two lines: a perf-script sample (uptime clock) and a unix timestamp
        f.write("[ perf record: Woken up 1 times to write data ]\n")
        f.write("[ perf record: Captured and wrote 0.012 MB perf.data (8 samples) ]\n")
        f.write("ls  4300 4999.500000:     1 cycles:  ffffffff8106315a native_write_msr_safe\n")
        f.write("1600000000.250000 0\n")
    if cpuinfo:
        with open(os.path.join(logdir, "cpuinfo.txt"), "w") as f:
            for t, mhz in cpuinfo:
                f.write("%.6f %.3f\n" % (t, mhz))


def make_logdir(logdir, seed, n_perf=0, n_gpu=0, n_net=0, cpu_col=False,
                units="s_ms", p2p_cols=False, non_ip_pct=1, perf_events="cycles",
                cpuinfo=None, n_api=0, api_corid=True):
    """Write a complete synthetic logdir. GPU / net sources are written as the
    text dumps (`gputrace.tmp`, `net.tmp`) plus the marker files the reference
    globs for (`gputrace1.nvvp`, `sofa.pcap`)."""
    write_side_files(logdir, perf_events=perf_events, cpuinfo=cpuinfo)
    if n_perf:
        with open(os.path.join(logdir, "perf.script"), "w") as f:
            f.write(perf_text(seed, n_perf, cpu_col))
    if n_gpu:
        with open(os.path.join(logdir, "gputrace.tmp"), "w") as f:
            f.write(gpu_text(seed, n_gpu, units, p2p_cols))
        open(os.path.join(logdir, "gputrace1.nvvp"), "w").close()
    if n_api:
        with open(os.path.join(logdir, "cuda_api_trace.tmp"), "w") as f:
            f.write(api_text(seed, n_api, api_corid))
        open(os.path.join(logdir, "gputrace1.nvvp"), "w").close()
    if n_net:
        with open(os.path.join(logdir, "net.tmp"), "w") as f:
            f.write(net_text(seed, n_net, non_ip_pct))
        open(os.path.join(logdir, "sofa.pcap"), "w").close()


# ----------------------------------------------------------------------------- side sources
# Raw side logs as the recorder writes them (sofa_record.py:62-135, 303-310), consumed by the
# reference's preprocess (sofa_preprocess.py:482-546 mpstat, 1013-1098 nvsmi_query, 1235-1337
# netstat).  Used to produce mpstat.csv / nvsmi_trace.csv / netstat.csv inputs for
# concurrency_breakdown (sofa_analyze.py:75-243).
def write_side_raw(logdir, seed, n_samples=60, n_cores=4, n_gpus=2, rate_hz=10):
    import datetime
    os.makedirs(logdir, exist_ok=True)
    t0 = TIME_BASE + 0.5
    dt = 1.0 / rate_hz
    # mpstat.txt: header + cumulative jiffies per (sample, cpu); cpu -1 = all
    cum = [[0] * 7 for _ in range(n_cores + 1)]
    with open(os.path.join(logdir, "mpstat.txt"), "w") as f:
        f.write("time,cpu,user,nice,system,idle,iowait,irq,softirq\n")
        for s in range(n_samples):
            t = t0 + s * dt
            for ci, cpu in enumerate([-1] + list(range(n_cores))):
                r = rnd(seed, 5, s * 64 + ci, 0)
                phase = (s // 15) % 3                      # busy / gpu-bound / idle phases
                usr = (r & 63) + (40 if phase == 0 else 3)
                sysj = ((r >> 6) & 15) + (10 if phase == 0 else 1)
                iow = ((r >> 10) & 7) if phase != 2 else 0
                irq = (r >> 13) & 3
                idl = 100 - min(99, usr + sysj + iow + irq) + ((r >> 15) & 7)
                inc = [usr, 0, sysj, idl, iow, irq, (r >> 18) & 1]
                for k in range(7):
                    cum[ci][k] += inc[k]
                f.write("%.6f,%d,%s\n" % (t, cpu, ",".join(str(v) for v in cum[ci])))
    # nvsmi_query.txt: first line skipped, "timestamp, name, index, sm %, mem %"
    with open(os.path.join(logdir, "nvsmi_query.txt"), "w") as f:
        f.write("timestamp, name, index, utilization.gpu [%], utilization.memory [%]\n")
        for s in range(n_samples):
            t = t0 + 0.013 + s * dt
            ts = datetime.datetime.fromtimestamp(t).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
            for g in range(n_gpus):
                r = rnd(seed, 6, s * 16 + g, 0)
                phase = (s // 15) % 3
                sm = (r % 30) + (65 if phase == 1 else (20 if phase == 0 else 0))
                mem = (r >> 8) % 60
                f.write("%s, Tesla V100-SXM2-16GB, %d, %d %%, %d %%\n" % (ts, g, sm, mem))
    # netstat.txt: time,tx_bytes,rx_bytes (cumulative)
    tx = rx = 0
    with open(os.path.join(logdir, "netstat.txt"), "w") as f:
        for s in range(n_samples + 1):
            t = t0 - dt + 0.031 + s * dt
            r = rnd(seed, 7, s, 0)
            tx += r % 1000000
            rx += (r >> 20) % 3000000
            f.write("%.6f,%d,%d\n" % (t, tx, rx))


def write_side_raw_1hz(logdir, seed, n_seconds=30, n_gpus=2):
    """vmstat.txt (`vmstat -w 1`) and nvsmi.txt (`nvidia-smi dmon -s u`) as the recorder captures them
    (sofa_record.py:249-251, 300-302): one sample per second, headers repeated every 20 samples."""
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "vmstat.txt"), "w") as f:
        for s in range(n_seconds):
            if s % 20 == 0:
                f.write("procs -----------------------memory---------------------- ---swap-- -----io---- -system-- --------cpu--------\n")
                f.write(" r  b         swpd         free         buff        cache   si   so    bi    bo   in   cs  us  sy  id  wa  st\n")
            r = rnd(seed, 10, s, 0)
            us, sy = r % 60, (r >> 6) % 20
            wa = (r >> 11) % 5
            f.write("%2d %2d %12d %12d %12d %12d %4d %4d %5d %5d %4d %4d %3d %3d %3d %3d %3d\n" % (
                (r >> 14) % 9, (r >> 18) % 3, 0, 100000000 + (r >> 20) % 5000000, 1234567, 23456789 + s * 1000,
                0, 0, (r >> 24) % 5000, (r >> 36) % 90000, 100 + (r >> 44) % 9000, 200 + (r >> 50) % 9000,
                us, sy, 100 - us - sy - wa, wa, 0))
        f.write("  3  0\n")                                 # a truncated last line (recorder killed mid-write)
    with open(os.path.join(logdir, "nvsmi.txt"), "w") as f:
        for s in range(n_seconds):
            if s % 20 == 0:
                f.write("# gpu    sm   mem   enc   dec\n# Idx     %     %     %     %\n")
            for g in range(n_gpus):
                r = rnd(seed, 11, s * 16 + g, 0)
                enc = "-" if (r >> 20) % 7 == 0 else str(r % 40)
                dec = "-" if (r >> 24) % 5 == 0 else str((r >> 8) % 30)
                f.write("%5d %5d %5d %5s %5s\n" % (g, (r >> 12) % 100, (r >> 30) % 100, enc, dec))


def write_diskstat(logdir, seed, n_samples=40, devs=("sda", "sda1", "nvme0n1", "loop0", "dm-0"), rate_hz=10):
    """diskstat.txt as the recorder appends it (sofa_record.py:96-121): one line per block device and sample,
    `time,dev,sectors_read,sectors_written,reads,writes,ms_reading,ms_writing` with cumulative counters.
    loop0 never moves (the reference drops devices whose traffic stays zero)."""
    os.makedirs(logdir, exist_ok=True)
    t0 = TIME_BASE + 0.5
    cum = {d: [1000 * (i + 1), 2000 * (i + 1), 10 * (i + 1), 20 * (i + 1), 5 * (i + 1), 7 * (i + 1)] for i, d in enumerate(devs)}
    with open(os.path.join(logdir, "diskstat.txt"), "w") as f:
        for s_ in range(n_samples):
            t = t0 + s_ / rate_hz + 0.0007 * (s_ % 3)
            for i, d in enumerate(devs):
                r = rnd(seed, 12, s_ * 16 + i, 0)
                if d != "loop0":
                    busy = (s_ // 10 + i) % 3
                    rd = (r % 4000) * (busy != 2)
                    wr = ((r >> 12) % 9000) * (busy != 0)
                    rio, wio = (rd + 7) // 8 if rd else 0, (wr + 15) // 16 if wr else 0
                    inc = [rd, wr, rio, wio, ((r >> 24) % 50) if rio else 0, ((r >> 30) % 120) if wio else 0]
                    for k in range(6):
                        cum[d][k] += inc[k]
                f.write("%.6f,%s,%s\n" % (t, d, ",".join(str(v) for v in cum[d])))


def write_strace_blktrace(logdir, seed, n_strace=300, n_blk=200):
    """strace.txt as `strace -q -T -tt(t) -f -o` writes it (pid prefix, epoch stamps, `<duration>` suffix) and
    blktrace.txt as `blkparse` prints it (sofa_record.py:337, 444); both opt-in recorder outputs."""
    os.makedirs(logdir, exist_ok=True)
    calls = ["read(3, \"abc\\n\", 4096) = 4", "write(1, \"x, y\", 4) = 4", "mmap(NULL, 262144, PROT_READ|PROT_WRITE, MAP_PRIVATE|MAP_ANONYMOUS, -1, 0) = 0x7f501f910000",
             "futex(0x7f50, FUTEX_WAIT_PRIVATE, 2, NULL) = 0", "openat(AT_FDCWD, \"/etc/ld.so.cache\", O_RDONLY|O_CLOEXEC) = 3", "ioctl(4, TCGETS, 0x7ffd) = -1 ENOTTY (Inappropriate ioctl for device)",
             "close(3) = 0", "nanosleep({tv_sec=0, tv_nsec=1000}, NULL) = 0", "poll([{fd=5, events=POLLIN}], 1, 100) = 1 ([{fd=5, revents=POLLIN}])"]
    t = TIME_BASE + 0.4
    with open(os.path.join(logdir, "strace.txt"), "w") as f:
        for i in range(n_strace):
            r = rnd(seed, 13, i, 0)
            t += (r % 5000) / 1e6
            pid = 4242 + (r >> 13) % 3
            call = calls[(r >> 16) % len(calls)]
            dur = ((r >> 24) % 200000) / 1e6
            if (r >> 50) % 23 == 0:
                f.write("%d %.6f +++ exited with 0 +++\n" % (pid, t))
            elif (r >> 45) % 17 == 0:
                f.write("[pid %5d] %.6f %s <%.6f>\n" % (pid, t, call, dur))
            elif (r >> 40) % 19 == 0:
                f.write("%d %.6f %s <unfinished ...>\n" % (pid, t, call.split(" = ")[0]))
            else:
                f.write("%d %.6f %s <%.6f>\n" % (pid, t, call, dur))
    with open(os.path.join(logdir, "blktrace.txt"), "w") as f:
        tb = 0.0
        seq = 0
        pending = []
        for i in range(n_blk):
            r = rnd(seed, 14, i, 0)
            tb += (r % 3000) / 1e6
            seq += 1
            cpu = (r >> 12) % 8
            blk = 691248304 + 1024 * ((r >> 16) % 40)
            if pending and (r >> 30) % 3 == 0:
                b0 = pending.pop(0)
                f.write("  8,0  %3d %8d %14.9f %5d  C   W %d + 1024 [0]\n" % (cpu, seq, tb, 0, b0))
            else:
                ev = "QGID"[(r >> 34) % 4]
                f.write("  8,0  %3d %8d %14.9f %5d  %s   W %d + 1024 [dd]\n" % (cpu, seq, tb, 31479, ev, blk))
                if ev == "D":
                    pending.append(blk)
        f.write("CPU0 (8,0):\n Reads Queued:           0,        0KiB\t Writes Queued:         100,    51200KiB\n")
        f.write("Total (8,0):\n IO unplugs:             5        \t Timer unplugs:           0\n")


def write_pystacks(logdir, seed, n=60):
    """pystacks.txt as `pyflame --flamechart -t` writes it: a microsecond time stamp line, then the stack line."""
    frames = ["train.py:main:10", "model.py:forward:55", "ops.py:conv2d:120", "data.py:next_batch:33", "(idle)", "optim.py:step:77"]
    t = int((TIME_BASE + 0.45) * 1e6)
    with open(os.path.join(logdir, "pystacks.txt"), "w") as f:
        for i in range(n):
            r = rnd(seed, 15, i, 0)
            t += 1000 + r % 50000
            depth = 1 + (r >> 16) % 4
            stack = ";".join(frames[(r >> (20 + 3 * k)) % len(frames)] for k in range(depth))
            f.write("%d\n%s\n" % (t, stack))


def side_csv_bodies(seed, n_samples, n_cores, n_gpus, rate_hz=10):
    """Header-less bodies of mpstat.csv, nvsmi_trace.csv, netstat.csv in the format the reference's
    preprocess writes them (tests/golden/side_small), for concurrency_breakdown at scale."""
    dt = 1.0 / rate_hz
    mp, nv, ns = [], [], []
    for s_ in range(n_samples):
        t = 0.6 + s_ * dt
        for c in range(n_cores):
            r = rnd(seed, 8, s_ * 256 + c, 0)
            mp.append("%.6f,-1,%.6f,%.6f,-1,-1,-1,-1,-1,-1,-1,mpstat_core%d (usr|sys|idl|iow|irq): |%3d|%3d|%3d|%3d|%3d|,0\n"
                      % (t, ((r >> 40) % 1000000) / 1e6, c, c, r % 61, (r >> 6) % 26, (r >> 12) % 100, (r >> 20) % 10, (r >> 24) % 4))
    for ev, tag in ((0, "sm"), (1, "mem")):
        for s_ in range(n_samples):
            for g in range(n_gpus):
                v = rnd(seed, 9, (s_ * 16 + g) * 2 + ev, 0) % 101
                nv.append("%.6f,%d,%d,%d,-1,-1,-1,-1,-1,-1,-1,GPUID_%s=%d_%d,%d\n" % (0.513 + s_ * dt, ev, v, g, tag, g, v, g))
    for ev, tag in ((0, "tx"), (1, "rx")):
        for s_ in range(n_samples):
            bw = (rnd(seed, 10, s_ * 2 + ev, 0) % 10**13) / 1e6
            ns.append("%.6f,%d,-1,-1,-1,-1,%.6f,-1,-1,-1,-1,network_bandwidth_%s(bytes):%d,0\n" % (0.531 + s_ * dt, ev, bw, tag, int(bw)))
    return "".join(mp).encode(), "".join(nv).encode(), "".join(ns).encode()
