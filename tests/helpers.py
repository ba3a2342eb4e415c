"""Shared test helpers: golden cases, oracle runner, libstdc++ demangler."""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from sofa_b200 import synth  # noqa: E402


def golden_cases():
    out = []
    for name in sorted(os.listdir(GOLDEN)):
        p = os.path.join(GOLDEN, name, "case.json")
        if os.path.isfile(p) and not name.startswith(("side_", "disk_", "oracle_", "driver_", "aisi_")):      # fixtures of other test families (side files, driver stdout, AISI)
            out.append(name)
    return out


def load_case(name):
    with open(os.path.join(GOLDEN, name, "case.json")) as f:
        return json.load(f)


def golden_bytes(name, fn):
    p = os.path.join(GOLDEN, name, fn)
    if not os.path.isfile(p):
        return None
    with open(p, "rb") as f:
        return f.read()


def demangle(name):
    """libstdc++ __cxa_demangle, i.e. what cxxfilt.demangle wraps."""
    if not name.startswith("_Z"):
        return name
    lib = ctypes.CDLL("libstdc++.so.6")
    libc = ctypes.CDLL("libc.so.6")
    lib.__cxa_demangle.restype = ctypes.c_void_p
    status = ctypes.c_int(0)
    p = lib.__cxa_demangle(name.encode(), None, None, ctypes.byref(status))
    if status.value != 0 or not p:
        raise ValueError("InvalidName: %s" % name)
    s = ctypes.string_at(p).decode()
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(p)
    return s


def build_case_logdir(name, tmpdir):
    case = load_case(name)
    logdir = os.path.join(str(tmpdir), "sofalog") + "/"
    synth.make_logdir(logdir, **case["make_logdir"])
    return case, logdir


SIDE_INPUTS = ("mpstat.txt", "nvsmi_query.txt", "netstat.txt", "perf.script", "sofa_time.txt", "perf_events_used.txt",
               "perf_timebase.txt", "misc.txt", "vmstat.txt", "nvsmi.txt",          # the last two only in some cases
               "diskstat.txt", "strace.txt", "blktrace.txt", "pystacks.txt")         # disk_small / oracle_strace_blk
SIDE_OUTPUTS = ("mpstat.csv", "usr_sys.csv", "nvsmi_trace.csv", "netstat.csv", "netbandwidth.csv", "cputrace.csv",
                "report.js", "vmstat.csv", "diskstat.csv", "diskstat_vector.csv", "diskstat_vector_ui.csv", "strace.csv",
                "blktrace.csv", "pystacks.csv")
PREPROCESS_CFG_KEYS = ("absolute_timestamp", "enable_vmstat", "enable_encode_decode", "nvsmi_time_zone", "plot_ratio")


def preprocess_overrides(case):
    return {k: v for k, v in case["cfg"].items() if k in PREPROCESS_CFG_KEYS}


def side_cases():
    """side_*: mpstat / vmstat / nvidia-smi / netstat; disk_small: + diskstat.txt; oracle_strace_blk: + strace.txt,
    blktrace.txt, pystacks.txt (all of them files written by the unmodified reference, make_golden_side.py)"""
    return sorted(d for d in os.listdir(GOLDEN) if d.startswith("side_")) + ["disk_small", "oracle_strace_blk"]


def build_side_logdir(name, tmpdir):
    """Copy the stored raw inputs of a side_* golden case into a fresh logdir; pin the time zone the
    fixtures were generated under (the reference reads nvidia-smi stamps as local time)."""
    import shutil
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    import sofa_b200.sofa_preprocess as sp
    sp.SYS_BLOCK = "/nonexistent_sofa_b200/"        # diskstat sector sizes: the fixtures were made where none of the names exist
    logdir = os.path.join(str(tmpdir), "sofalog") + "/"
    os.makedirs(logdir)
    for fn in SIDE_INPUTS:
        if os.path.isfile(os.path.join(GOLDEN, name, fn)):
            shutil.copy(os.path.join(GOLDEN, name, fn), logdir + fn)
    return load_case(name), logdir


STORAGE_OUTPUTS = ("diskstat.csv", "diskstat_vector.csv", "diskstat_vector_ui.csv", "blktrace.csv", "strace.csv", "pystacks.csv", "report.js")


def make_storage_logdir(a, n_disk=3000, n_strace=20000, n_blk=30000, n_py=5000):
    """A logdir whose diskstat / strace / blktrace / pystacks files span many thread blocks, plus the odd lines the
    reference tolerates: a completion without a dispatch, a start block that is not a number, a short line."""
    import time
    import sofa_b200.sofa_preprocess as sp
    sp.SYS_BLOCK = "/nonexistent_sofa_b200/"
    os.environ['TZ'] = 'UTC'
    time.tzset()
    synth.make_logdir(a, seed=31, n_perf=300)
    synth.write_side_raw(a, seed=32, n_samples=40, n_cores=2, n_gpus=2)
    synth.write_side_raw_1hz(a, seed=33, n_seconds=8, n_gpus=2)
    synth.write_diskstat(a, 34, n_samples=n_disk, devs=("xda", "xda1", "fake0n1", "loop0", "dm-x", "y,z".replace(",", "_")))
    synth.write_strace_blktrace(a, 35, n_strace=n_strace, n_blk=n_blk)
    synth.write_pystacks(a, 36, n=n_py)
    lines = open(a + "blktrace.txt").read().split("\n")
    lines.insert(50, "  8,0    2      999    0.000300000     0  C   W 123456789 + 1024 [0]")          # no dispatch to complete
    lines.insert(90, "  8,0    5     1000    0.000400000 31479  D   W notanumber + 1024 [dd]")      # -> start block 0, the warning
    lines.insert(95, "  8,0    5     1001    0.000500000     0  C   W alsonot + 8 [0] [0] [0] [0]")      # completes block 0
    lines.insert(120, "  8,0    1 short line")
    open(a + "blktrace.txt", "w").write("\n".join(lines))


def make_full_logdir(a, seed=0):
    """perf + nvprof + tcpdump text and every recorder side file in one small logdir (config-space fuzz)"""
    make_storage_logdir(a, n_disk=60, n_strace=300, n_blk=400, n_py=80)
    synth.make_logdir(a, seed=31 + seed, n_perf=300, n_gpu=200, n_net=150, non_ip_pct=3)
    synth.write_side_raw(a, seed=32, n_samples=40, n_cores=2, n_gpus=2)
    synth.write_side_raw_1hz(a, seed=33, n_seconds=8, n_gpus=2)


def random_overrides(rng):
    """a random point of the report options that reach sofa_preprocess (bin/sofa:161-312)"""
    over = {}
    if rng.random() < 0.5: over['plot_ratio'] = rng.choice([2, 3, 5, 11])
    if rng.random() < 0.4: over['absolute_timestamp'] = True
    if rng.random() < 0.4: over['cpu_time_offset'] = rng.choice([0.001, -0.25, 3.5])
    if rng.random() < 0.4: over['nvsmi_time_zone'] = rng.choice([-8, 3, 9])
    if rng.random() < 0.5: over['enable_vmstat'] = True
    if rng.random() < 0.5: over['enable_encode_decode'] = True
    if rng.random() < 0.3: over['strace_min_time'] = rng.choice([0.0, 0.01, 0.1])
    if rng.random() < 0.3: over['diskstat_filters'] = rng.choice(['xda', 'dm-x', 'nosuch'])
    if rng.random() < 0.6: over['cpu_filters'] = rng.choice(['idle:black', 'idle:black,tensorflow:orange', 'memcpy:red,nccl:blue,page_fault:green'])
    return over


ALL_PREPROCESS_OUTPUTS = tuple(sorted(set(STORAGE_OUTPUTS) | set(SIDE_OUTPUTS) | {'gputrace.csv', 'nettrace.csv'}))
