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
        if os.path.isfile(p) and not name.startswith(("side_", "disk_", "oracle_", "driver_", "aisi_")):      # side-file fixtures (disk_* / oracle_*: oracle only)
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
               "perf_timebase.txt", "misc.txt", "vmstat.txt", "nvsmi.txt")          # the last two only in some cases
SIDE_OUTPUTS = ("mpstat.csv", "usr_sys.csv", "nvsmi_trace.csv", "netstat.csv", "netbandwidth.csv", "cputrace.csv",
                "report.js", "vmstat.csv")
PREPROCESS_CFG_KEYS = ("absolute_timestamp", "enable_vmstat", "enable_encode_decode", "nvsmi_time_zone")


def preprocess_overrides(case):
    return {k: v for k, v in case["cfg"].items() if k in PREPROCESS_CFG_KEYS}


def side_cases():
    return sorted(d for d in os.listdir(GOLDEN) if d.startswith("side_"))


def build_side_logdir(name, tmpdir):
    """Copy the stored raw inputs of a side_* golden case into a fresh logdir; pin the time zone the
    fixtures were generated under (the reference reads nvidia-smi stamps as local time)."""
    import shutil
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    logdir = os.path.join(str(tmpdir), "sofalog") + "/"
    os.makedirs(logdir)
    for fn in SIDE_INPUTS:
        if os.path.isfile(os.path.join(GOLDEN, name, fn)):
            shutil.copy(os.path.join(GOLDEN, name, fn), logdir + fn)
    return load_case(name), logdir
