"""One process per GPU, one trace (logdir) per rank - the host logic of the N>1 path.

The preprocess path shards by trace with no data-path exchange: every rank
runs the same `sofa_preprocess` / `sofa_analyze` on its own logdirs (the reference's cluster
mode loops `cfg.logdir = <logdir>-<ip>/` serially, bin/sofa:358-365).  torch.distributed is
used only for the barrier and for reducing the step time (max) and event counts (sum); the
backend is NCCL on GPUs and gloo in the CPU tests.
"""
import os


def rank_env():
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def assign_logdirs(logdirs, rank, world):
    """Round-robin assignment of per-node logdirs (`sofalog-<ip>/`) to ranks, in list order."""
    return [d for i, d in enumerate(logdirs) if i % world == rank]


def cluster_logdirs(base_logdir, cluster_ip):
    """bin/sofa:358-365: '<logdir>-<ip>/' for every ip of --cluster_ip."""
    base = base_logdir.rstrip('/')
    return ["%s-%s/" % (base, ip) for ip in cluster_ip.split(',') if ip]


def aggregate(dist, events_local, ms_local, device=None):
    """Whole-job throughput: sum of events over ranks / max step time over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return events_local, ms_local, events_local / (ms_local * 1e-3)
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    e = torch.tensor([float(events_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    return e.item(), t.item(), e.item() / (t.item() * 1e-3)


def preprocess_cluster(cfg, logdirs, engine_factory=None):
    """Run sofa_preprocess over this rank's logdirs (one engine, sources re-uploaded per logdir)."""
    from .sofa_preprocess import sofa_preprocess
    base = cfg.logdir
    eng = engine_factory() if engine_factory else None
    try:
        for d in logdirs:
            cfg.logdir = d
            sofa_preprocess(cfg, engine=eng)
    finally:
        cfg.logdir = base
        if eng is not None:
            eng.close()


# ----------------------------------------------------------------------------- NUMA placement of a rank
def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(gpu_index):
    """NUMA node of GPU `gpu_index` (sysfs numa_node of its PCI function), or None when unknown."""
    import subprocess
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id",
                                       "--format=csv,noheader"], text=True, stderr=subprocess.DEVNULL).strip()
    except Exception:
        return None
    if not bus:
        return None
    dom, rest = bus.split(':', 1)
    path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
    try:
        with open(path) as f:
            node = int(f.read().strip())
    except Exception:
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa(gpu_index, world=1):
    """Pin this process (its threads and, through first touch + MPOL_PREFERRED, the pinned host arenas it allocates
    afterwards) to the NUMA node GPU `gpu_index` hangs off.  Eight ranks each moving tens of GB per step through
    host memory otherwise share whatever node the launcher left them on: round 1 measured 0.31 e2e scaling
    efficiency at 8 GPUs with unbound ranks.  The node's CPUs are divided among the ranks that share it.
    Returns a dict describing what was done (for the bench line); never raises."""
    info = {"gpu": gpu_index, "node": None, "cpus": None, "mempolicy": None}
    try:
        node = gpu_numa_node(gpu_index)
        if node is None:
            return info
        info["node"] = node
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED = 1
        nmask = (ctypes.c_ulong * 16)()
        nmask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, MPOL_PREFERRED, nmask, 1024)          # set_mempolicy (x86-64)
        info["mempolicy"] = "preferred" if rc == 0 else "errno %d" % ctypes.get_errno()
    except Exception as e:                                           # placement is an optimisation, never a failure
        info["error"] = str(e)
    return info
