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
