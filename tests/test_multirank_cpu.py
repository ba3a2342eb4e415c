"""world_size-2 gloo test of the N>1 host logic (no GPU): every rank preprocesses its own
logdirs (CPU build of the device row codecs as the engine - test infrastructure), results are
checked against the oracle per rank, and the throughput aggregation (sum of events / max of
step times) is checked across ranks."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, world, port, workdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hostsim_engine import HostSimEngine
        from helpers import demangle
        from sofa_b200 import synth, multirank
        from sofa_b200.sofa_config import default_report_config
        from oracle import sofa_oracle as orc
        assert multirank.rank_env() == (rank, rank, world)
        ips = ["10.0.0.%d" % i for i in range(1, 5)]
        base = os.path.join(workdir, "sofalog")
        dirs = multirank.cluster_logdirs(base, ",".join(ips))
        assert dirs[2] == base + "-10.0.0.3/"
        mine = multirank.assign_logdirs(dirs, rank, world)
        assert mine == [dirs[rank], dirs[rank + 2]]
        n_events = 0
        for i, d in enumerate(mine):
            seed = 1000 + 10 * rank + i
            synth.make_logdir(d, seed=seed, n_perf=150 + 10 * rank, n_gpu=120, n_net=40)
            synth.make_logdir(d.rstrip('/') + "_ref/", seed=seed, n_perf=150 + 10 * rank, n_gpu=120, n_net=40)
            n_events += 150 + 10 * rank + 120 + 40
        cfg = default_report_config(base)
        multirank.preprocess_cluster(cfg, mine, engine_factory=HostSimEngine)
        assert cfg.logdir == base + "/"
        for d in mine:
            ref = d.rstrip('/') + "_ref/"
            orc.preprocess(orc.OracleConfig(ref, nprocs=1), demangle=demangle)
            for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
                with open(d + fn, "rb") as f, open(ref + fn, "rb") as g:
                    assert f.read() == g.read(), (rank, d, fn)
        dist.barrier()
        ms_local = 100.0 * (rank + 1)
        ev, ms, thr = multirank.aggregate(dist, n_events, ms_local)
        assert ms == 100.0 * world
        assert ev == sum(2 * (150 + 10 * r + 160) for r in range(world))
        assert abs(thr - ev / (ms * 1e-3)) < 1e-9
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 400)
    with tempfile.TemporaryDirectory() as workdir:
        mp.spawn(_worker, args=(world, port, workdir), nprocs=world, join=True)
