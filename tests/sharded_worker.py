"""torchrun worker of test_sharded_trace_two_gpus: sofa_preprocess of ONE logdir by all ranks, then the parsed gpu and
packet shards are gathered on rank 0 (NCCL) where sofa_analyze runs on the resident rows."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    logdir, out = sys.argv[1], sys.argv[2]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    from sofa_b200.engine import CudaEngine
    from sofa_b200.sofa_config import default_report_config
    from sofa_b200 import sharded, hostlogic as hl
    from sofa_b200.sofa_analyze import sofa_analyze
    eng = CudaEngine(local)
    cfg = default_report_config(logdir, cpu_filters="idle:black,tensorflow:orange")
    sharded.sofa_preprocess_sharded(cfg, dist, engine=eng, device=device)
    moved = 0
    for kind in (hl.SRC_PERF, hl.SRC_GPU, hl.SRC_NET):
        moved += sharded.gather_source(eng, kind, dist, device, 0)
    assert moved > 0
    if dist.get_rank() == 0:
        import contextlib
        import io
        cfg._sofa_b200_engine = eng              # the gathered rows are resident: analyze them, do not re-ingest
        with contextlib.redirect_stdout(io.StringIO()):
            feats = sofa_analyze(cfg)
        with open(out, "w") as f:
            json.dump([[n, float(v)] for n, v in feats], f)
    else:
        eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
