"""`SOFA_Config` / `Filter`: the boundary type of the hot path (reference:
bin/sofa_config.py:1-74).  Attribute names and defaults are the reference's so a `cfg`
built by its CLI can be passed to this package unchanged.  Unlike the reference the
filter lists are per-instance (the reference's class-level lists leak between runs,
SURVEY.md A.5)."""

sofa_fieldnames = ["timestamp", "event", "duration", "deviceId", "copyKind", "payload",
                   "bandwidth", "pkt_src", "pkt_dst", "pid", "tid", "name", "category"]


class Filter:
    def __init__(self, keyword, color):
        self.keyword = keyword
        self.color = color


_DEFAULTS = dict(
    cluster_ip=None, perf_events='cycles', verbose=False, num_iterations=20, num_swarms=10,
    cpu_top_k=20, plot_ratio=1, viz_port=8000, cpu_time_offset=0, profile_all_cpus=False,
    enable_aisi=False, aisi_via_strace=False, strace_min_time=1e-6, enable_swarms=False,
    display_swarms=False, hsg_multifeatures=False, use_diff=False, swarm_diff=False,
    logdir='./sofalog/', base_logdir='', match_logdir='', enable_vmstat=False,
    enable_mpstat=True, enable_diskstat=True, enable_pcm=False, cuda_api_tracing=False,
    script_path='', potato_server=None, elapsed_time=0, absolute_timestamp=False, time_base=0,
    pid=-1, timeout=30, sys_mon_rate=10, is_idle_threshold=10, enable_strace=False,
    enable_py_stacks=False, spotlight_gpu=False, blktrace_device=None, nvsmi_data=False,
    roi_begin=0, roi_end=0, netstat_interface=None, nvsmi_time_zone=0, nvprof_inside=False,
    diskstat_filters=None, no_perf_events=False, enable_encode_decode=False, enable_tcpdump=False)


class SOFA_Config:
    columns = sofa_fieldnames

    def __init__(self):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        self.cpu_filters = []
        self.gpu_filters = []
        self.net_filters = []


def default_report_config(logdir, cpu_filters=None, gpu_filters=None, cuda_api_tracing=False):
    """cfg as `sofa report` leaves it (bin/sofa:258-291 default filters)."""
    cfg = SOFA_Config()
    cfg.logdir = logdir if logdir.endswith('/') else logdir + '/'
    cfg.cpu_time_offset = 0.0
    for pair in (cpu_filters or 'idle:black').split(','):
        cfg.cpu_filters.append(Filter(pair.split(':')[0], pair.split(':')[1]))
    cfg.gpu_filters.append(Filter('CUDA_COPY_H2D', 'Red'))
    cfg.gpu_filters.append(Filter('CUDA_COPY_D2H', 'Peru'))
    cfg.gpu_filters.append(Filter('CUDA_COPY_P2P', 'Purple'))
    cfg.cuda_api_tracing = cuda_api_tracing
    if cuda_api_tracing:
        cfg.gpu_filters.append(Filter('CUDA_API', 'DarkSlateGray'))            # bin/sofa:276-277
    for pair in (gpu_filters or '_fw_:royalblue,_bw_:darkviolet,AllReduceKernel:indigo').split(','):
        cfg.gpu_filters.append(Filter(pair.split(':')[0], pair.split(':')[1]))
    return cfg
