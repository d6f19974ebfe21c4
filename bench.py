"""bench.py - DCT importance-score throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--net resnet_50]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): ResNet-50 ImageNet 224x224 importance generation, batch 256 per GPU
(BASELINE configs[3], the configuration the metric is quoted on), seeded synthetic images and
random-init weights.  A step is one pass of the hot path over one batch: the 49 hook kernels
(one per hooked layer) accumulating per-channel DCT energies; the run ends with the single
all-reduce, the finalise kernel and the segmented top-k, all inside the timed region.

  value     images/s, whole job over all ranks, activations already resident in HBM (the hooked
            feature maps of one batch, 9.2 GB per GPU: larger than L2, nothing to flush)
  e2e       images/s of complete importance generation through the public API: pinned host batch ->
            H2D -> cuDNN fp32 forward with all hooks live -> D2H of the running score sums, every step
  roofline  dominant kernel (largest share of the step; the 56^2 layers' kernel on ResNet-50) timed with CUDA events
            inside the timed region; algorithmic bytes = 4*H*W per scored map
  cpu_baseline  the oracle port of the reference hooks on this box's host cores, bounded sample

`--impl reference` times the reference's CPU implementation (oracle/reference_port.py: the
reference is pure Python and its DCT dependency torch_dct is absent, so the port is the runnable
form) on a bounded sample per step, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'dct_importance_score_images_per_s'
UNIT = 'images/s'

WORKLOADS = {
    'resnet_50': dict(side=224, batch=256, rate='[0.]+[0.1]*3+[0.4]*7+[0.4]*9',
                      name='ResNet-50 ImageNet 224x224 importance generation, batch 256 per GPU, 49 hook sites'),
    'u2netp': dict(side=320, batch=12, rate='[0.40]*40',
                   name='U2-Netp DUTS 320x320 importance generation, batch 12 per GPU, 118 hook sites'),
    'resnet_56': dict(side=32, batch=256, rate='[0.]+[0.18]*29',
                      name='ResNet-56 CIFAR-10 importance generation, batch 256 per GPU, 55 hook sites'),
    'googlenet': dict(side=32, batch=128, rate='[0.4]+[0.85]*2+[0.9]*5+[0.9]*2',
                      name='GoogLeNet CIFAR-10 importance generation, batch 128 per GPU, 10 hook sites'),
    'densenet_40': dict(side=32, batch=256, rate='[0.]+[0.2]*12+[0.]+[0.2]*12+[0.]+[0.2]*12',
                        name='DenseNet-40 CIFAR-10 importance generation, batch 256 per GPU, 39 hook sites'),
    'vgg_16_bn': dict(side=32, batch=128, rate='[0.50]*7+[0.95]*5',
                      name='VGG-16-BN CIFAR-10 importance generation, batch 128 per GPU, 12 hook sites'),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=5)       # = the reference's default --limit
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    p.add_argument('--net', default='resnet_50', choices=tuple(WORKLOADS))
    p.add_argument('--batch', type=int, default=None)
    p.add_argument('--path', default='auto', choices=('auto', 'umma', 'simt'),
                   help="kernel path of the hook launches (the whole net must fit a forced path, so 'tmem' / 'large' are left to 'auto')")
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--graph', action='store_true', help='replay CUDA graphs (hook launches of a step; forward + hooks in the e2e leg): takes the host out of launch-bound nets')
    return p.parse_args()


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0))), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm),
                'window': 'warm-up + kernel-only timed region + end-to-end leg (nvidia-smi -lms 100)'}


def dist_setup(args):
    from dct_pruning_b200 import dist as ddist
    rank, local_rank, world = ddist.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    return rank, local_rank, world


def max_over_ranks(ms, device, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# ====================================================================== our arm
def run_ours(args):
    from dct_pruning_b200 import _lib
    from dct_pruning_b200.compress import get_compress_rate, selection_plan
    from dct_pruning_b200.generate import device_batches
    from dct_pruning_b200.hooks import ScoreSession
    from dct_pruning_b200.sites import VARIANT_INPUT, resolve_module
    from dct_pruning_b200.topk import topk_segmented
    from dct_pruning_b200.zoo import get_network

    rank, local_rank, world = dist_setup(args)
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    wl = WORKLOADS[args.net]
    B = args.batch or wl['batch']
    side = wl['side']
    torch.backends.cudnn.allow_tf32 = False            # activations fp32-exact, like the reference's
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    lib = _lib.load()
    _lib.check(lib.dctp_init())

    torch.manual_seed(0)
    net = get_network(args.net).to(device).eval()
    g = torch.Generator().manual_seed(1000 + rank)
    host_batch = torch.randn(B, 3, side, side, generator=g).pin_memory()

    # ---- capture the hooked feature maps of one batch (resident in HBM for the kernel-only leg)
    session = ScoreSession(net, args.net, path=args.path)
    acts = [None] * len(session.sites)
    handles = []
    for idx, site in enumerate(session.sites):
        def cap(module, inputs, output, idx=idx, take_input=(site.variant == VARIANT_INPUT)):
            acts[idx] = (inputs[0] if take_input else output).detach().clone()
        handles.append(resolve_module(net, site.module).register_forward_hook(cap))
    with torch.no_grad():
        net(host_batch.to(device))
    for h in handles:
        h.remove()
    torch.cuda.synchronize()
    act_bytes = sum(a.numel() * 4 for a in acts)
    # algorithmic bytes: only the scored channel window of each site is read
    site_bytes = []
    for a, site in zip(acts, session.sites):
        c = 12 if site.variant == 'D' else a.shape[1]
        site_bytes.append(4 * a.shape[0] * c * a.shape[2] * a.shape[3])
    alg_bytes_step = sum(site_bytes)

    # top-k plan on the flat score vector (segments = the files the reference's loader reads)
    plan = selection_plan(args.net, get_compress_rate(wl['rate']))

    def finish_run():
        scores = session.finalize_device(check=False)
        segs = {stem: (off, n) for stem, off, n in session.file_segments()}
        offsets, ks, pieces = [0], [], []
        for sel in plan:
            off, n = segs[sel.stem]
            pieces.append(scores[off:off + n])
            offsets.append(offsets[-1] + n)
            ks.append(sel.k)
        kept = topk_segmented(torch.cat(pieces), offsets, ks) if plan else []
        return scores, kept

    # clocks are sampled from here to the end of the end-to-end leg: both timed regions lie inside the window
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- warm-up (bases uploaded, slots allocated, clocks up)
    for _ in range(max(args.warmup, 3)):
        for idx, a in enumerate(acts):
            session.score(idx, a)
    finish_run()
    _lib.check(lib.dctp_check(None))
    session.reset()

    # which kernel a site's launch runs (mirrors the dispatch in csrc/dctp.cu for dense activations)
    def kernel_of(a):
        n = a.shape[2]
        if a.shape[2] == a.shape[3] and 96 <= n <= 320 and n % 16 == 0:
            return 'score_large_kernel (tcgen05, tiled, 16 warps)'
        if a.shape[2] != a.shape[3] or n > 128:
            return 'score_simt_kernel (fp32 CUDA cores)'
        if n > 64:
            return 'score_umma_kernel<128> (tcgen05, smem operands)'
        if n in (52, 56):
            return 'score_t_kernel<64,3> (tcgen05, TMEM-resident operands, 2 producer warpgroups)'
        if n % 2 == 0 and 52 <= n <= 64:
            return 'score_t_kernel<64,3> (tcgen05, TMEM-resident operands, register prefetch)'
        if n % 2 == 0 and 10 <= n <= 32 and a.numel() * 4 >= (32 << 20):
            return 'score_t_kernel<32,6> (tcgen05, TMEM-resident operands, cp.async staging)'
        mode = 0 if n % 4 == 0 else 1 if n % 2 == 0 else 2
        return 'score_umma_kernel<64,%d,1> (tcgen05 bf16x3, smem operands, register prefetch)' % mode
    site_kernel = [kernel_of(a) for a in acts]
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in acts]
          for _ in range(args.steps)]

    # ---- the timed region: K steps of back-to-back hook launches + the end-of-run kernels, nothing else on the stream
    barrier(world)
    launches0 = lib.dctp_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_graph = None
    if args.graph:                                # launch-bound nets: one CUDA graph per step takes the host out of the loop
        warm_stream = torch.cuda.Stream(device=device)
        warm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm_stream):
            for idx, a in enumerate(acts):
                session.score(idx, a)
        torch.cuda.current_stream().wait_stream(warm_stream)
        torch.cuda.synchronize()
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph):
            for idx, a in enumerate(acts):
                session.score(idx, a)
        session.reset()
        launches0 = lib.dctp_launch_count()
    t0.record()
    for step in range(args.steps):
        if step_graph is not None:
            step_graph.replay()
            for idx in range(len(acts)):
                session.images[idx] += B
        else:
            for idx, a in enumerate(acts):
                session.score(idx, a)
    scores, kept = finish_run()
    t1.record()
    barrier(world)
    launches = lib.dctp_launch_count() - launches0 + (args.steps * len(acts) if step_graph is not None else 0)   # replayed launches are not seen by the host counter
    _lib.check(lib.dctp_check(None))
    ms_total = max_over_ranks(t0.elapsed_time(t1), device, world)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- per-kernel durations: the same K steps again with an event pair around every launch (the events keep
    #      consecutive launches from overlapping, so these are isolated launch durations; not part of `value`)
    session.reset()
    for step in range(args.steps):
        for idx, a in enumerate(acts):
            ev[step][idx][0].record()
            session.score(idx, a)
            ev[step][idx][1].record()
    torch.cuda.synchronize()
    session.reset()

    per_site_ms = [statistics.mean(ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(args.steps)) for i in range(len(acts))]
    hbm_peak, _, peak_kind = measured_peaks()
    by_kernel = {}
    for i, name in enumerate(site_kernel):
        d = by_kernel.setdefault(name, {'launches_per_step': 0, 'ms_per_step': 0.0, 'bytes_per_step': 0})
        d['launches_per_step'] += 1
        d['ms_per_step'] += per_site_ms[i]
        d['bytes_per_step'] += site_bytes[i]
    for d in by_kernel.values():
        d['GBps'] = d['bytes_per_step'] / (d['ms_per_step'] / 1e3) / 1e9
        d['frac_hbm'] = d['GBps'] / hbm_peak
    dom_name = max(by_kernel, key=lambda k: by_kernel[k]['ms_per_step'])       # dominant = largest share of the step
    dom = [i for i, name in enumerate(site_kernel) if name == dom_name]
    dom_ms = by_kernel[dom_name]['ms_per_step']
    dom_bytes = by_kernel[dom_name]['bytes_per_step']
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            table = json.load(f)
        prefix = dom_name.split(' ')[0].rstrip('>')           # e.g. score_t_kernel<64,3 matches both load-width variants
        hits = [v for k, v in table.items() if isinstance(v, dict) and k.startswith(prefix)]
        if hits:
            traffic = sum(v['dram_bytes_per_launch'] * v['launches'] for v in hits) / sum(v['launches'] for v in hits)
    by_shape = {}
    for i, a in enumerate(acts):
        key = '%dx%dx%d' % (a.shape[1] if session.sites[i].variant != 'D' else 12, a.shape[2], a.shape[3])
        d = by_shape.setdefault(key, {'launches': 0, 'ms': 0.0, 'bytes': 0})
        d['launches'] += 1
        d['ms'] += per_site_ms[i]
        d['bytes'] += site_bytes[i]
    shape_table = {k: {'launches': v['launches'], 'ms': round(v['ms'], 4), 'GBps': round(v['bytes'] / (v['ms'] / 1e3) / 1e9, 1),
                       'frac_hbm': round(v['bytes'] / (v['ms'] / 1e3) / 1e9 / hbm_peak, 3)} for k, v in by_shape.items()}

    # ---- end to end: pinned host batch -> H2D -> forward with hooks live -> D2H of the running sums
    e2e = None
    if not args.no_e2e:
        session.reset()
        pinned_out = torch.empty(session.used + 1, dtype=torch.float64).pin_memory()
        n_e2e_warm = 2

        replay = None

        def e2e_steps(n):
            # every step: H2D of that step's batch from pinned memory (issued one step ahead on a copy stream, as the
            # package's own `generate.inference` does), forward with all hooks live, D2H of the running sums, sync
            if replay is not None:
                for _ in range(n):
                    replay(host_batch)             # H2D into the graph's input buffer, then one graph launch
                    pinned_out.copy_(session.flat[:session.used + 1], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                return
            for x in device_batches((host_batch for _ in range(n)), device):
                with torch.no_grad():
                    net(x)
                pinned_out.copy_(session.flat[:session.used + 1], non_blocking=True)
                torch.cuda.current_stream().synchronize()

        with session:
            if args.graph:
                replay = session.capture(host_batch.to(device))
            e2e_steps(n_e2e_warm)
            session.reset()
            barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            e2e_steps(args.steps)
            e2e_scores, _ = finish_run()
            host_scores = e2e_scores.cpu()
            e1.record()
            barrier(world)
            # forward alone, same loop without hooks, for the breakdown
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), device, world)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            with torch.no_grad():
                net(host_batch.to(device, non_blocking=True))
        f1.record()
        torch.cuda.synchronize()
        e2e = {'value': world * B * args.steps / (ms_e2e / 1e3), 'unit': UNIT,
               'h2d_bytes_per_step': int(host_batch.numel() * 4), 'd2h_bytes_per_step': int(pinned_out.numel() * 8),
               'ms_per_step': ms_e2e / args.steps, 'forward_only_ms_per_step': f0.elapsed_time(f1) / args.steps,
               'score_checksum': float(host_scores.double().sum()), 'cuda_graph': bool(args.graph)}

    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_hooks_baseline([a[:4].cpu() for a in acts], session.sites, budget_s=12.0)
        on_gpu = gpu_tensor_reference_baseline(acts, session.sites, budget_s=4.0)
        if on_gpu is not None:
            cpu['reference_hook_on_gpu_tensors'] = on_gpu

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16x3 split of fp32 on tcgen05 (fp32 accumulate), fp64 cross-image sums', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'net': args.net, 'batch_per_gpu': B, 'input_side': side, 'limit': args.steps,
                       'hook_sites': len(acts), 'activation_bytes_per_step': act_bytes, 'algorithmic_bytes_per_step': alg_bytes_step,
                       'l2': 'inputs (%.1f GB per step) exceed L2; no flush needed' % (act_bytes / 1e9),
                       'compress_rate': wl['rate'], 'path': args.path, 'cuda_graph': bool(args.graph), 'parallelism': 'batch-sharded x%d, 1 all-reduce per run' % world},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                         'traffic': traffic, 'peak_kind': peak_kind, 'kernel': dom_name,
                         'share_of_step': dom_ms / sum(per_site_ms),
                         'launches_per_step': len(dom), 'bytes_per_step': dom_bytes, 'ms_per_step': dom_ms,
                         'timing': 'CUDA event pair around every launch, over a repeat of the K timed steps on the same stream '
                                   '(the events keep consecutive launches from overlapping: isolated launch durations)'},
            'hook_path_GBps': alg_bytes_step * args.steps / (ms_total / 1e3) / 1e9,
            'by_kernel': {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in by_kernel.items()},
            'by_shape': shape_table,
            'clocks': clocks,
        }
        if e2e is not None:
            line['e2e'] = e2e
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    from dct_pruning_b200 import dist as ddist
    ddist.shutdown()


# ====================================================================== CPU legs
def cpu_hooks_baseline(acts_cpu, sites, budget_s):
    """The reference's hook (oracle port, op for op) over whole images until the budget is spent."""
    from oracle import reference_port as port
    torch.set_num_threads(os.cpu_count() or 1)
    hooks = {'O': port.hook_output, 'D': port.hook_densenet, 'I': port.hook_u2net_input}
    n_img = acts_cpu[0].shape[0]
    done, slices, t0 = 0, 0, time.perf_counter()
    for i in range(n_img):
        for a, site in zip(acts_cpu, sites):
            st = port.ScoreState()
            x = a[i:i + 1]
            hooks[site.variant](st)(None, (x,), x)
            slices += 12 if site.variant == 'D' else x.shape[1]
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {'value': done / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d image(s) x all %d hook sites (%d slices) through oracle.reference_port hooks, hooks only (no CNN forward), %.1f s'
                      % (done, len(sites), slices, dt),
            'us_per_slice': dt / slices * 1e6}


def gpu_tensor_reference_baseline(acts, sites, budget_s):
    """The reference hook in the mode its authors ran it: activations stay on the GPU, one cuFFT-based DCT and one
    `.item()` sync per (image, channel) slice (oracle port of utils/common.py:262-277).  One image, as many sites as fit."""
    from oracle import reference_port as port
    done_sites, slices, t0 = 0, 0, time.perf_counter()
    total_slices = sum(12 if s.variant == 'D' else a.shape[1] for a, s in zip(acts, sites))
    for a, site in zip(acts, sites):
        if site.variant != 'O':
            continue                               # the other two variants copy every slice to the host for cv2
        x = a[0:1]
        port.hook_output(port.ScoreState())(None, (x,), x)
        torch.cuda.synchronize()
        slices += x.shape[1]
        done_sites += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if slices == 0:
        return None
    us = dt / slices * 1e6
    return {'us_per_slice': us, 'images_per_s_extrapolated': 1e6 / (us * total_slices),
            'sample': '1 image x %d output-hook site(s) (%d slices) on GPU tensors, per-slice torch.fft DCT + .item() sync, %.1f s; '
                      'extrapolated linearly to all %d slices of an image' % (done_sites, slices, dt, total_slices)}


def run_reference(args):
    """The reference's own CPU implementation (per-site forward sweep + per-slice Python hook),
    restated in oracle/reference_port.py, on a bounded sample: one image per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from dct_pruning_b200.sites import hook_sites
    from dct_pruning_b200.zoo import get_network
    from oracle import reference_port as port
    wl = WORKLOADS[args.net]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = get_network(args.net).eval()
    sessions = [(s.module, s.variant, [(f.stem, f.lo, f.hi) for f in s.files]) for s in hook_sites(args.net, net)]
    B = 1

    def batches(seed):
        g = torch.Generator().manual_seed(seed)
        return [torch.randn(B, 3, wl['side'], wl['side'], generator=g)]

    def step(i):
        port.imp_score_port(net, sessions, lambda: batches(1000 + i), 1)

    warm = min(args.warmup, 1)
    for i in range(warm):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(warm + i)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    sample = ('%d image per step through the full reference flow (one forward sweep per hook site, %d sites, per-slice Python hook) '
              'on %d host threads' % (B, len(sessions), torch.get_num_threads()))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['name'], 'net': args.net, 'batch_per_step': B, 'input_side': wl['side']},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device; the product path has no CPU fallback')
        run_ours(a)
