#!/usr/bin/env python
"""bench.py -- images/sec of N-step FCN8 + DAE_h iterative inference (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic CamVid-shaped input:
BASELINE.json configs[1] -- batch 10 of 360x480 images, 11 classes: FCN8 forward (h = pool4, y0),
50 iterations of y <- clip(y + step*(DAE(y,h) - y), 0, 1) with step 0.05, then the metrics.py
confusion-matrix / Jaccard reduction.  Every rank processes its own batch (image sharding, weak
scaling); the only collective is the all-reduce of the int64 confusion matrix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config 2|3|4] [--precision mixed|bf16|fp32x3] [--scaling weak|strong] [--sections a,b,..]

The headline (`value`, `e2e`, `roofline`) is the PARITY-GRADE arithmetic, precision='mixed': fp32-accurate convs
(bf16 hi/lo operand pairs, three tensor-core products) in FCN8 and on the DAE's contracting path, bf16 on the
expanding path -- the variant the -m gpu tests hold to BASELINE.json's fp32 bar (2e-3 max-abs per iteration,
>= 99.9 % argmax agreement, confusion matrix bit-exact given the labels).  The all-bf16 throughput variant, with
its own stated tolerance, is measured in the same run and reported as `bf16_variant`.

`value`  : whole-job images/s with the batch already resident in HBM.
`e2e`    : the same through the public callables with HOST (pinned) buffers: H2D of the images and
           the one-hot targets and D2H of the metrics inside the timed region, every step (the copies
           of step i+1 ride a copy stream under step i's kernels, as with a prefetching iterator).
`roofline`: the dominant kernel, conv_igemm_pair_kernel (tcgen05 cta_group::2), CUDA-event timed.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference path (oracle/, PyTorch
           CPU fp32; Theano is not installable) on the host cores, on a bounded sample.
Other BASELINE.json configs ride the same line as sub-records measured in the same run:
`config3` (FC-DenseNet103 + DAE_h), `config4` (train_dae.py step, data-parallel over the N ranks with the
gradient all-reduce), `steps_sweep` (config 5: iterations 1..100); `--config 3|4` makes one of them the headline.
`--scaling strong`: a fixed set of 80 images (8 batches) is sharded over the ranks (sharding.shard_range).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCLS, H, W, BATCH, N_ITER, STEP = 11, 360, 480, 10, 50, 0.05
TH, TW = 224, 224                    # config 4: train_dae.py crop
LOGIT_GAIN, OUT_GAIN = 10.0, 0.1
STRONG_SET = 80                      # --scaling strong: images in the fixed set
METRIC = 'images/sec, N-step FCN8+DAE iterative inference 360x480'
WORKLOAD = 'FCN8 + DAE_h (n_filters=64, concat_h=pool4, trackind unpool), batch 10 x 360x480, 11 classes, 50 steps, step 0.05, metrics.py Jaccard'
WORKLOAD3 = 'FC-DenseNet103 + DAE_h (concat_h=pool4, padding 0), batch 10 x 360x480, 11 classes, 50 steps, step 0.05, metrics.py Jaccard'
WORKLOAD4 = 'train_dae.py DAE step (rmsprop lr 1e-3, crossentropy + squared_error, noise 0.5: the main pass + one noised mask pass per DePool2D), batch 10 x 224x224 per GPU, data-parallel gradient all-reduce'
WEIGHTS = 'random init (He-uniform FCN8 x logit gain 10, Glorot DAE x out gain 0.1)'
DTYPES = {'bf16': 'bf16',
          'fp32x3': 'fp32 operands as bf16 hi/lo pairs, 3 tensor-core products, fp32 accumulate',
          'mixed': 'fp32-accurate (bf16 hi/lo pairs x 3 tensor-core products, fp32 accumulate) in FCN8 and the DAE contracting path; bf16 operands / fp32 accumulate on the DAE expanding path'}
PARITY = {'mixed': 'tests/test_path_gpu.py: per-iteration y within 2e-3 max-abs, argmax agreement >= 99.9 %, vs the CPU oracle (full size: test_full_size_end_to_end_vs_oracle); confusion matrix bit-exact given the labels',
          'fp32x3': 'tests/test_path_gpu.py: per-iteration y within 2e-3 max-abs, argmax agreement >= 99.9 %, vs the CPU oracle; confusion matrix bit-exact given the labels',
          'bf16': 'own tolerance (tests/test_path_gpu.py::test_full_size_end_to_end_vs_oracle): bf16 rounding flips pool-mask ties, so per-iteration y is only within the stated bf16 bound of the oracle; confusion matrix bit-exact given the labels'}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# --------------------------------------------------------------------------- CPU reference arm
CPU_SAMPLE_ITERS = 50          # the whole loop: no extrapolation over iterations (one image of the batch is the sample)
CPU_ARM_BUDGET_S = 240.0       # --impl reference: the whole --steps K --warmup W run stays within a few minutes


def cpu_reference_sample(n_dae_iters=CPU_SAMPLE_ITERS):
    """One bounded sample of the workload on the host cores with the oracle: 1 image at 360x480,
    FCN8 forward + `n_dae_iters` of the 50 loop iterations + metrics; returns the per-image time
    extrapolated linearly to 50 iterations."""
    import torch
    from oracle import nets, weights, metrics as M
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    st = cpu_reference_sample.state
    if st is None:
        X, L, _ = weights.synthetic_batch(1, H, W, NCLS, seed=0)
        pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN)
        pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=OUT_GAIN)
        st = cpu_reference_sample.state = (X, L, pf, pd)
    X, L, pf, pd = st
    with torch.no_grad():
        t0 = time.perf_counter()
        h, y = nets.fcn8_forward(pf, X, NCLS)
        t1 = time.perf_counter()
        for _ in range(n_dae_iters):
            p = nets.dae_forward(pd, y, h, 100)
            g = y - p
            y = torch.clamp(y - STEP * g, 0.0, 1.0)
            float(torch.linalg.vector_norm(g, dim=1).mean())
        t2 = time.perf_counter()
        M.val_fn(y.numpy(), L.numpy(), NCLS, [NCLS])
        t3 = time.perf_counter()
    t_img = (t1 - t0) + (t2 - t1) / n_dae_iters * N_ITER + (t3 - t2)
    return t_img, cores, {'fcn8_s': t1 - t0, 'dae_iter_s': (t2 - t1) / n_dae_iters, 'metrics_s': t3 - t2}


cpu_reference_sample.state = None


def cpu_sample_text(n):
    if n >= N_ITER:
        return ('1 image 360x480 of the batch of 10: FCN8 forward + all %d DAE iterations + metrics timed with the PyTorch-CPU '
                'oracle (port of the Theano path); the reference iterates image by image, so images/s = 1 / that time' % N_ITER)
    return ('1 image 360x480: FCN8 forward + %d of the 50 DAE iterations + metrics timed with the PyTorch-CPU '
            'oracle (port of the Theano path), per-image time extrapolated linearly to 50 iterations' % n)


def cpu_reference_sample_config3(n_dae_iters=5):
    """config 3 on the host cores: FC-DenseNet103 forward on one batch-2 sample is avoided (batch-stat BN couples the
    batch): 1 image 360x480 through the oracle DenseNet + n of the 50 DAE iterations (padding 0)."""
    import torch
    from oracle import nets, weights, densenet, metrics as M
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    X, L, _ = weights.synthetic_batch(1, H, W, NCLS, seed=0)
    pn = densenet.synthetic_densenet_params(3, NCLS, seed=2, logit_gain=4.0)
    pd = weights.synthetic_dae_params(NCLS, 464, seed=1, out_gain=OUT_GAIN)
    with torch.no_grad():
        t0 = time.perf_counter()
        h, y = densenet.densenet_forward(pn, X, NCLS, layer=('pool4',))
        t1 = time.perf_counter()
        for _ in range(n_dae_iters):
            p = nets.dae_forward(pd, y, h, 0)
            y = torch.clamp(y - STEP * (y - p), 0.0, 1.0)
        t2 = time.perf_counter()
        M.val_fn(y.numpy(), L.numpy(), NCLS, [NCLS])
        t3 = time.perf_counter()
    return (t1 - t0) + (t2 - t1) / n_dae_iters * N_ITER + (t3 - t2), cores


def cpu_reference_sample_config4(n_images=2):
    """config 4 on the host cores: one oracle train step (autograd, rmsprop) on `n_images` 224x224 crops; returns s / image."""
    import torch
    from oracle import weights, train as OT
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    X, L, _ = weights.synthetic_batch(n_images, TH, TW, NCLS, seed=5)
    pd = [p.clone() for p in weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=OUT_GAIN)]
    accus = [torch.zeros_like(p) for p in pd]
    y = L[:, :NCLS].contiguous()
    hs = (((TH + 198) // 2 // 2 // 2) // 2, ((TW + 198) // 2 // 2 // 2) // 2)
    gen = torch.Generator().manual_seed(3)
    h = torch.relu(torch.randn((n_images, 512) + hs, generator=gen))
    nm, nk = torch.randn(y.shape, generator=gen) * 0.5, torch.randn((6,) + tuple(y.shape), generator=gen) * 0.5      # one draw per DePool2D
    t0 = time.perf_counter()
    OT.train_step(pd, accus, y, h, L, NCLS, 100, 1e-3, noise_main=nm, noise_mask=nk)
    return (time.perf_counter() - t0) / n_images, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.config == 3:
        sample = lambda: cpu_reference_sample_config3()[0]           # noqa: E731
        text, workload, batch = ('1 image 360x480: oracle FC-DenseNet103 forward + 5 of the 50 DAE iterations + metrics, '
                                 'extrapolated linearly to 50 iterations'), WORKLOAD3, BATCH
    elif args.config == 4:
        sample = lambda: cpu_reference_sample_config4()[0]           # noqa: E731
        text, workload, batch = '2 images 224x224: one oracle train step (autograd + rmsprop), time per image', WORKLOAD4, BATCH
    else:
        # one image with ALL 50 iterations per step when K + W of those fit the budget, else as many iterations as do
        _, _, parts = cpu_reference_sample(2)
        per_step = CPU_ARM_BUDGET_S / max(1, args.steps + args.warmup)
        n_it = int(max(5, min(N_ITER, (per_step - parts['fcn8_s'] - parts['metrics_s']) / parts['dae_iter_s'])))
        sample = lambda: cpu_reference_sample(n_it)[0]               # noqa: E731
        text, workload, batch = cpu_sample_text(n_it), WORKLOAD, BATCH
    for _ in range(args.warmup):
        sample()
    times = [sample() for _ in range(args.steps)]
    cores = len(os.sched_getaffinity(0))
    t = sum(times) / len(times)
    val = 1.0 / t
    line = {
        'impl': 'reference', 'metric': METRIC if args.config == 2 else metric_name(args.config), 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus,
        # a step of this arm is its bounded sample (seconds per image, or per sampled image), not the batch of 10
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'weights': WEIGHTS, 'images_per_step': 1, 'batch_of_the_workload': batch},
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': text},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(config):
    return {2: METRIC, 3: 'images/sec, N-step FC-DenseNet103+DAE iterative inference 360x480',
            4: 'images/sec, train_dae.py DAE train step 224x224 (data-parallel)'}[config]


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw.instant,enforced.power.limit')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = threading.Thread(target=self._run, daemon=True)

    def _query(self, fields):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + fields,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        return [p.strip() for p in out.strip().split(',')]

    def _run_nvml(self):
        """NVML in-process (about a millisecond per sample, so that even a sub-second timed region is sampled many times);
        same fields as the nvidia-smi query."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [getattr(nv, 'nvmlClocksEventReason' + n, None) or getattr(nv, 'nvmlClocksThrottleReason' + n)
                for n in ('HwSlowdown', 'HwThermalSlowdown', 'SwThermalSlowdown', 'SwPowerCap')]
        limit = nv.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
        while not self._stop.is_set():
            r = get_reasons(h)
            self.samples.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                                ['Active' if (r & b) else 'Not Active' for b in bits] +
                                [str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0), str(limit)])
            self._stop.wait(0.02)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass                    # no NVML binding / call refused: fall back to polling nvidia-smi
        fields = self.QUERY
        while not self._stop.is_set():
            try:
                parts = self._query(fields)
                if len(parts) < 8 and fields is self.QUERY:      # a driver without the power fields: clocks and reasons only
                    fields = ','.join(self.QUERY.split(',')[:6])
                    parts = self._query(fields)
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        out = {'sm_mhz': mhz[len(mhz) // 2] if mhz else None, 'sm_max_mhz': int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
               'reasons': reasons, 'samples': len(self.samples)}
        try:        # instantaneous board power under load against its enforced limit (explains sw_power_cap)
            watts = sorted(float(s[6]) for s in self.samples if len(s) >= 8)
            if watts:
                out['power_w'] = watts[len(watts) // 2]
                out['power_max_w'] = watts[-1]
                out['power_limit_w'] = float(self.samples[0][7])
        except ValueError:
            pass
        return out


# --------------------------------------------------------------------------- B200 arm
class Ctx(object):
    """Process-wide state of the B200 arm: rank / world, device, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.peaks = load_peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, n):
        """Barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks -> (ms, wall ms)."""
        torch = self.torch
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(n):
            fn()
        e.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([s.elapsed_time(e), wall * 1e3], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])


def build_inference(ctx, precision, segm='fcn8', kind='standard'):
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.FCDenseNet import build_fcdensenet
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import IterativeInference
    from iterative_inference_segm_b200 import synthetic as weights
    if kind == 'contextmod':       # the reference CLI's default DAE: context module conditioned on the image
        from iterative_inference_segm_b200.models.contextmod_dae import buildDAE_contextmod
        fcn = buildFCN8(3, None, n_classes=NCLS, layer=['input', 'probs_dimshuffle'],
                        params=weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN), precision=precision)
        dae = buildDAE_contextmod([None], None, NCLS, concat_h=['input'], noise=0.0,
                                  params=weights.synthetic_contextmod_params(NCLS, 3, seed=3), nb_features_to_concat=3)
        return fcn, dae, IterativeInference(dae, NCLS, [NCLS])
    if kind == 'fcn8':             # inference()'s own default kind (iterative_inference.py:64): the FCN8-shaped DAE, h = pool4
        from iterative_inference_segm_b200.models.fcn8_dae import buildFCN8_DAE
        fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'],
                        params=weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN), precision=precision)
        dae = buildFCN8_DAE([None], None, NCLS, nb_in_channels=NCLS, concat_h=['pool4'], noise=0.0, precision=precision,
                            params=weights.synthetic_fcn8_params(NCLS, NCLS, seed=6, logit_gain=LOGIT_GAIN, concat=('pool4', 512)),
                            nb_features_to_concat=512)
        return fcn, dae, IterativeInference(dae, NCLS, [NCLS])
    if segm == 'fcn8':
        fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'],
                        params=weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN), precision=precision)
        nb_h, padding = fcn[0].output_shape[1], 100
    else:
        fcn = build_fcdensenet(None, ['pool4'], 3, NCLS, params=weights.synthetic_densenet_params(3, NCLS, seed=2, logit_gain=4.0),
                               precision=precision)
        nb_h, padding = 464, 0
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=nb_h, padding=padding, concat_h=['pool4'],
                   noise=0.0, n_filters=64, conv_before_pool=1, additional_pool=2, skip=True, unpool_type='trackind',
                   params=weights.synthetic_dae_params(NCLS, nb_h, seed=1, out_gain=OUT_GAIN), precision=precision)
    return fcn, dae, IterativeInference(dae, NCLS, [NCLS])


def measure_inference(ctx, precision, segm='fcn8', n_iter=N_ITER, with_e2e=True, clocks=False, strong=False, kind='standard'):
    """Times the step (FCN forward + n_iter loop iterations + metrics + the confusion-matrix all-reduce) on this rank's
    batches.  Returns (record, objects for the roofline leg)."""
    torch, dist, args = ctx.torch, ctx.dist, ctx.args
    from iterative_inference_segm_b200 import _lib
    from iterative_inference_segm_b200 import synthetic as weights
    from iterative_inference_segm_b200.sharding import shard_range, allreduce_metrics
    fcn, dae, ii = build_inference(ctx, precision, segm, kind)
    fnet = fcn[0].net
    hkey = 'input' if kind == 'contextmod' else 'pool4' if segm == 'fcn8' else 'pool4_bf16'
    want = ('input', 'probs_dimshuffle') if kind == 'contextmod' else ('pool4', 'probs_dimshuffle')
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    if strong:       # a fixed set of STRONG_SET images = batches of 10; rank r owns batches [lo, hi)
        lo, hi = shard_range(STRONG_SET // BATCH, rank, world)
        my_batches = list(range(lo, hi))
    else:            # rank r holds images [r*BATCH, (r+1)*BATCH) of the synthetic set
        my_batches = [rank]
    hosts = []
    for b in my_batches:
        X, L, _ = weights.synthetic_batch(BATCH, H, W, NCLS, seed=100 + b)
        hosts.append((X.pin_memory(), L.pin_memory()))
    devs = [(X.to(dev), L.to(dev)) for X, L in hosts]
    cm_total = torch.zeros(NCLS * NCLS, dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)

    def batch_device(Xd, Ld):
        out = fnet.forward(Xd, want=want)
        res = ii.run(out[hkey], out['probs_dimshuffle'], STEP, n_iter, onehot=Ld)
        cm_total.add_(res['cm'].sum(0))
        counts.add_(res['counts'].sum(0))
        return res

    def step_device():
        cm_total.zero_(); counts.zero_()
        res = None
        for Xd, Ld in devs:
            res = batch_device(Xd, Ld)
        allreduce_metrics(cm_total, counts)          # the one collective of the inference path (sharding.py)
        return res

    # End to end: every step copies its own inputs from pinned host memory and reads its result back.  The copies
    # of batch i+1 are issued on a copy stream before batch i's kernels (two device buffers, like a prefetching data
    # iterator), so only the first batch's H2D is exposed; the D2H read of the result blocks the host every step.
    copy_stream = torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(devs[0][0]), torch.empty_like(devs[0][1])) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(i):
        b = i % 2
        Xh, Lh = hosts[i % len(hosts)]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])
            in_bufs[b][0].copy_(Xh, non_blocking=True)
            in_bufs[b][1].copy_(Lh, non_blocking=True)
            ev_ready[b].record(copy_stream)

    def run_e2e(n_steps):
        cur = torch.cuda.current_stream()
        for b in range(2):
            ev_free[b].record(cur)
        n = n_steps * len(hosts)
        issue_h2d(0)
        out = None
        for i in range(n):
            if i + 1 < n:
                issue_h2d(i + 1)
            b = i % 2
            if i % len(hosts) == 0:
                cm_total.zero_(); counts.zero_()
            cur.wait_event(ev_ready[b])
            res = batch_device(*in_bufs[b])
            ev_free[b].record(cur)
            if (i + 1) % len(hosts) == 0:
                allreduce_metrics(cm_total, counts)
                out = (cm_total.cpu(), counts.cpu(), res['n_exec'].cpu())      # the step's result leaves the device
        return out

    # ---- warm-up (also captures the CUDA graph) + launch census
    l0 = _lib.launch_count()
    res = step_device()
    torch.cuda.synchronize()
    census_first = _lib.launch_count() - l0
    for _ in range(max(args.warmup - 1, 0)):
        res = step_device()
    torch.cuda.synchronize()
    n_exec = res['n_exec'].cpu().tolist()
    assert all(n == n_iter for n in n_exec), 'early exit in the benchmark: %s' % n_exec
    l1 = _lib.launch_count()
    step_device()
    torch.cuda.synchronize()
    eager_per_step = _lib.launch_count() - l1
    # eager FCN / boundary launches are counted live; the loop is a graph replay of graph_kernel_nodes library kernels
    launches_per_step = eager_per_step + ii.graph_kernel_nodes * len(devs)

    clk = None
    if clocks:
        with ClockSampler(ctx.local) as cs:
            ms_dev, _ = ctx.timed(step_device, args.steps)
        clk = cs.summary()
    else:
        ms_dev, _ = ctx.timed(step_device, args.steps)
    imgs = BATCH * (STRONG_SET // BATCH if strong else world) * args.steps
    rec = {'precision': precision, 'value': imgs / (ms_dev * 1e-3), 'ms_per_step': ms_dev / args.steps,
           'executed_iterations': n_exec[0], 'launches_per_step': int(launches_per_step), 'census_first': int(census_first),
           'clocks': clk, 'batches_per_rank_step': len(devs)}
    if with_e2e:
        run_e2e(1)
        ms_e2e, wall_e2e = ctx.timed(lambda: run_e2e(args.steps), 1)
        ms_e2e = max(ms_e2e, wall_e2e)     # D2H reads block the host: wall clock covers them
        Xh, Lh = hosts[0]
        rec['e2e'] = {'value': imgs / (ms_e2e * 1e-3), 'unit': 'images/s',
                      'h2d_bytes_per_step': int((Xh.numel() + Lh.numel()) * 4 * len(hosts)),
                      'd2h_bytes_per_step': int(cm_total.numel() * 8 + counts.numel() * 8 + BATCH * 4), 'ms_per_step': ms_e2e / args.steps}
    rec['cm'] = cm_total.cpu().numpy()
    return rec, (fcn, dae, ii)


def roofline_leg(ctx, dae, ii, precision):
    """Every conv launch of one steady-state DAE application (the 49 of 50), event-timed eagerly on the launching
    stream; the dominant kernel is the CTA-pair implicit GEMM."""
    torch = ctx.torch
    from iterative_inference_segm_b200.profiling import KernelTimer
    peaks = ctx.peaks
    st = ii._buffers(BATCH, H, W, N_ITER, False)
    timer = KernelTimer()
    reps = 3
    upd = None if dae.net.split_up else dict(y=st['y'], active=st['active'], norm_acc=st['norm_acc'], step=STEP)   # as in the captured loop
    st['active'].fill_(1)
    dae.net.logits(st['h'], st['y_bf16'], full_down=True, update=upd)      # fills the iteration-invariant borders
    with timer.recording():
        for _ in range(reps + 1):
            dae.net.logits(st['h'], st['y_bf16'], full_down=False, update=upd)  # the steady-state application
    summ = timer.summary()
    fl = dae.net.executed_conv_flops(H, W, steady_state=True)               # fp32-conv (algorithmic) FLOPs, per image
    flt = dae.net.executed_conv_flops(H, W, steady_state=True, tensor=True)  # FLOPs the tensor cores execute
    convs = [(tag, sum(v[1:]) / len(v[1:])) for (name, tag), v in summ.items() if name == 'conv2d']   # launch order; drop the cold rep
    assert len(convs) == len(fl)
    conv_total_ms = sum(ms for _, ms in convs)
    other = {}
    for (name, tag), v in summ.items():
        other[name] = other.get(name, 0.0) + sum(v[1:]) / len(v[1:])
    # dominant kernel: the launches the library reports as kernel 1 (CTA pair) (tag[-1] = iiseg_last_conv_plan)
    dom = [(f, ft, ms) for f, ft, (tag, ms) in zip(fl, flt, convs) if tag[-1][0] == 1]
    dom_alg, dom_tensor, dom_ms = sum(d[0] for d in dom) * BATCH, sum(d[1] for d in dom) * BATCH, sum(d[2] for d in dom)
    achieved = dom_tensor / (dom_ms * 1e-3) / 1e12
    peak = peaks['bf16_tflops_sustained']
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')      # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic = tj.get('conv_igemm_pair_kernel/' + precision, tj.get('conv_igemm_pair_kernel<256>', {})).get('dram_bytes_per_launch')
    roof = {'bound': 'tensor',
            'kernel': 'conv_igemm_pair_kernel<256|128> (tcgen05 cta_group::2 implicit GEMM over CTA pairs): %d of the 12 conv launches of one '
                      'steady-state DAE application, batch 10 (conv2_1..conv6_1, up_conv6..up_conv3 as the plan selects); executed FLOPs on the '
                      'y-dependent / crop-dependent windows; precision %s' % (len(dom), precision),
            'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
            'peak_source': peaks['source'] + ' bf16_tflops_sustained', 'launch_ms': dom_ms / len(dom),
            'flops_per_launch': dom_tensor / len(dom),
            'flops_note': 'tensor-core FLOPs executed: an fp32-accurate MAC is three bf16 products (hi*hi + lo*hi + hi*lo); '
                          'fp32-conv (algorithmic) FLOPs of the same launches: %.4g per launch = %.1f TFLOP/s' % (
                              dom_alg / len(dom), dom_alg / (dom_ms * 1e-3) / 1e12),
            'all_12_conv_launches': {'achieved': sum(flt) * BATCH / (conv_total_ms * 1e-3) / 1e12, 'ms': conv_total_ms,
                                     'tensor_flops_per_application': sum(flt) * BATCH, 'fp32_conv_flops_per_application': sum(fl) * BATCH,
                                     'frac': sum(flt) * BATCH / (conv_total_ms * 1e-3) / 1e12 / peak,
                                     'note': 'includes the 16-channel first layer, the fused 2x2 pool + tie mask epilogues and the fused softmax/update epilogue of up_conv1'}}
    unpool_b = 0.0
    for (name, tag), v in summ.items():
        if name == 'unpool2':       # algorithmic bytes: out written once, the touched u / mask windows read once
            (n, uh, uw, c), (_, oh, ow, co) = tag
            unpool_b += n * oh * ow * co * 2 + n * ((oh + 1) // 2 + 1) * ((ow + 1) // 2 + 1) * co * 2.5
    breakdown = {'ms_per_dae_application': {k: round(v, 4) for k, v in other.items()},
                 'conv_ms_in_launch_order': [round(ms, 4) for _, ms in convs],
                 'conv_tensor_tflops_in_launch_order': [round(f * BATCH / (ms * 1e-3) / 1e12, 1) for f, (_, ms) in zip(flt, convs)],
                 'conv_kernel_in_launch_order': ['%s<%d>' % (('per_tap', 'pair', 'halo', 'npack')[tag[-1][0]], tag[-1][1]) for tag, _ in convs],
                 'unpool_gbs': unpool_b / (other.get('unpool2', 1e9) * 1e-3) / 1e9,
                 'unpool_frac_of_hbm': unpool_b / (other.get('unpool2', 1e9) * 1e-3) / 1e9 / peaks['hbm_gbs'],
                 'hbm_peak_gbs': peaks['hbm_gbs'],
                 'note': 'max-pool + tie mask are fused into the contracting-path conv epilogues; softmax + y update + norm into up_conv1'}
    return roof, breakdown


def contextmod_leg(ctx, strong=False):
    """kind='contextmod' (models/contextmod_dae.py, the reference CLI's default DAE; concat_h=['input']): FCN8 + 50 iterations
    of the context module on the same batch, plus the application's time against the fp32 FMA rate of the CUDA cores (the
    module is 11-channel fp32 stencil work: csrc/contextmod.cu)."""
    torch = ctx.torch
    rec, (fcn, dae, ii) = measure_inference(ctx, 'mixed', kind='contextmod', strong=strong)
    out = {'workload': 'FCN8 (fp32-accurate) + context-module DAE (concat_h=input, fp32), batch 10 x 360x480, 11 classes, 50 steps, step 0.05, metrics.py Jaccard',
           'value': rec['value'], 'unit': 'images/s', 'ms_per_step': rec['ms_per_step'], 'e2e': rec['e2e'],
           'dtype': 'f32 (CUDA-core FMA) in the context module', 'executed_iterations': rec['executed_iterations'],
           'parity': 'tests/test_path_gpu.py::test_contextmod_dae_vs_oracle: probabilities within 2e-5 of the fp32 oracle'}
    if ctx.rank == 0:
        net = dae.net
        X = torch.rand((BATCH, 3, H, W), device=ctx.dev)
        y = torch.softmax(torch.randn((BATCH, NCLS, H, W), device=ctx.dev), 1)
        net.logits(X, None, full_down=True, y_f32=y)
        n = 20
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(n):
            net.logits(X, None, full_down=False, y_f32=y)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        fl = sum(net.executed_conv_flops(H, W)) * BATCH
        props = torch.cuda.get_device_properties(ctx.dev)
        peak = props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12          # 128 FMA lanes per SM at the 1965 MHz boost clock
        out['application'] = {'ms': ms, 'launches': 7, 'fp32_tflops': fl / (ms * 1e-3) / 1e12, 'fp32_peak_nominal_tflops': peak,
                              'frac': fl / (ms * 1e-3) / 1e12 / peak,
                              'note': 'executed 2*MAC of the 8 conv layers (real 11 channels) over the event-timed application; peak = SMs x 128 lanes x 2 x 1.965 GHz (nominal, the GPU runs ~1.65-1.9 GHz under load)'}
    del fcn, dae, ii
    torch.cuda.empty_cache()
    return out


def config4_leg(ctx):
    """train_dae.py step (config 4): rank r trains on its own batch of 10 crops; global loss denominators and the 24
    gradient matrices are all-reduced over NCCL (sharding.World)."""
    torch, args = ctx.torch, ctx.args
    from iterative_inference_segm_b200 import synthetic as S, _kernels as K, _lib
    from iterative_inference_segm_b200.sharding import World
    from iterative_inference_segm_b200.train_dae import DAETrainer
    tr = DAETrainer(NCLS, 512, 100, S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=OUT_GAIN), learning_rate=1e-3, noise=0.5)
    X, L, _ = S.synthetic_batch(BATCH, TH, TW, NCLS, seed=5 + ctx.rank)
    L = L.to(ctx.dev)
    y = L[:, :NCLS].contiguous()
    gen = torch.Generator(device=ctx.dev).manual_seed(3 + ctx.rank)
    hs = (((TH + 198) // 2 // 2 // 2) // 2, ((TW + 198) // 2 // 2 // 2) // 2)
    h = K.pack_nchw(torch.relu(torch.randn((BATCH, 512) + hs, device=ctx.dev, generator=gen)), 512)
    nm = torch.randn(y.shape, device=ctx.dev, generator=gen)
    # one mask-noise draw PER DePool2D: the reference's training graph re-evaluates the contracting path up to each pool with
    # an independent GaussianNoiseLayer draw (layers/mylayers.py:91-93, tests/golden/ref_noise.npz) -- 1 + 6 noised passes
    nk = torch.randn((tr.geo.total,) + tuple(y.shape), device=ctx.dev, generator=gen)
    world = World() if ctx.world > 1 else None

    def step():
        if world is None:
            tr.step_graphed(h, y, L, nm, nk)
        else:
            tr.step_dp(h, y, L, nm, nk, world)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    tr.step(h, y, L, nm, nk, world=world)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    loss0 = tr.loss_value()
    n = max(args.steps, 5)
    ms, _ = ctx.timed(step, n)
    ar = None
    if world is not None:        # what the gradient exchange costs: the same eager step without it, and with one blocking all-reduce
        ms_local, _ = ctx.timed(lambda: tr.step(h, y, L, nm, nk), n)
        ms_block, _ = ctx.timed(lambda: tr.step(h, y, L, nm, nk, world=world), n)
        ar = dict(tr.dp_info(), ms_per_step_without_exchange=ms_local / n, ms_per_step_blocking_allreduce=ms_block / n,
                  exposed_ms_per_step=(ms - ms_local) / n,
                  note='the data-parallel step runs eagerly (NCCL work is not captured); the single-GPU number is a CUDA-graph replay')
    shared = None
    if world is None:            # for comparison: ONE shared mask pass for all levels (round 2's first build; less work than the reference does)
        for _ in range(3):
            tr.step_graphed(h, y, L, nm, nk[0])
        ms_sh, _ = ctx.timed(lambda: tr.step_graphed(h, y, L, nm, nk[0]), n)
        shared = {'ms_per_step': ms_sh / n, 'value': BATCH * n / (ms_sh * 1e-3)}
    rec = {'workload': WORKLOAD4, 'value': BATCH * ctx.world * n / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms / n,
           'steps': n, 'dtype': 'bf16 operands, fp32 accumulate, fp32 master weights', 'launches_per_step': int(launches),
           'loss': loss0, 'allreduce': ar,
           'mask_noise': 'one independent draw and one contracting-path pass (levels 1..p) per DePool2D, as in the reference graph',
           'one_shared_mask_pass': shared}
    return rec


def run_b200(args):
    import numpy as np
    from iterative_inference_segm_b200.csrc.build import build
    build()
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    from iterative_inference_segm_b200.functions import jaccard_from_cm
    rank, world = ctx.rank, ctx.world
    sections = set(args.sections.split(',')) if args.sections else {'headline', 'bf16', 'roofline', 'cpu', 'config3', 'config4', 'sweep', 'contextmod', 'fcn8dae'}
    strong = args.scaling == 'strong'
    line = None

    if args.config == 2:
        rec, (fcn, dae, ii) = measure_inference(ctx, args.precision, clocks=True, strong=strong)
        jac = jaccard_from_cm(rec['cm'])
        roof = breakdown = cpu_base = None
        if rank == 0 and 'roofline' in sections:
            roof, breakdown = roofline_leg(ctx, dae, ii, args.precision)
        ctx.barrier()
        line = {
            'metric': METRIC, 'value': rec['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': rec['ms_per_step'], 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': DTYPES[args.precision], 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'per_gpu_batch': BATCH, 'global_batch': STRONG_SET if strong else BATCH * world,
                       'parallelism': ('a fixed set of %d images sharded over %d rank(s) in whole batches of 10' % (STRONG_SET, world)) if strong
                       else 'image shards, dp%d' % world,
                       'precision': args.precision, 'parity': PARITY[args.precision], 'weights': WEIGHTS,
                       'l2': 'working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush',
                       'executed_iterations': rec['executed_iterations']},
            'e2e': rec['e2e'], 'gpu_launches': int(rec['launches_per_step'] * args.steps),
            'clocks': rec['clocks'], 'roofline': roof, 'cpu_baseline': None, 'breakdown': breakdown,
            'result': {'mean_jaccard': float(np.nanmean(jac[0] / jac[1])), 'first_step_launch_census': rec['census_first']},
        }
        del fcn, dae, ii
        torch.cuda.empty_cache()
        if 'bf16' in sections and args.precision != 'bf16':
            rb, (fcn, dae, ii) = measure_inference(ctx, 'bf16', strong=strong)
            rfb = None
            if rank == 0 and 'roofline' in sections:
                rfb, bdb = roofline_leg(ctx, dae, ii, 'bf16')
                rfb = {'achieved': rfb['achieved'], 'frac': rfb['frac'], 'launch_ms': rfb['launch_ms'], 'all_12_conv_launches': rfb['all_12_conv_launches'],
                       'conv_ms_in_launch_order': bdb['conv_ms_in_launch_order'], 'conv_tensor_tflops_in_launch_order': bdb['conv_tensor_tflops_in_launch_order'],
                       'unpool_gbs': bdb['unpool_gbs']}
            ctx.barrier()
            line['bf16_variant'] = {'value': rb['value'], 'unit': 'images/s', 'ms_per_step': rb['ms_per_step'], 'e2e': rb['e2e'],
                                    'dtype': DTYPES['bf16'], 'parity': PARITY['bf16'], 'roofline': rfb,
                                    'note': 'throughput variant: same kernels, all-bf16 operands'}
            del fcn, dae, ii
            torch.cuda.empty_cache()
        if 'sweep' in sections:          # config 5: iterations 1..100 (each its own captured graph), same batch
            sweep = {}
            for n_it in (1, 10, 50, 100):
                rs, objs = measure_inference(ctx, args.precision, n_iter=n_it, with_e2e=False, strong=strong)
                sweep[str(n_it)] = {'value': rs['value'], 'ms_per_step': rs['ms_per_step']}
                del objs
                torch.cuda.empty_cache()
            line['steps_sweep'] = {'unit': 'images/s', 'precision': args.precision, 'by_iterations': sweep}
        if 'config3' in sections:
            r3, objs = measure_inference(ctx, 'bf16', segm='densenet', strong=strong)
            line['config3'] = {'workload': WORKLOAD3, 'value': r3['value'], 'unit': 'images/s', 'ms_per_step': r3['ms_per_step'], 'e2e': r3['e2e'],
                               'dtype': 'bf16 operands, fp32 stacks and accumulation', 'executed_iterations': r3['executed_iterations']}
            del objs
            torch.cuda.empty_cache()
            r3p, objs = measure_inference(ctx, 'mixed', segm='densenet', with_e2e=False, strong=strong)
            line['config3']['parity_grade'] = {'value': r3p['value'], 'unit': 'images/s', 'ms_per_step': r3p['ms_per_step'],
                                               'dtype': 'fp32-accurate FC-DenseNet103 (bf16 hi/lo pairs x 3 products) + ' + DTYPES['mixed'],
                                               'parity': 'tests/test_densenet_gpu.py::test_densenet_fp32_accurate_variant_vs_oracle: probabilities within 2e-3, argmax >= 99.9 %'}
            del objs
            torch.cuda.empty_cache()
        if 'contextmod' in sections:
            line['contextmod'] = contextmod_leg(ctx, strong)
        if 'fcn8dae' in sections:
            rk, objs = measure_inference(ctx, 'mixed', kind='fcn8', strong=strong)
            line['fcn8_dae'] = {'workload': 'FCN8 + FCN8-shaped DAE (kind=fcn8, concat_h=pool4; a full VGG16/fc6/fc7 forward per iteration), batch 10 x 360x480, 11 classes, 50 steps, step 0.05, metrics.py Jaccard',
                                'value': rk['value'], 'unit': 'images/s', 'ms_per_step': rk['ms_per_step'], 'e2e': rk['e2e'],
                                'dtype': DTYPES['fp32x3'], 'executed_iterations': rk['executed_iterations'],
                                'parity': 'tests/test_path_gpu.py::test_fcn8_shaped_dae_vs_oracle: within 2e-3 / 99.9 % of the fp32 oracle'}
            del objs
            torch.cuda.empty_cache()
        if 'config4' in sections:
            line['config4'] = config4_leg(ctx)
        if rank == 0 and world == 1 and 'cpu' in sections and not args.no_cpu_baseline:
            cpu_reference_sample(2)                                   # warm-up: thread pool, oneDNN primitive caches
            t_img, cores, parts = cpu_reference_sample(CPU_SAMPLE_ITERS)            # ~15 s of host work: one whole image
            line['cpu_baseline'] = {'value': 1.0 / t_img, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                    'sample': cpu_sample_text(CPU_SAMPLE_ITERS), 'parts_s': {k: round(v, 3) for k, v in parts.items()}}
    elif args.config == 3:
        with ClockSampler(ctx.local) as cs:
            r3, objs = measure_inference(ctx, 'bf16', segm='densenet', strong=strong)
        jac = jaccard_from_cm(r3['cm'])
        line = {'metric': metric_name(3), 'value': r3['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': r3['ms_per_step'], 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
                'dtype': 'bf16 operands, fp32 stacks and accumulation', 'data': 'synthetic',
                'config': {'workload': WORKLOAD3, 'per_gpu_batch': BATCH, 'parallelism': 'image shards in whole batches (batch-stat BN), dp%d' % world,
                           'executed_iterations': r3['executed_iterations']},
                'e2e': r3['e2e'], 'gpu_launches': int(r3['launches_per_step'] * args.steps), 'clocks': cs.summary(),
                'result': {'mean_jaccard': float(np.nanmean(jac[0] / jac[1]))}}
    else:
        with ClockSampler(ctx.local) as cs:
            r4 = config4_leg(ctx)
        line = {'metric': metric_name(4), 'value': r4['value'], 'unit': 'images/s', 'n_gpus': world, 'steps': r4['steps'], 'warmup': 3,
                'ms_per_step': r4['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': r4['dtype'],
                'data': 'synthetic', 'config': {'workload': WORKLOAD4, 'per_gpu_batch': BATCH, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world},
                'gpu_launches': int(r4['launches_per_step'] * r4['steps']), 'clocks': cs.summary(), 'allreduce': r4['allreduce'], 'loss': r4['loss']}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[2, 3, 4],
                    help='BASELINE.json config that is the headline of the line: 2 FCN8+DAE (default), 3 FC-DenseNet103+DAE, 4 train step')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='strong: a fixed set of 80 images is sharded over the ranks in whole batches')
    ap.add_argument('--sections', default='', help='development runs: comma list of headline,bf16,roofline,cpu,config3,config4,sweep,contextmod,fcn8dae')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='development runs: skip the CPU oracle timing')
    ap.add_argument('--precision', default='mixed', choices=['mixed', 'bf16', 'fp32x3'],
                    help="arithmetic of the headline: mixed (parity-grade, default), bf16 (throughput variant), fp32x3 (every conv fp32-accurate)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
