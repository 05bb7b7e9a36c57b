#!/usr/bin/env python
"""bench.py -- images/sec of N-step FCN8 + DAE_h iterative inference (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic CamVid-shaped input:
BASELINE.json configs[1] -- batch 10 of 360x480 images, 11 classes: FCN8 forward (h = pool4, y0),
50 iterations of y <- clip(y + step*(DAE(y,h) - y), 0, 1) with step 0.05, then the metrics.py
confusion-matrix / Jaccard reduction.  Every rank processes its own batch (image sharding, weak
scaling); the only collective is the all-reduce of the int64 confusion matrix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : whole-job images/s with the batch already resident in HBM.
`e2e`    : the same through the public callables with HOST (pinned) buffers: H2D of the images and
           the one-hot targets and D2H of the metrics inside the timed region, every step (the copies
           of step i+1 ride a copy stream under step i's kernels, as with a prefetching iterator).
`roofline`: the tcgen05 conv kernel (all conv launches of one DAE application), CUDA-event timed.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference path (oracle/, PyTorch
           CPU fp32; Theano is not installable) on the host cores, on a bounded sample.
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
LOGIT_GAIN, OUT_GAIN = 10.0, 0.1
METRIC = 'images/sec, N-step FCN8+DAE iterative inference 360x480'
WORKLOAD = 'FCN8 + DAE_h (n_filters=64, concat_h=pool4, trackind unpool), batch 10 x 360x480, 11 classes, 50 steps, step 0.05, metrics.py Jaccard'


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# --------------------------------------------------------------------------- CPU reference arm
CPU_SAMPLE_ITERS = 10


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
    return ('1 image 360x480: FCN8 forward + %d of the 50 DAE iterations + metrics timed with the PyTorch-CPU '
            'oracle (port of the Theano path), per-image time extrapolated linearly to 50 iterations' % n)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_reference_sample()
    times = []
    for _ in range(args.steps):
        t_img, cores, parts = cpu_reference_sample()
        times.append(t_img)
    t = sum(times) / len(times)
    val = 1.0 / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * BATCH * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'weights': 'random init (He-uniform FCN8 x logit gain 10, Glorot DAE x out gain 0.1)'},
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': cpu_sample_text(CPU_SAMPLE_ITERS)},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


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

    def _run(self):
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    from iterative_inference_segm_b200.csrc.build import build
    build()
    from iterative_inference_segm_b200 import _lib
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import IterativeInference, jaccard_from_cm
    from iterative_inference_segm_b200.profiling import KernelTimer
    from iterative_inference_segm_b200 import synthetic as weights     # synthetic weight / data recipe

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()

    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=OUT_GAIN)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, precision=args.precision)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=fcn[0].output_shape[1], padding=100, concat_h=['pool4'],
                   noise=0.0, n_filters=64, conv_before_pool=1, additional_pool=2, skip=True, unpool_type='trackind',
                   params=pd, precision=args.precision)
    del pf, pd
    ii = IterativeInference(dae, NCLS, [NCLS])
    fnet = fcn[0].net

    # rank r holds images [r*BATCH, (r+1)*BATCH) of the synthetic set
    X, L, lab = weights.synthetic_batch(BATCH, H, W, NCLS, seed=100 + rank)
    X_host, L_host = X.pin_memory(), L.pin_memory()
    X_dev, L_dev = X_host.to(dev), L_host.to(dev)
    cm_total = torch.zeros(NCLS * NCLS + 2, dtype=torch.int64, device=dev)

    def step_device(Xd, Ld):
        out = fnet.forward(Xd, want=('pool4', 'probs_dimshuffle'))
        res = ii.run(out['pool4'], out['probs_dimshuffle'], STEP, N_ITER, onehot=Ld)
        cm_total[:NCLS * NCLS] = res['cm'].sum(0)
        cm_total[NCLS * NCLS:] = res['counts'].sum(0)
        if world > 1:
            dist.all_reduce(cm_total)
        return res

    # End to end: every step copies its own inputs from pinned host memory and reads its result back.  The copies
    # of step i+1 are issued on a copy stream before step i's kernels (two device buffers, like a prefetching data
    # iterator), so only the first step's H2D is exposed; the D2H read of the result blocks the host every step.
    copy_stream = torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(X_dev), torch.empty_like(L_dev)) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])
            in_bufs[b][0].copy_(X_host, non_blocking=True)
            in_bufs[b][1].copy_(L_host, non_blocking=True)
            ev_ready[b].record(copy_stream)

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        for b in range(2):
            ev_free[b].record(cur)
        issue_h2d(0)
        out = None
        for i in range(n):
            if i + 1 < n:
                issue_h2d(i + 1)
            b = i % 2
            cur.wait_event(ev_ready[b])
            res = step_device(*in_bufs[b])
            ev_free[b].record(cur)
            out = (cm_total.cpu(), res['n_exec'].cpu())
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(n):
            fn()
        e.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([s.elapsed_time(e), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    # ---- warm-up (also captures the CUDA graph) + launch census
    l0 = _lib.launch_count()
    res = step_device(X_dev, L_dev)
    torch.cuda.synchronize()
    census_first = _lib.launch_count() - l0
    for _ in range(max(args.warmup - 1, 0)):
        res = step_device(X_dev, L_dev)
    torch.cuda.synchronize()
    n_exec = res['n_exec'].cpu().tolist()
    assert all(n == N_ITER for n in n_exec), 'early exit in the benchmark: %s' % n_exec
    # launches per step: eager FCN8 + boundary launches are counted live; the loop is a graph replay
    l1 = _lib.launch_count()
    step_device(X_dev, L_dev)
    torch.cuda.synchronize()
    eager_per_step = _lib.launch_count() - l1
    launches_per_step = eager_per_step + ii.graph_kernel_nodes     # eager FCN8/boundary launches + graph kernel nodes

    # ---- timed region: device-resident inputs
    with ClockSampler(local) as clk:
        ms_dev, _ = timed(lambda: step_device(X_dev, L_dev), args.steps)
    clocks = clk.summary()
    # ---- timed region: end to end through host buffers (pinned), H2D + D2H inside
    run_e2e(2)
    ms_e2e, wall_e2e = timed(lambda: run_e2e(args.steps), 1)
    ms_e2e = max(ms_e2e, wall_e2e)     # D2H reads block the host: wall clock covers them

    imgs = BATCH * world * args.steps
    value = imgs / (ms_dev * 1e-3)
    e2e_value = imgs / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: every conv launch of one DAE application, event-timed eagerly
    roof, breakdown = None, None
    cpu_base = None
    if rank == 0 and args.precision == 'bf16':
        st = ii._buffers(BATCH, H, W, N_ITER, False)
        timer = KernelTimer()
        reps = 3
        upd = dict(y=st['y'], active=st['active'], norm_acc=st['norm_acc'], step=STEP)   # as in the captured loop
        st['active'].fill_(1)
        dae.net.logits(st['h'], st['y_bf16'], full_down=True, update=upd)      # fills the iteration-invariant borders
        with timer.recording():
            for _ in range(reps + 1):
                dae.net.logits(st['h'], st['y_bf16'], full_down=False, update=upd)  # the steady-state application (49 of 50)
        summ = timer.summary()
        fl = dae.net.executed_conv_flops(H, W, steady_state=True)               # executed FLOPs only, per image
        convs = [(tag, sum(v[1:]) / len(v[1:])) for (name, tag), v in summ.items() if name == 'conv2d']   # launch order; drop the cold rep
        assert len(convs) == len(fl)
        conv_total_ms = sum(ms for _, ms in convs)
        other = {}
        for (name, tag), v in summ.items():
            other[name] = other.get(name, 0.0) + sum(v[1:]) / len(v[1:])
        # dominant kernel: conv_igemm_pair_kernel<256> (per-tap implicit GEMM, 256 x 256 tiles over CTA pairs): the launches the
        # library reports as kernel 1 (CTA pair) with BN = 256 (tag[-1] = iiseg_last_conv_plan of that launch)
        dom = [(f, ms) for f, (tag, ms) in zip(fl, convs) if tag[-1][0] == 1 and tag[-1][1] == 256]
        dom_flops, dom_ms = sum(f for f, _ in dom) * BATCH, sum(ms for _, ms in dom)
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12
        peak = peaks['bf16_tflops_sustained']
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')      # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get('conv_igemm_pair_kernel<256>', {}).get('dram_bytes_per_launch')
        roof = {'bound': 'tensor', 'kernel': 'conv_igemm_pair_kernel<256> (tcgen05 cta_group::2 implicit GEMM, 256 x 256 tiles over CTA pairs; %d of the 12 conv launches of one steady-state DAE application, batch 10: conv3_1, conv4_1, conv6_1, up_conv6..up_conv4 -- conv5_1 runs the same kernel with 128-wide tiles; executed FLOPs on the y-dependent / crop-dependent windows)' % len(dom),
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                'peak_source': peaks['source'] + ' bf16_tflops_sustained', 'launch_ms': dom_ms / len(dom),
                'flops_per_launch': dom_flops / len(dom),
                'all_12_conv_launches': {'achieved': sum(fl) * BATCH / (conv_total_ms * 1e-3) / 1e12, 'ms': conv_total_ms,
                                         'flops_per_application': sum(fl) * BATCH,
                                         'note': 'includes the 16-channel first layer, the fused 2x2 pool + tie mask epilogues and the fused softmax/update epilogue of up_conv1'}}
        unpool_b = 0.0
        for (name, tag), v in summ.items():
            if name == 'unpool2':       # algorithmic bytes: out written once, the touched u / mask windows read once
                (n, uh, uw, c), (_, oh, ow, _) = tag
                unpool_b += n * oh * ow * c * 2 + n * ((oh + 1) // 2 + 1) * ((ow + 1) // 2 + 1) * c * 2.5
        breakdown = {'ms_per_dae_application': {k: round(v, 4) for k, v in other.items()},
                     'conv_ms_in_launch_order': [round(ms, 4) for _, ms in convs],
                     'conv_tflops_in_launch_order': [round(f * BATCH / (ms * 1e-3) / 1e12, 1) for f, (_, ms) in zip(fl, convs)],
                     'unpool_gbs': unpool_b / (other.get('unpool2', 1e9) * 1e-3) / 1e9,
                     'hbm_peak_gbs': peaks['hbm_gbs'],
                     'note': 'max-pool + tie mask are fused into the contracting-path conv epilogues; softmax + y update + norm into up_conv1'}
        if world == 1 and not args.no_cpu_baseline:
            cpu_reference_sample(2)                                   # warm-up: thread pool, oneDNN primitive caches
            t_img, cores, parts = cpu_reference_sample(25)            # ~7 s of host work
            cpu_base = {'value': 1.0 / t_img, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                        'sample': cpu_sample_text(25), 'parts_s': {k: round(v, 3) for k, v in parts.items()}}

    if rank == 0:
        cm = cm_total[:NCLS * NCLS].cpu().numpy()
        jac = jaccard_from_cm(cm)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'fp32 operands as bf16 hi/lo pairs, 3 tensor-core products, fp32 accumulate', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'per_gpu_batch': BATCH, 'global_batch': BATCH * world,
                       'parallelism': 'image shards, dp%d' % world,
                       'weights': 'random init (He-uniform FCN8 x logit gain 10, Glorot DAE x out gain 0.1)',
                       'l2': 'working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush',
                       'executed_iterations': n_exec[0]},
            'e2e': {'value': e2e_value, 'unit': 'images/s',
                    'h2d_bytes_per_step': int(X_host.numel() * 4 + L_host.numel() * 4),
                    'd2h_bytes_per_step': int(cm_total.numel() * 8 + BATCH * 4), 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches_per_step * args.steps),
            'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu_base, 'breakdown': breakdown,
            'result': {'mean_jaccard': float(__import__('numpy').nanmean(jac[0] / jac[1])), 'first_step_launch_census': int(census_first)},
        }
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
    ap.add_argument('--no-cpu-baseline', action='store_true', help='development runs: skip the CPU oracle timing')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32x3'],
                    help="fp32x3: the parity-grade variant (fp32 operands as bf16 hi/lo pairs, three tensor-core products); no roofline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
