"""Experiment: the 50-iteration loop of a batch of 10 as ONE graph on one stream, or as two half-batches of 5 captured on two
streams that run concurrently (the tail round of one half's kernel overlaps the other half's kernels).
    python tools/two_stream_experiment.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(precision='mixed', B=10, H=360, W=480, N=50):
    from iterative_inference_segm_b200 import synthetic
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import IterativeInference
    pd = synthetic.synthetic_dae_params(11, 512, seed=1, out_gain=0.1)
    mk = lambda: buildDAE([None], None, 11, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,      # noqa: E731
                          n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision=precision)
    daes = [mk(), mk(), mk()]
    iis = [IterativeInference(d, 11, [11]) for d in daes]
    hs = daes[0].net.h_spatial(H, W)
    h = torch.relu(torch.randn(B, 512, hs[0], hs[1], device='cuda'))
    y0 = torch.softmax(torch.randn(B, 11, H, W, device='cuda') * 3, 1)
    lab = torch.randint(0, 12, (B, H, W), device='cuda', dtype=torch.int32)

    def ev(fn, n=5):
        fn(); fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    one = lambda: iis[0].run(h, y0, 0.05, N, eps=0.0, labels=lab)           # noqa: E731
    t1 = ev(one)
    y_one = iis[0].run(h, y0, 0.05, N, eps=0.0, labels=lab)['y'].clone()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    half = B // 2
    parts = [(h[:half].contiguous(), y0[:half].contiguous(), lab[:half].contiguous()), (h[half:].contiguous(), y0[half:].contiguous(), lab[half:].contiguous())]

    def two():
        cur = torch.cuda.current_stream()
        evs = []
        for k in range(2):
            streams[k].wait_stream(cur)
            with torch.cuda.stream(streams[k]):
                iis[1 + k].run(*parts[k][:2], 0.05, N, eps=0.0, labels=parts[k][2])
        for k in range(2):
            cur.wait_stream(streams[k])
    t2 = ev(two)
    two(); torch.cuda.synchronize()
    y_two = torch.cat([iis[1]._state[list(iis[1]._state)[0]]['y'], iis[2]._state[list(iis[2]._state)[0]]['y']])
    print('precision %s: one graph (batch %d) %.2f ms; two half-batch graphs on two streams %.2f ms (%.1f %%); y identical: %s' % (
        precision, B, t1, t2, 100 * (t1 - t2) / t1, torch.equal(y_one, y_two)))
    # sequential halves for reference (no overlap)
    t3 = ev(lambda: (iis[1].run(*parts[0][:2], 0.05, N, eps=0.0, labels=parts[0][2]), iis[2].run(*parts[1][:2], 0.05, N, eps=0.0, labels=parts[1][2])))
    print('two half-batch graphs back to back on one stream: %.2f ms' % t3)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'mixed')
