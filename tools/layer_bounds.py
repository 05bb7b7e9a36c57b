"""Per-launch times of one steady-state DAE application (batch 10, 360x480), CUDA-event timed.
Run under IISEG_CONV_DBG=0|1|2|3 (bit0: no activation/weight loads, bit1: no MMAs) to see what
bounds each conv layer:  python tools/layer_bounds.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(B=10, H=360, W=480, precision='bf16'):
    from iterative_inference_segm_b200 import synthetic, _kernels as K
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.profiling import KernelTimer
    pd = synthetic.synthetic_dae_params(11, 512, seed=1, out_gain=0.1)
    dae = buildDAE([None], None, 11, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,
                   n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision=precision)
    net = dae.net
    hs = net.h_spatial(H, W)
    h = K.pack_nchw(torch.relu(torch.randn(B, 512, hs[0], hs[1], device='cuda')), net.h_pad, split=net.split)
    y = K.pack_nchw(torch.softmax(torch.randn(B, 11, H, W, device='cuda'), 1), net.y_cpad, split=net.split)
    yf = torch.softmax(torch.randn(B, 11, H, W, device='cuda'), 1)
    upd = None if (net.split_up or os.environ.get('UNFUSED')) else dict(y=yf, active=torch.ones(B, dtype=torch.int32, device='cuda'),
                                     norm_acc=torch.zeros(B, dtype=torch.int64, device='cuda'), step=0.05)
    net.logits(h, y, full_down=True)
    timer = KernelTimer()
    with timer.recording():
        for _ in range(4):
            net.logits(h, y, full_down=False, update=upd)
    summ = timer.summary()
    fl = net.executed_conv_flops(H, W, True)
    tot, i = 0.0, 0
    for (name, tag), v in summ.items():
        ms = sum(v[1:]) / len(v[1:])
        tot += ms
        extra = ''
        if name == 'conv2d':
            extra = '%7.1f TFLOP/s' % (fl[i] * B / ms / 1e9)
            i += 1
        print('%-8s %-70s %8.1f us %s' % (name, str(tag)[:70], ms * 1e3, extra))
    print('total %.1f us  [dbg=%s precision=%s]' % (tot * 1e3, os.environ.get('IISEG_CONV_DBG'), precision))


if __name__ == '__main__':
    main(precision=sys.argv[1] if len(sys.argv) > 1 else 'bf16')
