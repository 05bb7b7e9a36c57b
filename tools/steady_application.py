"""Profiling target: one full DAE application (fills borders) + 2 steady-state applications, eager."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as weights  # noqa: E402

def main(B=10, H=360, W=480, precision='bf16'):
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200 import _kernels as K
    pd = weights.synthetic_dae_params(11, 512, seed=1, out_gain=0.1)
    dae = buildDAE([None], None, 11, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0,
                   n_filters=64, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision=precision)
    net = dae.net
    hs = net.h_spatial(H, W)
    h = K.pack_nchw(torch.relu(torch.randn(B, 512, hs[0], hs[1], device='cuda')), net.h_pad, split=net.split)
    yf = torch.softmax(torch.randn(B, 11, H, W, device='cuda'), 1)
    y = K.pack_nchw(yf, net.y_cpad, split=net.split)
    upd = dict(y=yf, active=torch.ones(B, dtype=torch.int32, device='cuda'),
               norm_acc=torch.zeros(B, dtype=torch.int64, device='cuda'), step=0.05)      # as in the captured loop
    net.logits(h, y, full_down=True, update=upd)
    for _ in range(2):
        net.logits(h, y, full_down=False, update=upd)
    torch.cuda.synchronize()
    print('ok')

if __name__ == '__main__':
    main(precision=sys.argv[1] if len(sys.argv) > 1 else 'bf16')
