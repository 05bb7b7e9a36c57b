"""Debug tool (not collected): per-level comparison of the CUDA backward pass with the fp32 autograd oracle under
teacher-forced masks.  python tests/debug_train_grads.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train as OT
from tests.test_train_gpu import _setup, _dense_to_mask, NCLS

def main():
    from iterative_inference_segm_b200 import _kernels as K
    from iterative_inference_segm_b200.train_dae import DAETrainer
    cuda = torch.device('cuda')
    pd, h, y, L, nm, nk = _setup(cuda, structured=True)
    sigma, lr = 0.5, 1e-3
    acc = [torch.zeros_like(p) for p in pd]
    tap = {}
    loss_o, grads_o, _, _ = OT.train_step(pd, acc, y, h, L, NCLS, 100, lr, noise_main=sigma * nm, noise_mask=None, tap=tap)
    forced = {'masksA': [_dense_to_mask(m, cuda) for m in tap['masksA']], 'zmasks': [_dense_to_mask(z, cuda) for z in tap['zero']],
              'masksB': [_dense_to_mask(m, cuda) for m in tap['masksB']]}
    tr = DAETrainer(NCLS, 512, 100, pd, learning_rate=lr, noise=sigma)
    tr.keep_grads = {}
    tr.forward(K.pack_nchw(h.to(cuda), 512), y.to(cuda), nm.to(cuda), None, forced=forced)
    tr.backward(L.to(cuda))
    torch.cuda.synchronize()
    for p in range(6, 0, -1):
        a = tap['pre_act'][p - 1]
        go = a.grad                                            # [B,C,H,W]
        g_pool, g_a = tr.keep_grads[p]
        gd = g_a.float().cpu().permute(0, 3, 1, 2)[:, :go.shape[1]]
        rel = float((gd - go).norm() / go.norm())
        nz_o, nz_d = float((go != 0).float().mean()), float((gd != 0).float().mean())
        # where do they differ in support?
        only_o = float(((go != 0) & (gd == 0)).float().mean()); only_d = float(((go == 0) & (gd != 0)).float().mean())
        # forward activations
        pooled_d = tr.st['pools'][p - 1].float().cpu().permute(0, 3, 1, 2)[:, :go.shape[1]]
        pooled_o = torch.nn.functional.max_pool2d(torch.relu(a.detach()), 2, 2)
        print('level %d: g_a rel err %.4f | nonzero frac oracle %.4f device %.4f | only-oracle %.5f only-device %.5f | pooled rel err %.4f | pooled>0 mismatch %.5f' % (
            p, rel, nz_o, nz_d, only_o, only_d, float((pooled_d - pooled_o).norm() / pooled_o.norm()),
            float(((pooled_d > 0) != (pooled_o > 0)).float().mean())))

if __name__ == '__main__':
    main()
