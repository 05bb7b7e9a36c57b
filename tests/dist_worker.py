"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU, NCCL): runs the packaged
drop-in entry points on this rank's shard and writes rank 0's reduced results.  Not collected by pytest."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NCLS = 11
DAE_DICT = {'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0, 'concat_h': ['pool4'],
            'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1, 'additional_pool': 2, 'path_weights': '',
            'layer': 'probs_dimshuffle', 'exp_name': 'flip_final_', 'bn': 0}


def run(out_dir, distributed):
    from iterative_inference_segm_b200 import synthetic as S, _kernels as K
    from iterative_inference_segm_b200.data_loader import SyntheticSegmentationIterator
    from iterative_inference_segm_b200.iterative_inference import inference
    from iterative_inference_segm_b200.iterative_inference_valid import sweep
    from iterative_inference_segm_b200.sharding import World, shard_range
    from iterative_inference_segm_b200.train_dae import DAETrainer
    pf = S.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = S.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    mk = lambda: SyntheticSegmentationIterator(7, 2, 32, 40, NCLS, seed=5)          # noqa: E731  (4 batches, the last one short)
    kw = dict(dae_dict_updates=DAE_DICT, fcn_params=pf, dae_params=pd, verbose=False)
    a = inference('camvid', 'fcn8', 0.05, 3, data_iter=mk(), savepath=out_dir, loadpath=out_dir, **kw)
    res, mats = sweep('camvid', 'fcn8', steps=[0.05, 0.5], num_iter=2, data_iter=mk(), **kw)
    # data-parallel train step: bucketed + overlapped all-reduce == one blocking all-reduce, bit for bit
    world = World()
    B, H, W = 4, 64, 80
    _, L, _ = S.synthetic_batch(B, H, W, NCLS, seed=5)
    L = L.cuda(); y = L[:, :NCLS].contiguous()
    gen = torch.Generator(device='cuda').manual_seed(3)
    hs = (((H + 198) // 2 // 2 // 2) // 2, ((W + 198) // 2 // 2 // 2) // 2)
    h = K.pack_nchw(torch.relu(torch.randn((B, 512) + hs, device='cuda', generator=gen)), 512)
    nm = torch.randn(y.shape, device='cuda', generator=gen); nk = torch.randn(y.shape, device='cuda', generator=gen)
    lo, hi = shard_range(B, world.rank, world.size)
    sl = lambda t: t[lo:hi].contiguous()                                            # noqa: E731
    outs = []
    for mode in ('blocking', 'bucketed'):
        tr = DAETrainer(NCLS, 512, 100, pd, learning_rate=1e-3, noise=0.5)
        tr.BUCKET_BYTES = 8 << 20
        w = world if distributed else None
        if mode == 'blocking' or not distributed:
            tr.step(sl(h), sl(y), sl(L), sl(nm), sl(nk), world=w)
        else:
            tr.step_dp(sl(h), sl(y), sl(L), sl(nm), sl(nk), world=w)
        torch.cuda.synchronize()
        outs.append(([p.cpu().numpy() for p in tr.params()], tr.loss_value()))
    # adam + the dice term: its three whole-batch sums are all-reduced with the loss denominators between the two loss passes
    tr = DAETrainer(NCLS, 512, 100, pd, learning_rate=1e-3, noise=0.5, optimizer='adam', training_loss=('crossentropy', 'dice', 'squared_error'))
    tr.step(sl(h), sl(y), sl(L), sl(nm), sl(nk), world=world if distributed else None)
    torch.cuda.synchronize()
    loss_ad, w_ad = tr.loss_value(), tr.params()[-2].cpu().numpy()
    if world.rank == 0:
        np.savez(os.path.join(out_dir, 'result_%d.npz' % world.size), loss_adam_dice=np.array([loss_ad]), w_last_adam_dice=w_ad, cm=a['cm'], jacc=a['jacc_tot'], jacc_fcn=a['jacc_tot_fcn'],
                 it=np.array(a['iterative'], dtype=np.float64), n_exec=np.array(a['n_exec']), mats=mats, res=res,
                 loss=np.array([outs[0][1], outs[1][1]]),
                 same=np.array([all(np.array_equal(x, z) for x, z in zip(outs[0][0], outs[1][0]))]),
                 w0=outs[0][0][0], w_last=outs[0][0][-2], w0_init=np.asarray(pd[0]), w_last_init=np.asarray(pd[-2]))


if __name__ == '__main__':
    out_dir = sys.argv[1]
    distributed = 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1
    if distributed:
        local = int(os.environ['LOCAL_RANK'])
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    run(out_dir, distributed)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
