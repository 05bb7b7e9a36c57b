"""CPU tests of the host-side logic: synthetic data iterator contract, experiment naming,
print_results arithmetic, shard ranges, and the N>1 metric reduction with gloo (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_data_iterator_contract():
    from iterative_inference_segm_b200.data_loader import load_data
    it = load_data('camvid', {}, one_hot=True, batch_size=[10, 5, 10], which_set='test', n_images=12, height=8, width=10)
    assert it.nbatches == 2 and it.non_void_nclasses == 11 and it.void_labels == [11] and it.data_shape == (3, 8, 10)
    X, L = it.next()
    assert X.shape == (10, 3, 8, 10) and X.dtype == np.float32 and 0 <= X.min() and X.max() < 1
    assert L.shape == (10, 12, 8, 10) and np.array_equal(L.sum(1), np.ones((10, 8, 10), np.float32))
    X2, _ = it.next()
    assert X2.shape[0] == 2
    assert len(it.mask_labels) == 12 and it.cmap.shape == (12, 3)


def test_experiment_name_matches_reference_rules():
    from iterative_inference_segm_b200.helpers import build_experiment_name
    name = build_experiment_name('fcn8', kind='standard', concat_h=['pool4'], n_filters=64, conv_before_pool=1,
                                 additional_pool=2, skip=True, unpool_type='trackind', dropout=0, noise=0.5,
                                 from_gt=False, temperature=1.0, training_loss=['crossentropy', 'squared_error'],
                                 learning_rate=0.001, lr_anneal=0.99, weight_decay=0.0001, optimizer='rmsprop',
                                 data_aug=True, exp_name='flip_final_', layer='probs_dimshuffle', bn=0)
    assert name == ('flip_final_fcn8_standard_pool4_f64c1p2_skip_trackind_crossentropy_squared_error_fromfcn8_z0.5'
                    '_data_aug_T1.0_rmsprop_lr0.001_anneal0.99_decay0.0001_probs_dimshuffle')


def test_results_values():
    from iterative_inference_segm_b200.helpers import results_values
    jacc = np.array([[1., 0., 2.], [2., 0., 4.]], np.float32)      # class 1 absent -> nan, ignored by nanmean
    loss, acc, jm = results_values(3.0, 1.5, jacc, 3)
    assert loss == 1.0 and acc == 0.5 and abs(jm - 0.5) < 1e-7


def test_shard_range_partitions():
    from iterative_inference_segm_b200.sharding import shard_range
    for n in (1, 7, 10, 233):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from iterative_inference_segm_b200.sharding import shard_range, allreduce_metrics
from oracle import metrics as M
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
rng = np.random.RandomState(0)
N, C, H, W = 7, 11, 6, 8
y = rng.rand(N, C, H, W).astype(np.float32)
lab = rng.randint(0, C + 1, size=(N, H, W))
onehot = np.eye(C + 1, dtype=np.float32)[lab].transpose(0, 3, 1, 2)
lo, hi = shard_range(N, rank, world)
cm = torch.from_numpy(M.confusion_matrix(y[lo:hi], onehot[lo:hi], C))
c, v = M.accuracy_counts(y[lo:hi], onehot[lo:hi], [C])
counts = torch.tensor([c, v], dtype=torch.int64)
allreduce_metrics(cm, counts)
full = M.confusion_matrix(y, onehot, C)
cf, vf = M.accuracy_counts(y, onehot, [C])
assert np.array_equal(cm.numpy(), full), 'confusion matrix differs after all-reduce'
assert counts.tolist() == [cf, vf]
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_two_rank_metric_allreduce_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29543')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29543', str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('ok') == 2


def test_synthetic_recipe_matches_oracle_copy():
    """bench.py draws its synthetic weights / data from the product package; the oracle keeps an
    independent copy.  Both must produce identical tensors or parity tests and the bench diverge."""
    import torch
    from oracle import weights as OW
    from iterative_inference_segm_b200 import synthetic as S
    for a, b in zip(S.synthetic_fcn8_params(3, 11, seed=0, logit_gain=10.0), OW.synthetic_fcn8_params(3, 11, seed=0, logit_gain=10.0)):
        assert torch.equal(a, b)
    for a, b in zip(S.synthetic_dae_params(11, 512, seed=1, out_gain=0.1), OW.synthetic_dae_params(11, 512, seed=1, out_gain=0.1)):
        assert torch.equal(a, b)
    for a, b in zip(S.synthetic_contextmod_params(11, 3, seed=3), OW.synthetic_contextmod_params(11, 3, seed=3)):
        assert torch.equal(a, b)
    for a, b in zip(S.synthetic_batch(2, 12, 16, 11, seed=7), OW.synthetic_batch(2, 12, 16, 11, seed=7)):
        assert torch.equal(a, b)
    from oracle import densenet as OD
    pa, pb = S.synthetic_densenet_params(3, 11, seed=2, logit_gain=4.0), OD.synthetic_densenet_params(3, 11, seed=2, logit_gain=4.0)
    assert len(pa) == len(pb) == 590
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


_TRAIN_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from iterative_inference_segm_b200.sharding import World, shard_range
from oracle import train as OT, weights
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
W_ = World()
assert (W_.rank, W_.size) == (rank, world)
torch.set_num_threads(2)
NCLS, B, H, Wd = 11, 4, 8, 10
pd = weights.synthetic_dae_params(NCLS, 8, seed=1, n_filters=8, out_gain=0.1)
gen = torch.Generator().manual_seed(0)
L = torch.nn.functional.one_hot(torch.randint(0, NCLS + 1, (B, H, Wd), generator=gen), NCLS + 1).permute(0, 3, 1, 2).float()
y = L[:, :NCLS].contiguous()
h = torch.rand((B, 8, 12, 13), generator=gen)          # pool4-sized conditioning for an 8x10 image with pad 100 (206x208 -> 12x13)
noise = 0.5 * torch.randn(y.shape, generator=gen)
# single-device step on the whole batch
ps = [p.clone().requires_grad_(True) for p in pd]
loss_full = OT.loss_fn(OT.dae_forward_train(ps, y + noise, h, 100), L, NCLS)
g_full = torch.autograd.grad(loss_full, ps)
# data-parallel: local numerators over GLOBAL denominators (two all-reduces), then SUM of the gradients
lo, hi = shard_range(B, rank, world)
ps = [p.clone().requires_grad_(True) for p in pd]
logits = OT.dae_forward_train(ps, (y + noise)[lo:hi], h[lo:hi], 100)
p = torch.softmax(logits, 1)
t = L[lo:hi]
true = t.argmax(1); mask = (true != NCLS).float()
ce = -torch.log(torch.clamp(p, 1e-7, 1 - 1e-7).gather(1, (true * mask.long()).unsqueeze(1))).squeeze(1)
m2 = t[:, :NCLS].sum(1)
se = ((p - t[:, :NCLS]) ** 2).mean(1)
den = torch.tensor([float(mask.sum()), float(m2.sum())], dtype=torch.float64)
W_.allreduce_sum(den)                                   # the loss is a masked mean over the global batch
loss_local = (ce * mask).sum() / den[0].float() + (se * m2).sum() / den[1].float()
g = [x.clone() for x in torch.autograd.grad(loss_local, ps)]
for x in g:
    W_.allreduce_sum(x)
for a, b in zip(g, g_full):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-8), float((a - b).abs().max())
# the same with every loss term of the device step: + dice (three whole-batch sums I, T, P all-reduced with the denominators;
# each rank then differentiates with the GLOBAL sums as constants: dL/dp_1 = -(2 t (T + P + 1) - (2 I + 1)) / (T + P + 1)^2) and
# + ae_h (sum of squares and element count all-reduced: d/dc = 2 c / n_global) -- iiseg_loss_grad_terms / iiseg_sq_sum protocol
ps = [p.clone().requires_grad_(True) for p in pd]
ae = {}
loss_full = OT.loss_fn(OT.dae_forward_train(ps, y + noise, h, 100, ae_out=ae), L, NCLS, use_dice=True) + OT.ae_h_loss(ae)
g_full = torch.autograd.grad(loss_full, ps)
ps = [p.clone().requires_grad_(True) for p in pd]
ae = {}
logits = OT.dae_forward_train(ps, (y + noise)[lo:hi], h[lo:hi], 100, ae_out=ae)
p = torch.softmax(logits, 1)
ce = -torch.log(torch.clamp(p, 1e-7, 1 - 1e-7).gather(1, (true * mask.long()).unsqueeze(1))).squeeze(1)
se = ((p - t[:, :NCLS]) ** 2).mean(1)
t1, p1 = t[:, 1].to(torch.int32).float(), p[:, 1]
c = ae['h_hat'] - ae['h']
sums = torch.tensor([float(mask.sum()), float(m2.sum()), float((t1 * p1).sum()), float(t1.sum()), float(p1.sum()),
                     float((c * c).sum()), float(c.numel())], dtype=torch.float64)
W_.allreduce_sum(sums)
n_ce, n_mse, I, T, P, sq, cnt = [float(v) for v in sums]
S = T + P + 1.0
loss_global = None          # every rank can form the global loss from the reduced sums (what DAETrainer.loss_value does)
coef = (-(2.0 * t1 * S - (2.0 * I + 1.0)) / (S * S)).detach()
surrogate = (ce * mask).sum() / n_ce + (se * m2).sum() / n_mse + (coef * p1).sum() + (c * c).sum() / cnt
g = [x.clone() for x in torch.autograd.grad(surrogate, ps)]
for x in g:
    W_.allreduce_sum(x)
for a, b in zip(g, g_full):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-8), float((a - b).abs().max())
num = torch.tensor([float((ce * mask).sum()), float((se * m2).sum())], dtype=torch.float64)
W_.allreduce_sum(num)
loss_global = float(num[0]) / n_ce + float(num[1]) / n_mse - (2.0 * I + 1.0) / S + sq / cnt
assert abs(loss_global - float(loss_full)) < 1e-5 * abs(float(loss_full)), (loss_global, float(loss_full))
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_two_rank_train_step_allreduce_gloo(tmp_path):
    """Data-parallel DAE train step protocol (config 4): global loss denominators + SUM of per-rank gradients equals
    the single-device gradient on the concatenated batch (sharding.World over gloo, autograd oracle as the model); also with
    the dice term's whole-batch sums and the ae_h term's sum of squares / element count riding the same all-reduce."""
    script = tmp_path / 'train_worker.py'
    script.write_text(_TRAIN_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29547')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29547', str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('ok') == 2


def test_error_behaviour_matches_reference():
    """The reference raises ValueError for an unknown segmentation net / DAE kind and for a missing save directory
    (iterative_inference.py:51,147,179); the drop-in raises the same exceptions before touching the device."""
    import pytest
    from iterative_inference_segm_b200.iterative_inference import build_networks, inference, DAE_DICT_DEFAULTS
    with pytest.raises(ValueError):
        build_networks('resnet', dict(DAE_DICT_DEFAULTS), 11, 3, [11])
    with pytest.raises(ValueError):
        inference('camvid', 'fcn8', 0.05, 3, savepath=None)
    from iterative_inference_segm_b200.data_loader import load_data
    with pytest.raises(NotImplementedError):
        load_data('camvid', {}, one_hot=False)


def test_product_path_has_no_cpu_fallback():
    """Kernel wrappers refuse CPU tensors: there is no silent host path behind the C ABI."""
    import pytest
    import torch
    from iterative_inference_segm_b200 import _kernels as K
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16)
    with pytest.raises((AssertionError, RuntimeError, ValueError, TypeError)):
        K.maxpool2(x, True)


def test_cli_accepts_the_reference_flags(monkeypatch):
    """Every flag of the reference's CLI (iterative_inference.py:329-394) parses and reaches inference()."""
    import sys
    from iterative_inference_segm_b200 import iterative_inference as II
    seen = {}
    monkeypatch.setattr(II, 'inference', lambda *a, **k: seen.update(args=a, kw=k))
    monkeypatch.setattr(sys, 'argv', ['iterative_inference.py', '-dataset', 'camvid', '-segmentation_net', 'fcn8', '-step', '0.05',
                                      '-ne', '50', '-which_set', 'val', '-dae_dict', "{'kind': 'standard', 'noise': 0}",
                                      '-training_dict', "{'optimizer': 'rmsprop'}", '-full_im_ft', 'False', '-ae_h', 'False',
                                      '-data_augmentation', 'True', '-test_from_0_255', 'False'])
    II.main()
    assert seen['args'] == ('camvid', 'fcn8', 0.05, 50)
    assert seen['kw']['which_set'] == 'val' and seen['kw']['dae_dict_updates'] == {'kind': 'standard', 'noise': 0}
    assert seen['kw']['full_im_ft'] is False and seen['kw']['data_augmentation'] is True


def test_bench_clock_sampler_summary():
    """bench.ClockSampler: median SM clock, throttle reasons, instantaneous power (when the driver reports it)."""
    import bench
    cs = bench.ClockSampler(0)
    cs.samples = [['1710', '1965', 'Not Active', 'Not Active', 'Not Active', 'Active', '995.1', '1000.00'],
                  ['1725', '1965', 'Not Active', 'Not Active', 'Not Active', 'Active', '988.0', '1000.00'],
                  ['1717', '1965', 'Not Active', 'Not Active', 'Not Active', 'Not Active', '640.2', '1000.00']]
    out = cs.summary()
    assert out['sm_mhz'] == 1717 and out['sm_max_mhz'] == 1965 and out['reasons'] == ['sw_power_cap'] and out['samples'] == 3
    assert out['power_w'] == 988.0 and out['power_max_w'] == 995.1 and out['power_limit_w'] == 1000.0
    cs.samples = [['1965', '1965', 'Not Active', 'Not Active', 'Not Active', 'Not Active']]      # driver without power fields
    out = cs.summary()
    assert out['sm_mhz'] == 1965 and out['reasons'] == [] and 'power_w' not in out
    cs.samples = []
    assert cs.summary()['reasons'] == ['unsampled']


def test_camvid_directory_reader(tmp_path):
    """The on-disk CamVid reader (layout of dataset_loaders' CamvidDataset: <set>/NAME.png + <set>annot/NAME.png) keeps the
    iterator contract the scripts rely on (iterative_inference.py:117-125, 233-234) and shards whole batches."""
    from PIL import Image
    from iterative_inference_segm_b200.data_loader import load_data, CamvidDirectoryIterator
    rng = np.random.RandomState(0)
    root = tmp_path / 'camvid'
    imgs, labs = [], []
    for s, n in (('test', 5), ('train', 3)):
        os.makedirs(root / s); os.makedirs(root / (s + 'annot'))
        for i in range(n):
            img = rng.randint(0, 256, size=(12, 16, 3)).astype(np.uint8)
            lab = rng.randint(0, 12, size=(12, 16)).astype(np.uint8)
            Image.fromarray(img).save(str(root / s / ('f%02d.png' % i)))
            Image.fromarray(lab).save(str(root / (s + 'annot') / ('f%02d.png' % i)))
            if s == 'test':
                imgs.append(img); labs.append(lab)
    it = load_data('camvid', {}, one_hot=True, batch_size=[10, 5, 2], which_set='test', path=str(root))
    assert it.nbatches == 3 and it.non_void_nclasses == 11 and it.void_labels == [11] and it.data_shape == (3, 12, 16)
    assert len(it.cmap) == 12 and len(it.mask_labels) == 12
    X, L = it.next()
    assert X.shape == (2, 3, 12, 16) and X.dtype == np.float32 and L.shape == (2, 12, 12, 16) and L.dtype == np.float32
    assert np.array_equal(X[0], imgs[0].transpose(2, 0, 1).astype(np.float32) / np.float32(255.0))
    assert np.array_equal(L[1].argmax(0), labs[1]) and np.all(L.sum(1) == 1)
    it.next()
    X3, _ = it.next()
    assert X3.shape[0] == 1                                  # the last, short batch
    X4, _ = it.next()
    assert np.array_equal(X4, X)                             # next epoch starts over in the same order
    it255 = load_data('camvid', {}, one_hot=True, batch_size=[10, 5, 2], which_set='test', path=str(root), return_0_255=True)
    assert float(it255.next()[0].max()) > 1.0
    # training-time augmentation: random crop + flip, labels stay aligned with the pixels
    tr = load_data('camvid', {'crop_size': (8, 10), 'horizontal_flip': 0.5}, one_hot=True, batch_size=[3, 1, 1], which_set='train',
                   path=str(root), seed=3)
    Xc, Lc = tr.next()
    assert Xc.shape == (3, 3, 8, 10) and Lc.shape == (3, 12, 8, 10) and tr.data_shape == (3, 8, 10)
    # shards: two ranks see disjoint batches that together are the whole set, in order
    a = CamvidDirectoryIterator(str(root), 'test', 2, shard=(0, 2), use_threads=False)
    b = CamvidDirectoryIterator(str(root), 'test', 2, shard=(1, 2), use_threads=False)
    assert a.nbatches + b.nbatches == 3
    got = [a.next()[0] for _ in range(a.nbatches)] + [b.next()[0] for _ in range(b.nbatches)]
    assert np.array_equal(np.concatenate(got), np.stack([im.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0) for im in imgs]))
    with pytest.raises(IOError):
        load_data('camvid', {}, one_hot=True, which_set='val', path=str(root))


def test_trainer_graph_identity_includes_hyperparameters():
    """lr / sigma / lmb / rho / eps are baked into a captured graph, so they are part of its identity."""
    import inspect
    from iterative_inference_segm_b200 import train_dae
    src = inspect.getsource(train_dae.DAETrainer.step_graphed)
    assert 'self.lr' in src and 'self.sigma' in src and 'ent[2] != hp' in src
    assert hasattr(train_dae, 'train') and hasattr(train_dae, 'main') and hasattr(train_dae.DAETrainer, 'step_dp')
