"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCES (read from /root/reference, unmodified) in this
container -- outputs of the reference itself, not of the oracle restatement.

    python -m tests.golden.make_reference_golden            # needs /root/reference; the fixtures it writes are committed

How: oracle/refrun installs stand-ins for `theano` and `lasagne` (neither is installable here: Python 2, README.md:17-22) and
an import hook that loads the reference's Python-2 modules in memory.  This script then calls the reference's drivers,

    iterative_inference.py:inference()            (FCN8 -> pred_dae_fn -> per-image loop -> val_fn; writes batch<i>.npz)
    iterative_inference_valid.py:inference()      (the same loop + valid_mat; writes iterations<step>.npz)

which build the nets with models/fcn8.py:buildFCN8, models/DAE_h.py:buildDAE (fcn_down / fcn_up / model_helpers /
layers/mylayers.py), models/contextmod_dae.py:buildDAE_contextmod, models/fcn8_dae.py:buildFCN8_DAE, load the checkpoints with
lasagne.layers.set_all_param_values, compile pred_fcn_fn / pred_dae_fn / de_fn / val_fn (metrics.py) and run the loop body.
What is supplied from outside the reference: the dataset iterator (dataset_loaders is an absent dependency; a seeded
synthetic CamVid-shaped iterator with the attributes iterative_inference.py:117-125 reads) and the checkpoints (seeded
synthetic weights from oracle/weights.py, saved positionally as np.savez(*params), the format train_dae.py:436-445 writes).
Nothing from /root/reference is copied; only the arrays the reference computed are stored.

Each fixture stores the case description (json), the arrays the reference saved (`Y_ii`, `Y_fcn`, `valid_mat`, `res`) and its
captured stdout (per-iteration `rec acc jaccard` lines and the print_results blocks).  Inputs and weights are regenerated
from the seeds in the description.
"""
import contextlib
import getpass
import io
import json
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
WORK = os.path.join(HERE, '_work')
NCLS = 11

TRAINING_DICT = {'training_loss': ['crossentropy', 'squared_error'], 'learning_rate': 0.001, 'lr_anneal': 0.99,
                 'weight_decay': 0.0001, 'optimizer': 'rmsprop'}


def dae_dict(**kw):
    d = {'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0, 'concat_h': ['pool4'],
         'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1, 'additional_pool': 2, 'path_weights': '',
         'layer': 'probs_dimshuffle', 'exp_name': 'ref_', 'bn': 0}
    d.update(kw)
    return d


# name -> description.  `dae` = dae_dict passed to the reference; `weights` = how the synthetic checkpoint is drawn
# (tests/test_oracle.py:_reference_case_params rebuilds it from this); sizes are small so that the whole file regenerates in
# a few minutes on CPU and each fixture replays in seconds.
CASES = {
    'ref_standard_trackind': dict(script='inference', dae=dae_dict(), H=32, W=40, B=2, nbatches=2, num_iter=6, step=0.05,
                                  weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_valid_trackind': dict(script='valid', dae=dae_dict(), H=37, W=45, B=2, nbatches=2, num_iter=5, step=0.1,
                               weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_early_exit': dict(script='valid', dae=dae_dict(), H=32, W=40, B=2, nbatches=1, num_iter=8, step=0.5,
                           weights=dict(fn='dae', seed=1, out_gain=0.002)),
    'ref_unpool_standard': dict(script='inference', dae=dae_dict(unpool_type='standard'), H=32, W=40, B=2, nbatches=1, num_iter=4,
                                step=0.05, weights=dict(fn='dae', seed=4, out_gain=0.1)),
    'ref_unpool_inverse': dict(script='inference', dae=dae_dict(unpool_type='inverse'), H=32, W=40, B=2, nbatches=1, num_iter=4,
                               step=0.05, weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_noskip': dict(script='inference', dae=dae_dict(skip=False), H=32, W=40, B=2, nbatches=1, num_iter=4, step=0.05,
                       weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_bn': dict(script='inference', dae=dae_dict(bn=1), H=37, W=45, B=2, nbatches=1, num_iter=4, step=0.05,
                   weights=dict(fn='dae', seed=1, out_gain=0.1, bn_seed=21)),
    'ref_conv_before_pool2': dict(script='inference', dae=dae_dict(conv_before_pool=2), H=32, W=40, B=2, nbatches=1, num_iter=4,
                                  step=0.05, weights=dict(fn='dae', seed=6, out_gain=0.1)),
    'ref_concat_input': dict(script='inference', dae=dae_dict(concat_h=['input'], additional_pool=3), H=32, W=40, B=2, nbatches=1,
                             num_iter=4, step=0.05, weights=dict(fn='dae', seed=8, out_gain=0.1)),
    'ref_pool3': dict(script='inference', dae=dae_dict(concat_h=['pool3'], additional_pool=1), H=32, W=40, B=2, nbatches=1, num_iter=4,
                      step=0.05, weights=dict(fn='dae', seed=9, out_gain=0.1, nb_h=256)),
    'ref_contextmod': dict(script='inference', dae=dae_dict(kind='contextmod', concat_h=['input']), H=32, W=40, B=2, nbatches=1,
                           num_iter=4, step=0.05, weights=dict(fn='contextmod', seed=3)),
    'ref_fcn8_dae': dict(script='inference', dae=dae_dict(kind='fcn8', concat_h=['pool4']), H=32, W=40, B=1, nbatches=1, num_iter=3,
                         step=0.05, weights=dict(fn='fcn8_dae', seed=6, logit_gain=10.0)),
    'ref_temperature': dict(script='fcn8_only', temperature=2.5, H=32, W=40, B=2, nbatches=1),
    # BASELINE.json configs[1], one image of it (the reference iterates image by image): 360x480, 50 iterations, step 0.05.
    # The fixture keeps the argmax labels in full and every third pixel of the probabilities (`sub`).
    'ref_full_size': dict(script='inference', dae=dae_dict(), H=360, W=480, B=1, nbatches=1, num_iter=50, step=0.05, sub=3,
                          weights=dict(fn='dae', seed=1, out_gain=0.1)),
    # noise > 0 (the valid script's default 0.5): every DePool2D's mask sub-graph draws its own Gaussian noise, even at inference
    # (layers/mylayers.py:91-93).  The stand-in's draws are logged, so the fixture records which draw fed which DePool2D level
    # in which function call (`noise_k`) and a test regenerates the same numbers (oracle/refrun/stubs/theano/sandbox/rng_mrg.py).
    'ref_noise': dict(script='inference', dae=dae_dict(noise=0.5), H=32, W=40, B=2, nbatches=1, num_iter=3, step=0.05,
                      weights=dict(fn='dae', seed=1, out_gain=0.1)),
    # FC-DenseNet103 conditioning (models/FCDenseNet.py:Network / build_fcdensenet are the reference's; its four layer helpers
    # come from the absent FC_DenseNet package and are restated in oracle/refrun/stubs/FC_DenseNet/layers.py)
    'ref_densenet': dict(script='inference', segm_net='densenet', dae=dae_dict(), H=64, W=96, B=2, nbatches=1, num_iter=3, step=0.05,
                         weights=dict(fn='dae', seed=1, out_gain=0.1, nb_h=464), densenet=dict(seed=2, logit_gain=4.0, bn_seed=5)),
    # train_dae.py:train(): two epochs of two rmsprop steps each (lr annealed in between) + validation, resumed from a seeded
    # checkpoint; noise = 0 (the MRG stream is not reproduced)
    # ... and with noise = 0.5 (BASELINE config 4): the stand-in's draws are logged, so the fixture records which draw was the
    # main GaussianNoiseLayer's and which fed which DePool2D, in training AND in validation (whose masks are noised too)
    'ref_train_noise': dict(script='train', dae=dae_dict(noise=0.5), H=32, W=40, B=2, nbatches=2, val_nbatches=1, num_epochs=2,
                            learning_rate=0.001, lr_anneal=0.99, lmb=1, training_loss=['crossentropy', 'squared_error'],
                            weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_train_adam': dict(script='train', optimizer='adam', dae=dae_dict(), H=32, W=40, B=2, nbatches=2, val_nbatches=1, num_epochs=2,
                           learning_rate=0.001, lr_anneal=0.99, lmb=1, training_loss=['crossentropy', 'squared_error'],
                           weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_train_aeh': dict(script='train', ae_h=True, dae=dae_dict(), H=32, W=40, B=2, nbatches=2, val_nbatches=1, num_epochs=2,
                          learning_rate=0.001, lr_anneal=0.99, lmb=1, training_loss=['crossentropy', 'squared_error'],
                          weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_train_dice': dict(script='train', dae=dae_dict(), H=32, W=40, B=2, nbatches=2, val_nbatches=1, num_epochs=2,
                           learning_rate=0.001, lr_anneal=0.99, lmb=1, training_loss=['crossentropy', 'dice', 'squared_error'],
                           weights=dict(fn='dae', seed=1, out_gain=0.1)),
    'ref_train': dict(script='train', dae=dae_dict(), H=32, W=40, B=2, nbatches=2, val_nbatches=1, num_epochs=2, learning_rate=0.001,
                      lr_anneal=0.99, lmb=1, training_loss=['crossentropy', 'squared_error'], weights=dict(fn='dae', seed=1, out_gain=0.1)),
}

FCN8_WEIGHTS = dict(seed=0, logit_gain=10.0)


def case_dae_params(case):
    """The synthetic DAE checkpoint of a case, as a list of torch tensors in the reference's positional order."""
    from oracle import weights
    w, d = case['weights'], case['dae']
    if w['fn'] == 'contextmod':
        return weights.synthetic_contextmod_params(NCLS, 3, seed=w['seed'])
    if w['fn'] == 'fcn8_dae':
        return weights.synthetic_fcn8_params(NCLS, NCLS, seed=w['seed'], logit_gain=w['logit_gain'], concat=(d['concat_h'][0], 512))
    nb_h = 3 if d['concat_h'][-1] == 'input' else w.get('nb_h', 512)
    pd = weights.synthetic_dae_params(NCLS, nb_h, seed=w['seed'], out_gain=w['out_gain'], concat_h=tuple(d['concat_h']),
                                      additional_pool=d['additional_pool'], unpool_type=d['unpool_type'],
                                      conv_before_pool=d['conv_before_pool'], n_filters=d['n_filters'])
    if d['bn']:
        n_levels = (int(d['concat_h'][-1][-1]) if 'pool' in d['concat_h'][-1] else 0) + d['additional_pool']
        pd = weights.with_batchnorm(pd, n_levels, d['unpool_type'], seed=w['bn_seed'])
    return pd


DENSENET_HARDCODED_PATH = '/data/lisatmp4/romerosa/itinf/models/camvid/DenseNet103/weights/FC-DenseNet103_weights.npz'          # models/FCDenseNet.py:198


def case_densenet_params(case):
    """Synthetic FC-DenseNet103 checkpoint (oracle/densenet.py recipe) with NON-trivial BatchNorm arrays, so that a wrong
    position of beta / gamma / mean / inv_std in the positional checkpoint cannot go unnoticed."""
    import torch
    from oracle import densenet as OD
    w = case['densenet']
    params = OD.synthetic_densenet_params(3, NCLS, seed=w['seed'], logit_gain=w['logit_gain'])
    gen = torch.Generator().manual_seed(w['bn_seed'])
    k = 0
    for name, kind, ws in OD.densenet_param_shapes(3, NCLS):
        if kind == 'bnconv':
            c = params[k].shape[0]
            params[k] = 0.2 * torch.randn(c, generator=gen)                    # beta
            params[k + 1] = 0.75 + 0.5 * torch.rand(c, generator=gen)          # gamma
            params[k + 2] = torch.randn(c, generator=gen)                      # mean    (stored averages: never read,
            params[k + 3] = 0.5 + torch.rand(c, generator=gen)                 # inv_std  batch_norm_use_averages=False)
            k += 6
        else:
            k += 2
    return params


def case_batch(case, i, which='test'):
    """Batch i of a case: (X, L one-hot with the void channel) as numpy float32 (oracle/weights.py:synthetic_batch)."""
    from oracle import weights
    X, L, _ = weights.synthetic_batch(case['B'], case['H'], case['W'], NCLS, seed={'test': 100, 'train': 100, 'val': 500}[which] + i)
    return X.numpy(), L.numpy()


def param_digest(i, after, before):
    """What the fixture keeps of a trained parameter array: a strided sample of up to 4096 values, and the sum / L2 norm of its
    change from the initial checkpoint (float64)."""
    flat = np.asarray(after, np.float32).reshape(-1)
    idx = np.unique(np.linspace(0, flat.size - 1, min(flat.size, 4096)).astype(np.int64))
    delta = flat.astype(np.float64) - np.asarray(before, np.float64).reshape(-1)
    return {'p%d_sample' % i: flat[idx], 'p%d_delta' % i: np.array([delta.sum(), np.sqrt((delta ** 2).sum()), np.abs(delta).max()])}


def noise_log(case, get_output_calls, draws):
    """Which logged draw fed which DePool2D in which function call.  The graph of pred_dae (and of de = y - pred_dae) is the
    top-level get_output(dae, deterministic=True) call: its main GaussianNoiseLayer is the identity, so the random nodes it
    created are exactly the DePool2D sub-graphs', in get_all_layers order up_P .. up_1.  -> noise_k [calls, levels]: entry
    [c, p - 1] = index k of the draw that level p's mask pass used in the c-th pred_dae_fn / de_fn call."""
    total = (int(case['dae']['concat_h'][-1][-1]) if 'pool' in case['dae']['concat_h'][-1] else 0) + case['dae']['additional_pool']
    cands = [g['created'] for g in get_output_calls if g['kwargs'] == {'deterministic': True} and g['created'][1] - g['created'][0] == total]
    assert len(cands) == 1, cands
    a, b = cands[0]
    rows = {}
    for d in draws:
        if a <= d['created'] < b:
            level = total - (d['created'] - a)                     # created in the order up_P .. up_1
            rows.setdefault(d['call'], {})[level] = (d['k'], d['shape'])
    calls = sorted(rows)
    assert all(sorted(rows[c]) == list(range(1, total + 1)) for c in calls)
    return {'noise_k': np.array([[rows[c][l][0] for l in range(1, total + 1)] for c in calls], np.int64),
            'noise_batch': np.array([rows[c][1][1][0] for c in calls], np.int64)}


def train_noise_log(case, get_output_calls, draws):
    """train_dae.py with noise > 0.  The training graph is get_output(dae_lays, batch_norm_use_averages=False): it creates the
    main GaussianNoiseLayer node (first: the layer sits at the bottom of the net) and one node per DePool2D (up_P .. up_1); the
    validation graph get_output(dae_lays, deterministic=True, batch_norm_use_averages=False) only the DePool2D ones.
    -> train_k [train_fn calls, 1 + P] (main, level 1 .. P) and val_k [val_fn calls, P] of draw indices."""
    total = (int(case['dae']['concat_h'][-1][-1]) if 'pool' in case['dae']['concat_h'][-1] else 0) + case['dae']['additional_pool']
    tr = [g['created'] for g in get_output_calls if g['kwargs'] == {'batch_norm_use_averages': False} and g['created'][1] - g['created'][0] == total + 1]
    va = [g['created'] for g in get_output_calls if g['kwargs'] == {'deterministic': True, 'batch_norm_use_averages': False}
          and g['created'][1] - g['created'][0] == total]
    assert len(tr) == 1 and len(va) == 1, (tr, va)
    rows_t, rows_v = {}, {}
    for d in draws:
        if tr[0][0] <= d['created'] < tr[0][1]:
            j = d['created'] - tr[0][0]                               # 0 = main, then up_P .. up_1
            rows_t.setdefault(d['call'], {})[0 if j == 0 else total + 1 - j] = d['k']
        elif va[0][0] <= d['created'] < va[0][1]:
            rows_v.setdefault(d['call'], {})[total - (d['created'] - va[0][0])] = d['k']
    ct, cv = sorted(rows_t), sorted(rows_v)
    assert all(sorted(rows_t[c]) == list(range(total + 1)) for c in ct) and all(sorted(rows_v[c]) == list(range(1, total + 1)) for c in cv)
    return {'train_k': np.array([[rows_t[c][j] for j in range(total + 1)] for c in ct], np.int64),
            'val_k': np.array([[rows_v[c][l] for l in range(1, total + 1)] for c in cv], np.int64),
            'call_order': np.array([0 if c in rows_t else 1 for c in sorted(ct + cv)], np.int64)}


class SyntheticCamvidIterator(object):
    """The attributes and the `next()` protocol iterative_inference.py:117-125,249 use from a dataset_loaders iterator."""

    def __init__(self, case, which='test'):
        self.case, self.which = case, which
        self.cmap = np.zeros((NCLS + 1, 3), np.float32)
        self.nbatches = case['val_nbatches'] if which == 'val' else case['nbatches']
        self.non_void_nclasses = NCLS
        self.void_labels = [NCLS]
        self.data_shape = (3, case['H'], case['W'])
        self.mask_labels = ['class%d' % i for i in range(NCLS)] + ['void']
        self.i = 0

    def next(self):
        X, L = case_batch(self.case, self.i % self.nbatches, self.which)
        self.i += 1
        return X, L

    __next__ = next


def install_environment():
    """Stand-ins for what the reference imports besides its own modules."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.refrun import py2import
    py2import.install()
    getpass.getuser = lambda: 'romerosa'          # iterative_inference.py:31-51 raises for unknown users

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    current = {}
    def load_data(dataset, *a, **kw):
        if kw.get('which_set', 'all') == 'all':          # train_dae.py:127-133 unpacks (train, val, test)
            return [SyntheticCamvidIterator(current['case'], 'train'), SyntheticCamvidIterator(current['case'], 'val'), None]
        return SyntheticCamvidIterator(current['case'])
    module('data_loader', load_data=load_data)
    module('distutils', dir_util=module('distutils.dir_util', copy_tree=lambda *a, **kw: None))      # removed in Python 3.12
    module('skimage', color=module('skimage.color', rgb2gray=None, gray2rgb=None), img_as_float=None)
    module('seaborn')
    return current


def run_case(name, case, current, write=True):
    from oracle import weights
    current['case'] = case
    shutil.rmtree(WORK, ignore_errors=True)
    wdir = os.path.join(WORK, 'weights', 'camvid')
    os.makedirs(wdir)
    weights.save_npz(os.path.join(wdir, 'fcn8_model.npz'), weights.synthetic_fcn8_params(3, NCLS, **FCN8_WEIGHTS))
    out = {'case': np.array(json.dumps(case))}
    buf = io.StringIO()
    if case['script'] == 'fcn8_only':
        # models/fcn8.py:193-198 (temperature): the builder itself, called as iterative_inference.py:135-139 calls it
        import lasagne
        import theano
        import theano.tensor as T
        from models.fcn8 import buildFCN8
        x = T.tensor4('input_x_var')
        with contextlib.redirect_stdout(buf):
            fcn = buildFCN8(3, input_var=x, n_classes=NCLS, void_labels=[NCLS], path_weights=os.path.join(wdir, 'fcn8_model.npz'),
                            trainable=False, load_weights=True, layer=['pool4', 'probs_dimshuffle'], temperature=case['temperature'])
            fn = theano.function([x], lasagne.layers.get_output(fcn, deterministic=True, batch_norm_use_averages=False))
            h, y = fn(case_batch(case, 0)[0])
        out.update(pool4=h, Y_fcn=y)
    elif case['script'] == 'train':
        import helpers
        import train_dae
        train_dae.WEIGHTS_PATH = os.path.join(WORK, 'weights') + os.sep
        d = dict(case['dae'], concat_h=list(case['dae']['concat_h']))
        with contextlib.redirect_stdout(io.StringIO()):
            exp_name = helpers.build_experiment_name('fcn8', training_loss=case['training_loss'], data_aug=True,
                                                     learning_rate=case['learning_rate'], lr_anneal=case['lr_anneal'], weight_decay=1e-4,
                                                     optimizer=case.get('optimizer', 'rmsprop'), ae_h=case.get('ae_h', False), **d)
        ldir = os.path.join(WORK, 'load', 'camvid', exp_name)
        os.makedirs(ldir)
        weights.save_npz(os.path.join(ldir, 'dae_model_best.npz'), case_dae_params(case))          # resume=True reads it (train_dae.py:186)
        import lasagne.layers as LL
        from theano.sandbox import rng_mrg
        rng_mrg.STATE['evaluated'] = 0
        n_go, n_draws = len(LL.GET_OUTPUT_LOG), len(rng_mrg.STATE['log'])
        with contextlib.redirect_stdout(buf):
            train_dae.train('camvid', 'fcn8', learning_rate=case['learning_rate'], lr_anneal=case['lr_anneal'], weight_decay=1e-4,
                            num_epochs=case['num_epochs'], max_patience=100, optimizer=case.get('optimizer', 'rmsprop'), training_loss=list(case['training_loss']),
                            batch_size=[case['B']] * 3, ae_h=case.get('ae_h', False), dae_dict_updates=dict(case['dae'], concat_h=list(case['dae']['concat_h'])),
                            data_augmentation={'crop_size': None}, savepath=os.path.join(WORK, 'save'), loadpath=os.path.join(WORK, 'load'),
                            resume=True, lmb=case['lmb'])
        sdir = os.path.join(WORK, 'save', 'camvid', exp_name)
        saved = sorted(f for f in os.listdir(sdir) if f.startswith('dae_model_'))
        assert len(saved) == 1, saved          # epoch 1 writes dae_model_best.npz (improved) or dae_model_last.npz
        out['saved_as'] = np.array(saved[0])
        init = [p.numpy() for p in case_dae_params(case)]
        with np.load(os.path.join(sdir, saved[0])) as f:          # 55 M parameters: store a digest, not the arrays
            assert len(f.files) == len(init)
            for i in range(len(f.files)):
                out.update(param_digest(i, f['arr_%d' % i], init[i]))
        with np.load(os.path.join(sdir, saved[0].replace('model', 'errors'))) as f:
            out['err_train'], out['err_valid'], out['jacc_val'], out['mse_val'] = [np.asarray(f['arr_%d' % i]) for i in range(4)]
        out['output_log'] = np.array(open(os.path.join(sdir, 'output.log')).read())
        if case['dae']['noise'] > 0:
            out.update(train_noise_log(case, LL.GET_OUTPUT_LOG[n_go:], rng_mrg.STATE['log'][n_draws:]))
    else:
        import helpers
        mod = __import__('iterative_inference' if case['script'] == 'inference' else 'iterative_inference_valid')
        mod.WEIGHTS_PATH = os.path.join(WORK, 'weights') + os.sep
        d = dict(case['dae'], concat_h=list(case['dae']['concat_h']))
        with contextlib.redirect_stdout(io.StringIO()):
            exp_name = helpers.build_experiment_name('fcn8', data_aug=False, ae_h=False, **dict(list(d.items()) + list(TRAINING_DICT.items())))
        ldir = os.path.join(WORK, 'load', 'camvid', exp_name)
        os.makedirs(ldir)
        weights.save_npz(os.path.join(ldir, 'dae_model_best.npz'), case_dae_params(case))
        segm_net = case.get('segm_net', 'fcn8')
        if segm_net == 'densenet':          # build_fcdensenet restores from a hard-coded absolute path: serve it from the work directory
            weights.save_npz(os.path.join(wdir, 'densenet.npz'), case_densenet_params(case))
            real_load = np.load
            np.load = lambda path, *a, **kw: real_load(os.path.join(wdir, 'densenet.npz') if path == DENSENET_HARDCODED_PATH else path, *a, **kw)
            os.rename(ldir, ldir.replace(exp_name, exp_name.replace('fcn8', 'densenet', 1)))
            exp_name = exp_name.replace('fcn8', 'densenet', 1)
        import lasagne.layers as LL
        from theano.sandbox import rng_mrg
        rng_mrg.STATE['evaluated'] = 0          # draw k of THIS case (the fixture does not depend on what ran before it)
        n_go, n_draws = len(LL.GET_OUTPUT_LOG), len(rng_mrg.STATE['log'])
        with contextlib.redirect_stdout(buf):
            res = mod.inference('camvid', segm_net, learn_step=case['step'], num_iter=case['num_iter'],
                                dae_dict_updates=dict(case['dae'], concat_h=list(case['dae']['concat_h'])), training_dict=dict(TRAINING_DICT),
                                data_augmentation=False, which_set='test', ae_h=False, savepath=os.path.join(WORK, 'save'),
                                loadpath=os.path.join(WORK, 'load'))
        if case['dae']['noise'] > 0:
            out.update(noise_log(case, LL.GET_OUTPUT_LOG[n_go:], rng_mrg.STATE['log'][n_draws:]))
        if segm_net == 'densenet':
            np.load = real_load
        if case['script'] == 'inference':
            sdir = os.path.join(WORK, 'save', 'camvid', exp_name, 'img_plots')
            for i in range(case['nbatches']):
                with np.load(os.path.join(sdir, 'testbatch%d.npz' % i)) as f:      # `savepath+'batch'+str(i)`: no separator
                    X, L = case_batch(case, i)
                    assert np.array_equal(f['X'], X) and np.array_equal(f['L'], L)
                    if 'sub' in case:
                        k = case['sub']
                        out['Y_ii_%d' % i], out['Y_fcn_%d' % i] = f['Y_ii'][:, :, ::k, ::k], f['Y_fcn'][:, :, ::k, ::k]
                        out['labels_ii_%d' % i] = f['Y_ii'].argmax(1).astype(np.uint8)
                        out['labels_fcn_%d' % i] = f['Y_fcn'].argmax(1).astype(np.uint8)
                    else:
                        out['Y_ii_%d' % i] = f['Y_ii']
                        out['Y_fcn_%d' % i] = f['Y_fcn']
        else:
            sdir = os.path.join(WORK, 'save', 'camvid', exp_name, 'img_plots', str(case['step']), 'test')
            with np.load(os.path.join(sdir, 'iterations%s.npz' % str(case['step']))) as f:
                out['valid_mat'] = f['arr_0']
            out['res'] = np.asarray(res)
    out['stdout'] = np.array(buf.getvalue())
    if write:
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    shutil.rmtree(WORK, ignore_errors=True)
    return out


def check(names):
    """`--check NAME...`: execute the reference again and compare with the committed fixtures (nothing is written).  Arrays must
    be identical; the captured stdout may differ only in the absolute paths it prints."""
    current = install_environment()
    global WORK
    WORK = os.path.join(HERE, '_work_check')
    for name in names:
        out = run_case(name, CASES[name], current, write=False)
        with np.load(os.path.join(HERE, name + '.npz')) as f:
            assert sorted(f.files) == sorted(out), (sorted(f.files), sorted(out))
            for k in f.files:
                if k == 'stdout':
                    strip = lambda t: [ln for ln in str(t).split('\n') if '_work' not in ln]          # noqa: E731
                    assert strip(f[k]) == strip(out[k]), 'stdout of %s differs' % name
                else:
                    assert np.array_equal(f[k], out[k]), '%s: %s differs from the committed fixture' % (name, k)
        print('%s: the reference, executed again, reproduces the committed fixture exactly' % name)


if __name__ == '__main__':
    if sys.argv[1:2] == ['--check']:
        check(sys.argv[2:])
        sys.exit(0)
    current = install_environment()
    names = sys.argv[1:] or list(CASES)
    for name in names:
        out = run_case(name, CASES[name], current)
        print(name, {k: (v.shape if v.ndim else '...') for k, v in out.items()})
