#!/usr/bin/env python
"""Drop-in for the reference's iterative_inference.py (`inference`, `main`).

Same entry point and arguments (`iterative_inference.py:56-59,329-394`): dataset,
segmentation net, step, number of iterations, `dae_dict` (kind, concat_h, n_filters,
additional_pool, unpool_type, noise, temperature, ...), which_set, save/load paths.  What
changes is where the work happens: the per-image Python loop over `de_fn`/`val_fn`
(`:258-282`) becomes one device-resident CUDA-graph replay per batch
(`functions.IterativeInference`); `fused=False` runs the literal per-image loop over the four
callables instead (kept for cross-checking the fused loop).

Extra keyword arguments (not in the reference): `data_iter` (an iterator with the
`dataset_loaders` attribute contract; default: synthetic CamVid-shaped data), `fcn_params` /
`dae_params` (in-memory checkpoints instead of `.npz` paths), `verbose`.
"""
import argparse
import os

import numpy as np
import torch

from . import functions as F
from .data_loader import load_data
from .helpers import build_experiment_name, print_results, results_values
from .models.DAE_h import buildDAE
from .models.contextmod_dae import buildDAE_contextmod
from .models.fcn8_dae import buildFCN8_DAE
from .models.fcn8 import buildFCN8
from .models.FCDenseNet import build_fcdensenet

_EPSILON = 1e-3      # iterative_inference.py:53

DAE_DICT_DEFAULTS = {'kind': 'fcn8', 'dropout': 0.0, 'skip': True, 'unpool_type': 'standard', 'n_filters': 64,
                     'conv_before_pool': 1, 'additional_pool': 0, 'concat_h': ['input'], 'noise': 0.0,
                     'from_gt': True, 'temperature': 1.0, 'layer': 'probs_dimshuffle', 'exp_name': '', 'bn': 0}


def build_networks(segm_net, dae_dict, n_classes, nb_in_channels, void_labels, weights_path=None, loadpath=None,
                   dataset='camvid', fcn_params=None, dae_params=None, precision='bf16', stochastic_masks=None):
    """The network-construction block of `inference` (iterative_inference.py:127-179).
    `precision` ('bf16' | 'fp32x3' | 'mixed') selects the arithmetic of both nets (see models/DAE_h.py).
    `stochastic_masks`: the reference's non-deterministic DePool2D mask sub-graphs when dae_dict['noise'] > 0
    (layers/mylayers.py:91-93): None = as the reference (on), False = the deterministic graph, with a warning."""
    if segm_net == 'fcn8':
        fcn = buildFCN8(nb_in_channels, None, n_classes=n_classes, void_labels=void_labels,
                        path_weights=os.path.join(weights_path or '', dataset, 'fcn8_model.npz'),
                        trainable=False, load_weights=True, layer=dae_dict['concat_h'] + [dae_dict['layer']],
                        params=fcn_params, precision=precision)
        padding = 100
    elif segm_net == 'densenet':
        fcn = build_fcdensenet(None, dae_dict['concat_h'], nb_in_channels, n_classes, output_d='4d', from_gt=False,
                               weight_path=os.path.join(weights_path or '', dataset, 'DenseNet103', 'weights',
                                                        'FC-DenseNet103_weights.npz'), params=fcn_params, precision=precision)
        padding = 0
    elif segm_net == 'fcn_fcresnet':
        raise NotImplementedError
    else:
        raise ValueError
    if dae_dict['kind'] == 'standard':
        dae = buildDAE([None] * len(dae_dict['concat_h']), None, n_classes,
                       nb_features_to_concat=fcn[0].output_shape[1], padding=padding, trainable=True,
                       void_labels=void_labels, load_weights=True, path_weights=loadpath or '',
                       model_name='dae_model_best.npz', concat_h=dae_dict['concat_h'], noise=dae_dict['noise'],
                       n_filters=dae_dict['n_filters'], conv_before_pool=dae_dict['conv_before_pool'],
                       additional_pool=dae_dict['additional_pool'], dropout=dae_dict['dropout'],
                       skip=dae_dict['skip'], unpool_type=dae_dict['unpool_type'], bn=dae_dict['bn'],
                       params=dae_params, precision=precision, stochastic_masks=stochastic_masks)
    elif dae_dict['kind'] == 'contextmod':       # iterative_inference.py:171-177 -- the reference CLI's default kind
        dae = buildDAE_contextmod([None] * len(dae_dict['concat_h']), None, n_classes, path_weights=loadpath or '',
                                  model_name='dae_model_best.npz', trainable=True, load_weights=True,
                                  noise=dae_dict['noise'], concat_h=dae_dict['concat_h'], params=dae_params,
                                  nb_features_to_concat=fcn[0].output_shape[1])
    elif dae_dict['kind'] == 'fcn8':             # iterative_inference.py:165-170
        dae = buildFCN8_DAE([None] * len(dae_dict['concat_h']), None, n_classes, nb_in_channels=n_classes,
                            path_weights=loadpath or '', model_name='dae_model_best.npz', trainable=True, load_weights=True,
                            pretrained=True, pascal=False, concat_h=dae_dict['concat_h'], noise=dae_dict['noise'],
                            params=dae_params, precision=precision, nb_features_to_concat=fcn[0].output_shape[1])
    else:
        raise ValueError('Unknown dae kind')
    return fcn, dae


class _BatchStager(object):
    """Host -> device staging of (X, L) batches: two pinned buffers per tensor and a copy stream, so that the next batch
    is fetched from the iterator and copied while the current batch's loop runs on the device (the reference's iterator
    prefetches with threads, data_loader.py:58)."""

    def __init__(self, data_iter, n_batches, device, shard=None):
        self.it, self.n, self.dev = data_iter, n_batches, device
        self.stream = torch.cuda.Stream(device=device)
        self.pinned = [{}, {}]
        self.k = 0
        self.lo, self.hi = shard if shard is not None else (0, n_batches)      # this rank's batches [lo, hi)
        self.taken = 0

    def _stage(self, slot, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        buf = self.pinned[slot].get(name)
        if buf is None or tuple(buf.shape) != arr.shape:
            buf = self.pinned[slot][name] = torch.empty(arr.shape, dtype=torch.float32).pin_memory()
        buf.copy_(torch.from_numpy(arr))
        with torch.cuda.stream(self.stream):
            return buf.to(self.dev, non_blocking=True)

    def fetch(self):
        """Next (X, L, Xd, Ld, ready_event) or None; the device tensors are valid after `ready_event`."""
        while self.k < self.n and not (self.lo <= self.k < self.hi):
            self.it.next()              # another rank's batch: advance the shared iterator order past it
            self.k += 1
        if self.k >= self.n:
            return None
        slot = self.taken % 2
        self.k += 1
        self.taken += 1
        X, L = self.it.next()
        self.stream.synchronize()                 # the pinned buffers of this slot were last used two fetches ago
        Xd, Ld = self._stage(slot, 'X', X), self._stage(slot, 'L', L)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return X, L, Xd, Ld, ev


def inference(dataset, segm_net, learn_step=0.005, num_iter=500, dae_dict_updates={}, training_dict={},
              data_augmentation=False, which_set='test', ae_h=False, full_im_ft=False, savepath=None,
              loadpath=None, test_from_0_255=False, data_iter=None, fcn_params=None, dae_params=None,
              weights_path=None, fused=True, verbose=True, save_batches=False, precision='bf16', stochastic_masks=None):
    dae_dict = dict(DAE_DICT_DEFAULTS)
    dae_dict.update(dae_dict_updates)
    exp_name = build_experiment_name(segm_net, data_aug=data_augmentation, ae_h=ae_h,
                                     **dict(list(dae_dict.items()) + list(training_dict.items())))
    if savepath is None:
        raise ValueError('A saving directory must be specified')
    savepath = os.path.join(savepath, dataset, exp_name, 'img_plots', which_set)
    loadpath = os.path.join(loadpath or '', dataset, exp_name)
    os.makedirs(savepath, exist_ok=True)

    if data_iter is None:
        data_iter = load_data(dataset, {}, one_hot=True, batch_size=[10, 5, 10], return_0_255=test_from_0_255,
                              which_set=which_set)
    n_batches_test = data_iter.nbatches
    n_classes = data_iter.non_void_nclasses
    void_labels = data_iter.void_labels
    nb_in_channels = data_iter.data_shape[0]

    fcn, dae = build_networks(segm_net, dae_dict, n_classes, nb_in_channels, void_labels, weights_path, loadpath,
                              dataset, fcn_params, dae_params, precision=precision, stochastic_masks=stochastic_masks)
    pred_fcn_fn = F.function_pred_fcn(fcn)
    pred_dae_fn = F.function_pred_dae(dae)
    de_fn = F.function_de(dae)
    val_fn = F.function_val(n_classes, void_labels)
    loop = F.IterativeInference(dae, n_classes, void_labels)

    tot = {k: 0 for k in ('rec', 'acc', 'jacc', 'rec_fcn', 'acc_fcn', 'jacc_fcn', 'rec_dae', 'acc_dae', 'jacc_dae')}
    n_exec_all = []
    cm_total = np.zeros((n_classes, n_classes), np.int64)
    say = print if verbose else (lambda *a, **k: None)
    say('Inference step: ' + str(learn_step) + ' num iter ' + str(num_iter))
    # Multi-GPU (torch.distributed initialised, one process per GPU): the batches are sharded contiguously over the
    # ranks (sharding.shard_range; whole batches, so per-batch means and DenseNet's batch statistics are those of the
    # single-process run), weights are replicated and the only exchange is the final SUM all-reduce of the totals.
    from .sharding import World, shard_range
    world = World()
    b_lo, b_hi = shard_range(n_batches_test, world.rank, world.size)
    stager = _BatchStager(data_iter, n_batches_test, torch.device('cuda', torch.cuda.current_device()), (b_lo, b_hi))
    nxt = stager.fetch()
    for i in range(b_lo, b_hi):
        X, L, Xd, Ld, ready = nxt
        torch.cuda.current_stream().wait_event(ready)
        Xd.record_stream(torch.cuda.current_stream()); Ld.record_stream(torch.cuda.current_stream())
        nxt = None
        pred = pred_fcn_fn(Xd)
        Y, H = pred[-1], pred[:-1]

        acc_fcn, jacc_fcn, rec_fcn = val_fn(Y, Ld)                      # metrics before iterative inference
        tot['acc_fcn'] += acc_fcn; tot['jacc_fcn'] = tot['jacc_fcn'] + jacc_fcn; tot['rec_fcn'] += rec_fcn
        Y_dae = pred_dae_fn(*(H + [Y]))                                 # one plain DAE pass
        acc_dae, jacc_dae, rec_dae = val_fn(Y_dae, Ld)
        tot['acc_dae'] += acc_dae; tot['jacc_dae'] = tot['jacc_dae'] + jacc_dae; tot['rec_dae'] += rec_dae

        if fused:
            res = loop.run(H[0], Y, learn_step, num_iter, eps=_EPSILON, onehot=Ld)
            nxt = stager.fetch()                  # the loop is queued: fetch + copy the next batch under it
            Y_ii = res['y']
            n_exec_all += res['n_exec'].cpu().tolist()
            cm = res['cm'].sum(0).cpu().numpy().reshape(n_classes, n_classes)
            cnt = res['counts'].sum(0).cpu().numpy()
            se = res['sqerr'].sum(0).cpu().numpy()
            acc, jacc, rec = (np.float32(np.float32(cnt[0]) / np.float32(cnt[1])), F.jaccard_from_cm(cm),
                              np.float32(se[0] / se[1]))
        else:
            outs = []
            for im in range(Xd.shape[0]):                               # iterative_inference.py:258-282
                h_im = [el[im:im + 1] for el in H]
                y_im = Y[im:im + 1]
                n_exec = 0
                for it in range(num_iter):
                    grad = de_fn(*(h_im + [y_im]))
                    y_im = torch.clamp(y_im - learn_step * grad, 0.0, 1.0)
                    n_exec += 1
                    norm = float(torch.linalg.vector_norm(grad, dim=1).mean())
                    if norm < _EPSILON:
                        break
                outs.append(y_im)
                n_exec_all.append(n_exec)
            Y_ii = torch.cat(outs, dim=0)
            acc, jacc, rec = val_fn(Y_ii, Ld)
            cm = np.zeros_like(cm_total)
        if nxt is None:
            nxt = stager.fetch()
        cm_total += cm
        tot['acc'] += acc; tot['jacc'] = tot['jacc'] + jacc; tot['rec'] += rec
        if verbose:
            print_results('>>>>> FCN:', tot['rec_fcn'], tot['acc_fcn'], tot['jacc_fcn'], i - b_lo + 1)
            print_results('>>>>> FCN+DAE:', tot['rec_dae'], tot['acc_dae'], tot['jacc_dae'], i - b_lo + 1)
            print_results('>>>>> ITERATIVE INFERENCE:', tot['rec'], tot['acc'], tot['jacc'], i - b_lo + 1)
        if save_batches:                                                # iterative_inference.py:293
            np.savez(os.path.join(savepath, 'batch' + str(i) + '.npz'), X=X, L=L, Y_ii=Y_ii.cpu().numpy(),
                     Y_fcn=Y.cpu().numpy())
    if world.size > 1:          # totals over all shards: integer counts (exact in float64 / int64) and per-batch means
        dev = torch.device('cuda', torch.cuda.current_device())
        keys = sorted(tot)
        parts = [np.broadcast_to(np.asarray(tot[k], dtype=np.float64), (2, n_classes) if 'jacc' in k else ()).reshape(-1) for k in keys]
        flat = torch.tensor(np.concatenate(parts), dtype=torch.float64, device=dev)
        world.allreduce_sum(flat)
        cm_dev = torch.from_numpy(cm_total).to(dev)
        world.allreduce_sum(cm_dev)
        cm_total = cm_dev.cpu().numpy()
        flat, off = flat.cpu().numpy(), 0
        for k, part in zip(keys, parts):
            v = flat[off:off + part.size]
            tot[k] = v.reshape(2, n_classes).astype(np.float32) if 'jacc' in k else np.float32(v[0])
            off += part.size
        gathered = [None] * world.size
        torch.distributed.all_gather_object(gathered, n_exec_all)
        n_exec_all = [n for g in gathered for n in g]
    nb = n_batches_test
    return {'fcn': results_values(tot['rec_fcn'], tot['acc_fcn'], tot['jacc_fcn'], nb),
            'fcn_dae': results_values(tot['rec_dae'], tot['acc_dae'], tot['jacc_dae'], nb),
            'iterative': results_values(tot['rec'], tot['acc'], tot['jacc'], nb),
            'jacc_tot': tot['jacc'], 'jacc_tot_fcn': tot['jacc_fcn'], 'cm': cm_total, 'n_exec': n_exec_all,
            'savepath': savepath}


def _flag(v):
    """argparse `type=bool` treats every non-empty string as True (the reference's flags do); this one parses it."""
    return str(v).strip().lower() in ('1', 'true', 'yes', 'y', 'on')


def _literal_dict(v):
    import ast
    d = ast.literal_eval(v) if isinstance(v, str) else v
    if not isinstance(d, dict):
        raise argparse.ArgumentTypeError('expected a dict literal, e.g. "{\'noise\': 0}"')
    return d


def main():
    """Same flags as the reference's CLI (iterative_inference.py:329-394).  `-dae_dict` / `-training_dict` take a Python
    dict literal; the defaults are the benchmark's DAE (kind=standard, concat_h=[pool4]); the reference's default
    (kind=contextmod, concat_h=[input]) is selected with -dae_dict "{'kind': 'contextmod', 'concat_h': ['input']}"."""
    parser = argparse.ArgumentParser(description='Iterative inference.')
    parser.add_argument('-dataset', type=str, default='camvid', help='Dataset.')
    parser.add_argument('-segmentation_net', type=str, default='fcn8', help='Segmentation network: fcn8 | densenet')
    parser.add_argument('-step', type=float, default=1.0, help='step')
    parser.add_argument('--num_iter', '-ne', type=int, default=1, help='Max number of iterations')
    parser.add_argument('-which_set', type=str, default='test', help='Inference set')
    parser.add_argument('-dae_dict', type=_literal_dict,
                        default={'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0,
                                 'concat_h': ['pool4'], 'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1,
                                 'additional_pool': 2, 'path_weights': '', 'layer': 'probs_dimshuffle',
                                 'exp_name': 'flip_final_', 'bn': 0}, help='DAE kind and parameters')
    parser.add_argument('-training_dict', type=_literal_dict,
                        default={'training_loss': ['crossentropy'], 'learning_rate': 0.0001, 'lr_anneal': 0.99,
                                 'weight_decay': 0.0001, 'optimizer': 'rmsprop'}, help='Training parameters')
    parser.add_argument('-full_im_ft', type=_flag, default=False, help='Whether to finetune at full image resolution')
    parser.add_argument('-ae_h', type=_flag, default=False, help='Whether to reconstruct intermediate h')
    parser.add_argument('-data_augmentation', type=_flag, default=True, help='Data augmentation tag of the experiment name')
    parser.add_argument('-test_from_0_255', type=_flag, default=False, help='Whether images are within the 0-255 range')
    parser.add_argument('-savepath', type=str, default='./iiseg_out/')
    parser.add_argument('-loadpath', type=str, default='./iiseg_models/')
    parser.add_argument('-weights_path', type=str, default='./iiseg_models/')
    parser.add_argument('-precision', type=str, default='bf16', choices=['bf16', 'fp32x3', 'mixed'])
    parser.add_argument('-stochastic_masks', type=_flag, default=None,
                        help="DePool2D masks from a separately noised pass when dae_dict['noise'] > 0, as the reference's graph does")
    args = parser.parse_args()
    inference(args.dataset, args.segmentation_net, float(args.step), int(args.num_iter), which_set=args.which_set,
              savepath=args.savepath, loadpath=args.loadpath, weights_path=args.weights_path,
              full_im_ft=args.full_im_ft, test_from_0_255=args.test_from_0_255, ae_h=args.ae_h,
              dae_dict_updates=args.dae_dict, data_augmentation=args.data_augmentation,
              training_dict=args.training_dict, precision=args.precision, stochastic_masks=args.stochastic_masks)


if __name__ == '__main__':
    main()
