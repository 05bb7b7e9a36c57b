#!/usr/bin/env python
"""Drop-in for iterative_inference_valid.py: the step-size / iteration-count sweep.

The reference rebuilds and recompiles both networks once per step value
(`iterative_inference_valid.py:373-382`) and accumulates `valid_mat[:, :, it] += jacc_iter`
from a host `val_fn` call per image per iteration (`:265-288`).  Here the networks are built
once, `h` and `y0` of each batch are computed once and reused for all step values, and the
per-iteration confusion matrices are accumulated on the device inside the captured loop
(int64, exact); an image that early-exits at iteration `it` contributes nothing to
`valid_mat[..., it:]`, exactly like the reference's `break` before `val_fn` (`:282-288`).
"""
import numpy as np
import torch

from . import functions as F
from .data_loader import load_data
import argparse
import os

from .iterative_inference import DAE_DICT_DEFAULTS, _EPSILON, _flag, _literal_dict, build_networks

STEPS = [.01, .02, .05, .08, .1, .5, 1.]      # iterative_inference_valid.py:373


def sweep(dataset, segm_net, steps=STEPS, num_iter=50, dae_dict_updates={}, which_set='val', data_iter=None,
          fcn_params=None, dae_params=None, weights_path=None, loadpath=None, verbose=True,
          precision='bf16', savepath=None, eps=_EPSILON, stochastic_masks=None):
    """Returns (all_results[len(steps), num_iter], valid_mats[len(steps), 2, C, num_iter]).  `eps` is the convergence
    threshold of the loop (the reference's _EPSILON = 1e-3, iterative_inference_valid.py:53)."""
    dae_dict = dict(DAE_DICT_DEFAULTS)
    dae_dict.update(dae_dict_updates)
    if data_iter is None:
        data_iter = load_data(dataset, {}, one_hot=True, batch_size=[10, 10, 10], which_set=which_set)
    n_classes, void_labels = data_iter.non_void_nclasses, data_iter.void_labels
    fcn, dae = build_networks(segm_net, dae_dict, n_classes, data_iter.data_shape[0], void_labels, weights_path,
                              loadpath, dataset, fcn_params, dae_params, precision=precision, stochastic_masks=stochastic_masks)
    pred_fcn_fn = F.function_pred_fcn(fcn)
    loop = F.IterativeInference(dae, n_classes, void_labels)
    valid_mats = np.zeros((len(steps), 2, n_classes, num_iter))        # float64 like the reference (:231)
    # multi-GPU: contiguous batch shards per rank, one SUM all-reduce of valid_mats at the end (integer counts held in
    # float64: exact, order-free)
    from .sharding import World, shard_range
    world = World()
    b_lo, b_hi = shard_range(data_iter.nbatches, world.rank, world.size)
    for b in range(data_iter.nbatches):
        X, L = data_iter.next()
        if not (b_lo <= b < b_hi):
            continue
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
        Ld = torch.from_numpy(np.ascontiguousarray(L, dtype=np.float32)).cuda()
        pred = pred_fcn_fn(Xd)                                         # once per batch, shared by every step value
        Y, H = pred[-1], pred[:-1]
        for si, s in enumerate(steps):
            res = loop.run(H[0], Y, s, num_iter, eps=eps, onehot=Ld, per_iter_metrics=True)
            for it, acc in enumerate(res['iter']):
                cms = acc.cm.cpu().numpy().reshape(-1, n_classes, n_classes)
                for cm in cms:                                         # per image: val_fn(y_im, t_im) of the reference
                    if cm.sum() > 0:
                        valid_mats[si, :, :, it] += F.jaccard_from_cm(cm)
    if world.size > 1:
        vm = torch.from_numpy(valid_mats).cuda()
        world.allreduce_sum(vm)
        valid_mats = vm.cpu().numpy()
    if savepath is not None and world.rank == 0:   # the reference writes one `iterations<step>.npz` per step value (:303)
        os.makedirs(savepath, exist_ok=True)
        for si, s in enumerate(steps):
            np.savez(os.path.join(savepath, 'iterations' + str(s) + '.npz'), valid_mats[si])
    with np.errstate(divide='ignore', invalid='ignore'):
        all_results = np.nanmean(valid_mats[:, 0] / valid_mats[:, 1], axis=1)
    sweep.last_graph_captures = loop.graph_captures         # one captured graph serves every step value
    if verbose:
        best = all_results.max(1)
        print('Best step: ' + str(steps[int(best.argmax())]))
        print('Result: ' + str(best.max()))
        print('Num iters: ' + str(int(all_results.argmax(1)[int(best.argmax())]) + 1))
    return all_results, valid_mats


def inference(dataset, segm_net, learn_step=0.005, num_iter=500, dae_dict_updates={}, training_dict={},
              data_augmentation=False, which_set='test', ae_h=False, full_im_ft=False, savepath=None, loadpath=None,
              test_from_0_255=False, **kw):
    """Single-step form with the reference's signature and directory layout (iterative_inference_valid.py:56-93): the DAE
    checkpoint is read from `<loadpath>/<dataset>/<exp_name>/dae_model_best.npz`, `iterations<step>.npz` (valid_mat, :303) is
    written to `<savepath>/<dataset>/<exp_name>/img_plots/<step>/<which_set>/`; returns
    `res = nanmean(valid_mat[0] / valid_mat[1], axis=0)` (`:297`)."""
    from .helpers import build_experiment_name
    dae_dict = dict(DAE_DICT_DEFAULTS)
    dae_dict.update(dae_dict_updates)
    exp_name = build_experiment_name(segm_net, data_aug=data_augmentation, ae_h=ae_h,
                                     **dict(list(dae_dict.items()) + list(training_dict.items())))
    exp_name += '_ftsmall' if full_im_ft else ''
    if savepath is None:
        raise ValueError('A saving directory must be specified')
    savepath = os.path.join(savepath, dataset, exp_name, 'img_plots', str(learn_step), which_set)
    loadpath = os.path.join(loadpath or '', dataset, exp_name)
    res, _ = sweep(dataset, segm_net, [learn_step], num_iter, dae_dict_updates, which_set, loadpath=loadpath,
                   savepath=savepath, **kw)
    return res[0]


def main():
    """The reference's CLI (iterative_inference_valid.py:312-394): same flags and defaults; the seven step values are swept in
    ONE job (networks built once, h / y0 of a batch shared by all steps) instead of seven rebuilds."""
    parser = argparse.ArgumentParser(description='Iterative inference.')
    parser.add_argument('-dataset', type=str, default='camvid', help='Dataset.')
    parser.add_argument('-segmentation_net', type=str, default='fcn8', help='Segmentation network.')
    parser.add_argument('-step', type=float, default=0.05, help='step (ignored: the sweep covers STEPS, like the reference)')
    parser.add_argument('--num_iter', '-ne', type=int, default=50, help='Max number of iterations')
    parser.add_argument('-which_set', type=str, default='val', help='Inference set')
    parser.add_argument('-dae_dict', type=_literal_dict,
                        default={'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0.5,
                                 'concat_h': ['pool4'], 'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1,
                                 'additional_pool': 2, 'path_weights': '', 'layer': 'probs_dimshuffle',
                                 'exp_name': 'flip_final_', 'bn': 0}, help='DAE kind and parameters')
    parser.add_argument('-training_dict', type=_literal_dict,
                        default={'training_loss': ['crossentropy', 'squared_error'], 'learning_rate': 0.001, 'lr_anneal': 0.99,
                                 'weight_decay': 0.0001, 'optimizer': 'rmsprop'}, help='Training parameters')
    parser.add_argument('-full_im_ft', type=_flag, default=False)
    parser.add_argument('-ae_h', type=_flag, default=False)
    parser.add_argument('-data_augmentation', type=_flag, default=True)
    parser.add_argument('-test_from_0_255', type=_flag, default=False)
    parser.add_argument('-savepath', type=str, default='./iiseg_out/')
    parser.add_argument('-loadpath', type=str, default='./iiseg_models/')
    parser.add_argument('-weights_path', type=str, default='./iiseg_models/')
    args = parser.parse_args()
    sweep(args.dataset, args.segmentation_net, STEPS, int(args.num_iter), args.dae_dict, args.which_set,
          weights_path=args.weights_path, loadpath=args.loadpath, savepath=args.savepath)


if __name__ == '__main__':
    main()
