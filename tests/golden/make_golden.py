"""Generates the golden vectors under tests/golden/ from the CPU oracle.

The reference (Theano/Lasagne, Python 2) cannot run in this image, so these vectors are
outputs of the oracle restatement, not of the reference itself ("parity unpinned"); they pin
the oracle against regressions and give the GPU tests fixed inputs/outputs.  Weights are
regenerated from seeds (oracle/weights.py), only activations are stored.

    python -m tests.golden.make_golden
"""
import os

import numpy as np
import torch

from oracle import loop, metrics as M, nets, weights

HERE = os.path.dirname(os.path.abspath(__file__))
H, W, NCLS = 32, 40, 11
LOGIT_GAIN, OUT_GAIN = 10.0, 0.1


def _setup():
    X, L, lab = weights.synthetic_batch(1, H, W, NCLS, seed=0)
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=LOGIT_GAIN)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=OUT_GAIN)
    return X, L, lab, pf, pd


def fcn8_case():
    X, L, lab, pf, _ = _setup()
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    return {'X': X.numpy(), 'pool4': h.numpy(), 'probs': y0.numpy()}


def dae_case():
    X, L, lab, pf, pd = _setup()
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    p = nets.dae_forward(pd, y0, h, 100)
    return {'h': h.numpy(), 'y': y0.numpy(), 'p': p.numpy()}


def loop_case():
    X, L, lab, pf, pd = _setup()
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    Y, n_exec, bm, valid_mat = loop.inference_batch(pd, h, y0, 0.05, 4, 100, L=L.numpy(), n_classes=NCLS,
                                                    void_labels=[NCLS])
    return {'labels': lab.numpy().astype(np.int32), 'y_final': Y.numpy(), 'n_exec': np.array(n_exec, np.int32),
            'cm': M.confusion_matrix(Y.numpy(), L.numpy(), NCLS), 'acc': np.float32(bm[0]), 'jacc': bm[1],
            'mse': np.float32(bm[2]), 'valid_mat': valid_mat}


CASES = {'fcn8_32x40': fcn8_case, 'dae_32x40': dae_case, 'loop_32x40': loop_case}

if __name__ == '__main__':
    for name, fn in CASES.items():
        out = fn()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, {k: v.shape for k, v in out.items()})
