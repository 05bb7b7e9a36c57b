"""N > 1 on real GPUs (NCCL): the packaged drop-ins `inference()` / `sweep()` shard their batches over the ranks and
all-reduce the totals; the data-parallel train step all-reduces its gradients in buckets under backward.  Needs two
GPUs on the box (skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).  The CPU-side
protocol (shard ranges, metric reduction, loss denominators) is covered with gloo in tests/test_host_logic.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(out_dir, n):
    cmd = [sys.executable]
    if n > 1:
        cmd += ['-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
                '--master-port', '29517']
    cmd += [os.path.join(ROOT, 'tests', 'dist_worker.py'), out_dir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return np.load(os.path.join(out_dir, 'result_%d.npz' % n))


def test_two_ranks_equal_one_rank(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    one = _launch(str(tmp_path), 1)
    two = _launch(str(tmp_path), 2)
    # inference(): integer totals are order-free -> identical on 1 and 2 ranks; per-batch means only differ by float summation order
    assert np.array_equal(one['cm'], two['cm']) and np.array_equal(one['jacc'], two['jacc']) and np.array_equal(one['jacc_fcn'], two['jacc_fcn'])
    assert sorted(one['n_exec'].tolist()) == sorted(two['n_exec'].tolist()) and len(two['n_exec']) == 7
    assert np.allclose(one['it'], two['it'], rtol=1e-6, atol=1e-7)
    # sweep(): valid_mat sums of integer counts
    assert np.array_equal(one['mats'], two['mats']) and np.array_equal(one['res'], two['res'], equal_nan=True)
    # train step: bucketed / overlapped all-reduce == blocking all-reduce bit for bit; 2 ranks == 1 rank up to summation order
    assert bool(two['same'][0]) and bool(one['same'][0])
    assert abs(one['loss'][0] - two['loss'][0]) < 1e-6 * abs(one['loss'][0])
    for k in ('w0', 'w_last'):            # the weight UPDATE of two ranks vs one, relative L2 (bf16 gradients, different summation order)
        upd = np.linalg.norm(one[k] - one[k + '_init'])
        assert np.linalg.norm(one[k] - two[k]) < 0.05 * upd, (k, np.linalg.norm(one[k] - two[k]), upd)
    # adam + crossentropy / dice / squared_error: the dice term's whole-batch sums (I, T, P) are global -> the same loss on 1 and 2 ranks
    assert abs(one['loss_adam_dice'][0] - two['loss_adam_dice'][0]) < 1e-6 * abs(one['loss_adam_dice'][0])
    assert abs(one['loss_adam_dice'][0] - one['loss'][0]) > 1e-4          # the dice term is in the loss
    upd = np.linalg.norm(one['w_last_adam_dice'] - one['w_last_init'])
    d = np.linalg.norm(one['w_last_adam_dice'] - two['w_last_adam_dice'])
    print('adam + dice, 2 ranks vs 1: weight update differs by %.3f of its norm' % (d / upd))
    assert d < 0.2 * upd          # adam's first step is lr * sign(g): only elements with ~0 gradient may differ (summation order)
