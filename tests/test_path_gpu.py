"""End-to-end GPU parity of the hot path against the CPU oracle on the same seeded inputs and
weights: FCN8 forward, one DAE application (teacher-forced), the N-step loop with metrics.

Arithmetic: bf16 operands, fp32 accumulation (dtype "bf16").  Tolerances are the bf16 variant's
own: the pool mask is a discontinuous function of the conv outputs, so bf16
rounding flips some tie decisions; bounds below are max-abs on probabilities and argmax
agreement, measured with tests/parity_report.py and given ~2x headroom.  The integer reductions
(confusion matrix, counts) are bit-exact GIVEN the labels, which is asserted separately by
feeding the oracle the device's own y."""
import os

import numpy as np
import pytest
import torch

from oracle import loop, metrics as M, nets, weights

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NCLS = 11
TOL_FCN_PROBS = 3e-2      # y0 max-abs, bf16 FCN8 (logit gain 10 amplifies bf16 rounding)
TOL_DAE_P = 1e-2          # one DAE application, max-abs on p
TOL_LOOP_Y = 1e-2         # y after the loop, max-abs
MIN_ARGMAX = 0.99         # argmax agreement


def _nets(cuda, precision='bf16'):
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, precision=precision)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=fcn[0].output_shape[1], padding=100,
                   concat_h=['pool4'], noise=0.0, n_filters=64, conv_before_pool=1, additional_pool=2,
                   skip=True, unpool_type='trackind', params=pd, precision=precision)
    return pf, pd, fcn, dae


@pytest.fixture(scope='module')
def built(cuda):
    return _nets(cuda)


def test_fcn8_forward_vs_golden(cuda, built):
    from iterative_inference_segm_b200.functions import function_pred_fcn
    pf, pd, fcn, dae = built
    g = np.load(os.path.join(GOLD, 'fcn8_32x40.npz'))
    h, y0 = function_pred_fcn(fcn)(g['X'])
    assert h.shape == g['pool4'].shape and y0.shape == g['probs'].shape and h.dtype == np.float32
    scale = float(np.abs(g['pool4']).max())
    assert float(np.abs(h - g['pool4']).max()) < 3e-2 * scale
    assert float(np.abs(y0 - g['probs']).max()) < TOL_FCN_PROBS
    assert float((y0.argmax(1) == g['probs'].argmax(1)).mean()) >= MIN_ARGMAX


def test_dae_application_vs_golden(cuda, built):
    """Teacher-forced: the oracle's own (h, y) in, compare p = DAE(y, h)."""
    from iterative_inference_segm_b200.functions import function_pred_dae, function_de
    pf, pd, fcn, dae = built
    g = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    p = function_pred_dae(dae)(g['h'], g['y'])
    assert p.shape == g['p'].shape
    assert float(np.abs(p - g['p']).max()) < TOL_DAE_P
    assert np.allclose(p.sum(1), 1.0, atol=1e-5)
    de = function_de(dae)(g['h'], g['y'])
    assert float(np.abs(de - (g['y'] - p)).max()) < 1e-6


def test_loop_vs_oracle_and_exact_metrics(cuda, built):
    from iterative_inference_segm_b200.functions import IterativeInference, jaccard_from_cm
    pf, pd, fcn, dae = built
    gd = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    gl = np.load(os.path.join(GOLD, 'loop_32x40.npz'))
    ii = IterativeInference(dae, NCLS, [NCLS])
    h = torch.from_numpy(gd['h']).to(cuda)
    y0 = torch.from_numpy(gd['y']).to(cuda)
    labels = torch.from_numpy(gl['labels']).to(cuda)
    for use_graph in (False, True):
        res = ii.run(h, y0, 0.05, 4, labels=labels, per_iter_metrics=True, use_graph=use_graph)
        torch.cuda.synchronize()
        y = res['y'].cpu().numpy()
        assert res['n_exec'].cpu().tolist() == gl['n_exec'].tolist()
        assert float(np.abs(y - gl['y_final']).max()) < TOL_LOOP_Y
        assert float((y.argmax(1) == gl['y_final'].argmax(1)).mean()) >= MIN_ARGMAX
        # integer reductions: bit-exact given identical labels -> oracle metrics of the DEVICE's y
        onehot = np.eye(NCLS + 1, dtype=np.float32)[gl['labels']].transpose(0, 3, 1, 2)
        cm = res['cm'].sum(0).cpu().numpy().reshape(NCLS, NCLS)
        assert np.array_equal(cm, M.confusion_matrix(y, onehot, NCLS))
        assert np.array_equal(jaccard_from_cm(cm), M.jaccard(y, onehot, NCLS))
        c, v = M.accuracy_counts(y, onehot, [NCLS])
        assert res['counts'].sum(0).cpu().tolist() == [c, v]
        assert len(res['iter']) == 4
    # graph replay is deterministic
    y1 = ii.run(h, y0, 0.05, 4, labels=labels, per_iter_metrics=True)['y'].clone()
    y2 = ii.run(h, y0, 0.05, 4, labels=labels, per_iter_metrics=True)['y'].clone()
    assert torch.equal(y1, y2)


def test_windowed_down_path_is_exact(cuda, built):
    """Recomputing only the y-dependent windows of the contracting path (iterations 2..N) gives
    bit-identical logits to recomputing everything: the hoisted borders are iteration-invariant."""
    pf, pd, fcn, dae = built
    gd = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    from iterative_inference_segm_b200 import _kernels as K
    net = dae.net
    h = K.pack_nchw(torch.from_numpy(gd['h']).to(cuda), net.h_pad)
    y_a = K.pack_nchw(torch.from_numpy(gd['y']).to(cuda), net.y_cpad)
    y_b = K.pack_nchw(torch.softmax(torch.randn(1, NCLS, 32, 40, device=cuda), 1), net.y_cpad)
    net.logits(h, y_a, full_down=True)                       # fills the borders for this h
    win = net.logits(h, y_b, full_down=False).clone()        # different y, windows only
    full = net.logits(h, y_b, full_down=True).clone()
    assert torch.equal(win, full)


def test_early_exit_semantics(cuda, built):
    """eps huge: every image does exactly one update, then is frozen; no per-iteration metrics."""
    from iterative_inference_segm_b200.functions import IterativeInference
    pf, pd, fcn, dae = built
    gd = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    gl = np.load(os.path.join(GOLD, 'loop_32x40.npz'))
    ii = IterativeInference(dae, NCLS, [NCLS])
    h = torch.from_numpy(gd['h']).to(cuda)
    y0 = torch.from_numpy(gd['y']).to(cuda)
    labels = torch.from_numpy(gl['labels']).to(cuda)
    one = ii.run(h, y0, 0.05, 1, eps=0.0, labels=labels, use_graph=False)['y'].clone()
    res = ii.run(h, y0, 0.05, 3, eps=1e9, labels=labels, per_iter_metrics=True, use_graph=False)
    assert res['n_exec'].cpu().tolist() == [1]
    assert torch.equal(res['y'], one)
    assert all(int(a.cm.sum()) == 0 for a in res['iter'])
    assert int(res['cm'].sum()) > 0          # the final batch-level val_fn still runs on the frozen y


def _tiny_iter(n_images=3, batch=2):
    from iterative_inference_segm_b200.data_loader import SyntheticSegmentationIterator
    return SyntheticSegmentationIterator(n_images, batch, 32, 40, NCLS, seed=5)


DAE_DICT = {'kind': 'standard', 'dropout': 0, 'skip': True, 'unpool_type': 'trackind', 'noise': 0, 'concat_h': ['pool4'],
            'from_gt': False, 'n_filters': 64, 'conv_before_pool': 1, 'additional_pool': 2, 'path_weights': '',
            'layer': 'probs_dimshuffle', 'exp_name': 'flip_final_', 'bn': 0}


def test_inference_script_dropin(cuda, tmp_path):
    """inference(...) with the reference's arguments: fused CUDA-graph loop == literal per-image loop over
    the four callables, and both track the oracle's inference_batch."""
    from iterative_inference_segm_b200.iterative_inference import inference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    kw = dict(dae_dict_updates=DAE_DICT, savepath=str(tmp_path), loadpath=str(tmp_path), fcn_params=pf, dae_params=pd,
              verbose=False)
    a = inference('camvid', 'fcn8', 0.05, 3, data_iter=_tiny_iter(), fused=True, save_batches=True, **kw)
    b = inference('camvid', 'fcn8', 0.05, 3, data_iter=_tiny_iter(), fused=False, **kw)
    assert a['n_exec'] == b['n_exec'] == [3, 3, 3]
    assert np.array_equal(a['jacc_tot'], b['jacc_tot'])             # same kernels, same labels -> identical counts
    assert abs(a['iterative'][0] - b['iterative'][0]) < 1e-6 and a['iterative'][1] == b['iterative'][1]
    assert os.path.exists(os.path.join(a['savepath'], 'batch0.npz'))
    # oracle: same data, same weights
    it = _tiny_iter()
    jacc_o = 0
    agree, total = 0, 0
    for i in range(it.nbatches):
        X, L = it.next()
        h, y0 = nets.fcn8_forward(pf, torch.from_numpy(X), NCLS)
        Y, n_exec, bm, _ = loop.inference_batch(pd, h, y0, 0.05, 3, 100, L=L, n_classes=NCLS, void_labels=[NCLS])
        jacc_o = jacc_o + bm[1]
        saved = np.load(os.path.join(a['savepath'], 'batch%d.npz' % i))
        agree += (saved['Y_ii'].argmax(1) == Y.numpy().argmax(1)).sum()
        total += Y.numpy().argmax(1).size
        assert float(np.abs(saved['Y_ii'] - Y.numpy()).max()) < 5e-2    # bf16 FCN8 (logit gain 10) + bf16 DAE
    assert agree / total >= 0.98
    # confusion-matrix totals: same pixel count, per-class counts within the argmax disagreement
    assert a['jacc_tot'][1].sum() >= jacc_o[1].sum() * 0.9


def test_valid_sweep_per_iteration_matrices(cuda):
    """iterative_inference_valid, internal consistency: the per-iteration Jaccard accumulators of the captured loop equal
    val_fn on the per-iteration y (the comparison with the ORACLE's valid_mat is test_valid_sweep_matches_oracle_valid_mat)."""
    from iterative_inference_segm_b200.iterative_inference_valid import sweep
    from iterative_inference_segm_b200.functions import IterativeInference, function_pred_fcn, function_val
    from iterative_inference_segm_b200.iterative_inference import build_networks, DAE_DICT_DEFAULTS
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    res, mats = sweep('camvid', 'fcn8', steps=[0.05, 0.5], num_iter=2, dae_dict_updates=DAE_DICT,
                      data_iter=_tiny_iter(2, 2), fcn_params=pf, dae_params=pd, verbose=False)
    assert res.shape == (2, 2) and mats.shape == (2, 2, NCLS, 2)
    # recompute iteration 1 of step 0.05 by hand: one update, then val_fn per image
    dd = dict(DAE_DICT_DEFAULTS); dd.update(DAE_DICT)
    fcn, dae = build_networks('fcn8', dd, NCLS, 3, [NCLS], fcn_params=pf, dae_params=pd)
    X, L = _tiny_iter(2, 2).next()
    Xd, Ld = torch.from_numpy(X).cuda(), torch.from_numpy(L).cuda()
    h, y0 = function_pred_fcn(fcn)(Xd)
    y1 = IterativeInference(dae, NCLS, [NCLS]).run(h, y0, 0.05, 1, onehot=Ld, use_graph=False)['y']
    val_fn = function_val(NCLS, [NCLS])
    expect = sum(val_fn(y1[i:i + 1], Ld[i:i + 1])[1] for i in range(2))
    assert np.array_equal(mats[0, :, :, 0], expect.astype(np.float64))


def test_metrics_module_matches_oracle(cuda):
    from iterative_inference_segm_b200 import metrics as GM
    rng = np.random.RandomState(4)
    y = rng.rand(2, NCLS, 9, 7).astype(np.float32)
    lab = rng.randint(0, NCLS + 1, size=(2, 9, 7))
    t = np.eye(NCLS + 1, dtype=np.float32)[lab].transpose(0, 3, 1, 2).copy()
    y2d = y.transpose(0, 2, 3, 1).reshape(-1, NCLS)
    t2d = t.transpose(0, 2, 3, 1).reshape(-1, NCLS + 1)
    assert np.array_equal(GM.jaccard(y2d, t2d, NCLS, one_hot=True), M.jaccard(y, t, NCLS))
    assert GM.accuracy(y2d, t2d, [NCLS], one_hot=True) == M.accuracy(y, t, [NCLS])
    assert abs(float(GM.squared_error(y, t, NCLS)) - float(M.squared_error(y, t, NCLS))) < 1e-6


# ---------------------------------------------------------------------------
# fp32-accurate variant (precision='fp32x3'): BASELINE.json's fp32/TF32 bar -- per-iteration
# probabilities within 2e-3 max-abs, argmax agreement >= 99.9 % -- against the same oracle vectors.
# ---------------------------------------------------------------------------
TOL_F32 = 2e-3
MIN_ARGMAX_F32 = 0.999


@pytest.fixture(scope='module', params=['fp32x3', 'mixed'])
def built_f32(cuda, request):
    """Both fp32-grade variants are held to the same bar: 'fp32x3' (every conv fp32-accurate) and 'mixed'
    (fp32-accurate FCN8 + contracting path, bf16 expanding path with the fused softmax/update epilogue)."""
    return _nets(cuda, request.param)


def test_fp32x3_fcn8_forward_vs_golden(cuda, built_f32):
    from iterative_inference_segm_b200.functions import function_pred_fcn
    pf, pd, fcn, dae = built_f32
    g = np.load(os.path.join(GOLD, 'fcn8_32x40.npz'))
    h, y0 = function_pred_fcn(fcn)(g['X'])
    scale = float(np.abs(g['pool4']).max())
    assert float(np.abs(h - g['pool4']).max()) < 1e-4 * scale
    assert float(np.abs(y0 - g['probs']).max()) < TOL_F32
    assert float((y0.argmax(1) == g['probs'].argmax(1)).mean()) >= MIN_ARGMAX_F32


def test_fp32x3_dae_and_loop_vs_golden(cuda, built_f32):
    from iterative_inference_segm_b200.functions import function_pred_dae, IterativeInference
    pf, pd, fcn, dae = built_f32
    g = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    gl = np.load(os.path.join(GOLD, 'loop_32x40.npz'))
    p = function_pred_dae(dae)(g['h'], g['y'])
    assert float(np.abs(p - g['p']).max()) < TOL_F32
    ii = IterativeInference(dae, NCLS, [NCLS])
    for use_graph in (False, True):
        res = ii.run(torch.from_numpy(g['h']).to(cuda), torch.from_numpy(g['y']).to(cuda), 0.05, 4,
                     labels=torch.from_numpy(gl['labels']).to(cuda), use_graph=use_graph)
        y = res['y'].cpu().numpy()
        assert res['n_exec'].cpu().tolist() == gl['n_exec'].tolist()
        assert float(np.abs(y - gl['y_final']).max()) < TOL_F32
        assert float((y.argmax(1) == gl['y_final'].argmax(1)).mean()) >= MIN_ARGMAX_F32


def test_fp32x3_free_running_loop_vs_oracle_64x80(cuda, built_f32):
    """FCN8 + 10 free-running iterations on 2 images at 64x80, every iteration compared with the oracle
    run live on the CPU (a few seconds): per-iteration y within 2e-3, argmax agreement >= 99.9 %."""
    from iterative_inference_segm_b200.functions import function_pred_fcn, IterativeInference
    pf, pd, fcn, dae = built_f32
    X, L, lab = weights.synthetic_batch(2, 64, 80, NCLS, seed=3)
    h_o, y_o = nets.fcn8_forward(pf, X, NCLS)
    h_d, y_d = function_pred_fcn(fcn)(X.to(cuda))
    assert float((y_d.cpu() - y_o).abs().max()) < TOL_F32
    ii = IterativeInference(dae, NCLS, [NCLS])
    y = y_d
    for it in range(10):
        p_o = nets.dae_forward(pd, y_o, h_o, 100)
        y_o = torch.clamp(y_o - 0.05 * (y_o - p_o), 0, 1)
        y = ii.run(h_d, y, 0.05, 1, eps=0.0, use_graph=False)['y'].clone()
        assert float((y.cpu() - y_o).abs().max()) < TOL_F32, it
        assert float((y.cpu().argmax(1) == y_o.argmax(1)).float().mean()) >= MIN_ARGMAX_F32, it


@pytest.mark.parametrize('precision', ['bf16', 'mixed'])
def test_fused_update_epilogue_equals_standalone_kernel(cuda, built, precision):
    """up_conv1 with the softmax tail + y update fused in its epilogue gives bit-identical y (and the same
    executed-iteration counts) as logits -> iiseg_softmax_update; the per-iteration norms agree to
    fp32 rounding (fixed-point integer sum vs ordered partial sums).  'mixed': the epilogue writes the
    (hi | lo) pair of y that the fp32-accurate first conv reads."""
    from iterative_inference_segm_b200.functions import IterativeInference
    pf, pd, fcn, dae = built if precision == 'bf16' else _nets(cuda, precision)
    gd = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    h = torch.from_numpy(gd['h']).to(cuda)
    y0 = torch.from_numpy(np.concatenate([gd['y'], gd['y'][:, ::-1].copy()])).to(cuda)     # 2 images
    h2 = torch.cat([h, h])
    res = {}
    for fuse in (True, False):
        ii = IterativeInference(dae, NCLS, [NCLS], fuse_update=fuse)
        r = ii.run(h2, y0, 0.05, 5, eps=0.0, use_graph=False)
        res[fuse] = (r['y'].clone(), r['norm_hist'].clone(), r['n_exec'].clone())
    assert torch.equal(res[True][0], res[False][0])
    assert torch.equal(res[True][2], res[False][2])
    assert torch.allclose(res[True][1], res[False][1], rtol=1e-5, atol=0)
    # frozen images stay untouched in the fused path too
    ii = IterativeInference(dae, NCLS, [NCLS], fuse_update=True)
    one = ii.run(h2, y0, 0.05, 1, eps=0.0, use_graph=False)['y'].clone()
    many = ii.run(h2, y0, 0.05, 3, eps=1e9, use_graph=False)
    assert many['n_exec'].cpu().tolist() == [1, 1] and torch.equal(many['y'], one)


def test_config1_224x224_one_image_10_steps(cuda, built, built_f32):
    """BASELINE.json configs[0]: FCN8 + DAE_h, 1 synthetic 224x224 image, 11 classes, 10 steps, step 0.05 -- the
    reference's own CPU-runnable case -- free-running against the oracle, both arithmetic variants."""
    from iterative_inference_segm_b200.functions import function_pred_fcn, IterativeInference
    X, L, lab = weights.synthetic_batch(1, 224, 224, NCLS, seed=11)
    pf, pd = built[0], built[1]
    h_o, y_o = nets.fcn8_forward(pf, X, NCLS)
    for _ in range(10):
        y_o = torch.clamp(y_o - 0.05 * (y_o - nets.dae_forward(pd, y_o, h_o, 100)), 0, 1)
    for (_, _, fcn, dae), tol, agree in ((built_f32, TOL_F32, MIN_ARGMAX_F32), (built, 5e-2, 0.98)):
        h_d, y0_d = function_pred_fcn(fcn)(X.to(cuda))
        res = IterativeInference(dae, NCLS, [NCLS]).run(h_d, y0_d, 0.05, 10, labels=lab.to(torch.int32).to(cuda))
        y = res['y'].cpu()
        assert res['n_exec'].cpu().tolist() == [10]
        assert float((y - y_o).abs().max()) < tol, float((y - y_o).abs().max())
        assert float((y.argmax(1) == y_o.argmax(1)).float().mean()) >= agree
        onehot = np.eye(NCLS + 1, dtype=np.float32)[lab.numpy()].transpose(0, 3, 1, 2)
        assert np.array_equal(res['cm'].sum(0).cpu().numpy().reshape(NCLS, NCLS), M.confusion_matrix(y.numpy(), onehot, NCLS))


def test_full_size_properties_360x480(cuda, built):
    """BASELINE.json configs[1] size (360x480, 11 classes), checked through size-independent properties:
    probabilities are a distribution, the confusion matrix counts exactly the non-void pixels, replays are
    bit-identical, and images do not interact -- an image iterated inside a batch of 3 gives bit-identical
    y to the same image iterated alone (no cross-image coupling, position-independent K order)."""
    from iterative_inference_segm_b200.functions import IterativeInference
    pf, pd, fcn, dae = built
    X, L, lab = weights.synthetic_batch(3, 360, 480, NCLS, seed=21)
    net = fcn[0].net
    out = net.forward(X.to(cuda), want=('pool4', 'probs_dimshuffle'))
    h, y0 = out['pool4'].clone(), out['probs_dimshuffle'].clone()
    assert torch.allclose(y0.sum(1), torch.ones_like(y0[:, 0]), atol=1e-5) and float(y0.min()) >= 0.0
    labels = lab.to(torch.int32).to(cuda)
    ii = IterativeInference(dae, NCLS, [NCLS])
    r1 = ii.run(h, y0, 0.05, 5, labels=labels)
    y_a, cm_a = r1['y'].clone(), r1['cm'].clone()
    r2 = ii.run(h, y0, 0.05, 5, labels=labels)
    assert torch.equal(r2['y'], y_a) and torch.equal(r2['cm'], cm_a)                       # deterministic replay
    assert float(y_a.min()) >= 0.0 and float(y_a.max()) <= 1.0
    assert cm_a.sum(1).cpu().tolist() == [int((lab[i] != NCLS).sum()) for i in range(3)]     # every non-void pixel counted once
    assert r1['counts'][:, 1].cpu().tolist() == [int((lab[i] != NCLS).sum()) for i in range(3)]
    alone = IterativeInference(dae, NCLS, [NCLS]).run(h[1:2].contiguous(), y0[1:2].contiguous(), 0.05, 5,
                                                      labels=labels[1:2].contiguous())
    assert torch.equal(alone['y'][0], y_a[1]) and torch.equal(alone['cm'][0], cm_a[1])        # images do not interact


def test_batch_composition_does_not_change_results(cuda, built):
    """Images are independent (no cross-image coupling, fixed accumulation order per output): three images run as
    one batch of 3 or as batches of 2 + 1 give bit-identical y and identical int64 confusion matrices -- the property
    image sharding across GPUs relies on."""
    from iterative_inference_segm_b200.functions import IterativeInference
    _, _, fcn, dae = built
    X, L, _ = weights.synthetic_batch(3, 32, 40, NCLS, seed=9)
    X, L = X.to(cuda), L.to(cuda)
    ii = IterativeInference(dae, NCLS, [NCLS])
    net = fcn[0].net

    def run(sl):
        out = net.forward(X[sl].contiguous(), want=('pool4', 'probs_dimshuffle'))
        r = ii.run(out['pool4'], out['probs_dimshuffle'], 0.05, 4, onehot=L[sl].contiguous())
        return r['y'].clone(), r['cm'].clone()
    y_all, cm_all = run(slice(0, 3))
    y_a, cm_a = run(slice(0, 2))
    y_b, cm_b = run(slice(2, 3))
    assert torch.equal(y_all[:2], y_a) and torch.equal(y_all[2:], y_b)
    assert torch.equal(cm_all.sum(0), cm_a.sum(0) + cm_b.sum(0))


# ---------------------------------------------------------------------------
# Full BASELINE size, end to end, against the live oracle
# ---------------------------------------------------------------------------
# bf16 variant's own tolerance at 360x480 x 50 iterations, true pipeline (device FCN8 -> device loop): measured with
# tests/parity_report.py on a B200 (profiles/r02_parity_360x480_50it_bf16.txt) and asserted with ~25 % headroom.
# Measured (1 image, seed 0): FCN8 y0 max-abs 2.40e-2 / argmax 99.43 %; true pipeline worst y 2.28e-2 (iteration 1: FCN8's y0
# error, decaying under the iteration) / worst argmax 98.93 % (iteration 50); teacher-forced p worst 6.2e-3; loop alone
# (oracle h, y0 in) worst y 2.4e-3 / argmax 99.33 %.
BF16_FULL = {'fcn_y0': 3.0e-2, 'true_y': 2.9e-2, 'true_argmax': 0.986, 'tf_p': 7.8e-3, 'loop_y': 3.0e-3, 'loop_argmax': 0.9915}


@pytest.mark.parametrize('precision', ['mixed', 'bf16'])
def test_full_size_end_to_end_vs_oracle(cuda, precision):
    """BASELINE.json configs[1] size: one 360x480 image, 50 iterations, step 0.05.  The TRUE pipeline -- device FCN8 ->
    device loop as one CUDA-graph replay, y recorded after every iteration -- against oracle FCN8 -> oracle loop run
    live on the host (~15 s), every iteration compared; plus the teacher-forced application (oracle y_k in) and the
    loop alone from the oracle's h / y0.  'mixed' is held to north_star's fp32 bar on every iteration; 'bf16' to its
    own measured tolerance."""
    from tests.parity_report import measure
    r = measure(360, 480, 50, 0.05, precision, n_images=1, verbose=False)
    if precision == 'mixed':
        assert r['fcn_y0_maxabs'] < TOL_F32 and r['fcn_argmax'] >= MIN_ARGMAX_F32, r
        assert r['true_y'] < TOL_F32 and r['true_argmax'] >= MIN_ARGMAX_F32, r       # every iteration, end to end
        assert r['tf_p'] < TOL_F32, r                                                 # every teacher-forced application
        assert r['loop_y'] < TOL_F32 and r['loop_argmax'] >= MIN_ARGMAX_F32, r
    else:
        assert r['fcn_y0_maxabs'] < BF16_FULL['fcn_y0'], r
        assert r['true_y'] < BF16_FULL['true_y'] and r['true_argmax'] >= BF16_FULL['true_argmax'], r
        assert r['tf_p'] < BF16_FULL['tf_p'], r
        assert r['loop_y'] < BF16_FULL['loop_y'] and r['loop_argmax'] >= BF16_FULL['loop_argmax'], r


def test_valid_sweep_matches_oracle_valid_mat(cuda):
    """iterative_inference_valid.py's accumulator against the ORACLE's (oracle.loop.inference_batch, restating
    iterative_inference_valid.py:231,286-297): same data, same weights, two step values, three iterations, an eps
    chosen so that some images leave the loop early.  Executed-iteration pattern identical; per-iteration [TP; union]
    counts equal up to the few pixels whose argmax differs (fp32-grade arithmetic: <= 0.1 % of the pixels)."""
    from iterative_inference_segm_b200.iterative_inference_valid import sweep
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    steps, n_it = [0.05, 0.5], 3
    it = _tiny_iter(3, 2)
    batches = [it.next() for _ in range(it.nbatches)]
    # first-iteration norms of every image in the oracle -> an eps that freezes about half of them after iteration 1
    norms = []
    for X, L in batches:
        h, y0 = nets.fcn8_forward(pf, torch.from_numpy(X), NCLS)
        for im in range(X.shape[0]):
            g = y0[im:im + 1] - nets.dae_forward(pd, y0[im:im + 1], h[im:im + 1], 100)
            norms.append(float(torch.linalg.vector_norm(g, dim=1).mean()))
    ns = sorted(norms)
    eps = 0.5 * (ns[0] + ns[1])               # between two images' norms: no decision sits on the threshold
    assert ns[0] < eps < ns[1]
    want = np.zeros((len(steps), 2, NCLS, n_it))
    n_exec_o = []
    for X, L in batches:
        h, y0 = nets.fcn8_forward(pf, torch.from_numpy(X), NCLS)
        for si, s in enumerate(steps):
            _, n_exec, _, vm = loop.inference_batch(pd, h, y0, s, n_it, 100, L=L, n_classes=NCLS, void_labels=[NCLS], eps=eps)
            want[si] += vm
            n_exec_o.append(n_exec)
    res, mats = sweep('camvid', 'fcn8', steps=steps, num_iter=n_it, dae_dict_updates=DAE_DICT, data_iter=_tiny_iter(3, 2),
                      fcn_params=pf, dae_params=pd, verbose=False, precision='mixed', eps=eps)
    assert sweep.last_graph_captures <= 2            # one graph per batch shape (2 and 1 images), NOT one per step value
    assert any(n < n_it for ne in n_exec_o for n in ne), 'eps did not trigger an early exit: the test would not cover it'
    # an image that left the loop contributes nothing afterwards: the zero pattern of the denominators is the oracle's
    assert np.array_equal(mats[:, 1].sum(1) == 0, want[:, 1].sum(1) == 0)
    npix = 32 * 40
    assert float(np.abs(mats - want).max()) <= max(2.0, 1e-3 * npix * 3), float(np.abs(mats - want).max())
    assert np.array_equal(mats[:, 1].sum(1) > 0, want[:, 1].sum(1) > 0)


def test_temperature_divides_the_upsample_layer(cuda):
    """models/fcn8.py:193-198: temperature T divides upsample.W and .b, i.e. y0 = softmax(logits / T).  T = 2.5 against
    the oracle's fcn8_forward(temperature=2.5), fp32-grade arithmetic; T is only applied with load_weights (as in the
    reference)."""
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.functions import function_pred_fcn
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    X, _, _ = weights.synthetic_batch(2, 32, 40, NCLS, seed=7)
    _, y_T = nets.fcn8_forward(pf, X, NCLS, temperature=2.5)
    _, y_1 = nets.fcn8_forward(pf, X, NCLS)
    assert float((y_T - y_1).abs().max()) > 0.05                       # the temperature matters on this input
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, load_weights=True,
                    temperature=2.5, precision='mixed')
    _, y_d = function_pred_fcn(fcn)(X.to(cuda))
    assert float((y_d.cpu() - y_T).abs().max()) < TOL_F32
    fcn1 = buildFCN8(3, None, n_classes=NCLS, layer=['pool4', 'probs_dimshuffle'], params=pf, load_weights=False,
                     temperature=2.5, precision='mixed')                # reference: T ignored without load_weights
    _, y_d1 = function_pred_fcn(fcn1)(X.to(cuda))
    assert float((y_d1.cpu() - y_1).abs().max()) < TOL_F32


def test_checkpoints_load_from_disk_in_the_reference_layout(cuda, tmp_path):
    """Positional .npz checkpoints (np.savez(path, *get_all_param_values(net)): arr_0..arr_k; models/fcn8.py:177-180,
    models/DAE_h.py:52-57) written where the reference's scripts look for them -- <weights_path>/<dataset>/fcn8_model.npz and
    <loadpath>/<dataset>/<experiment name>/dae_model_best.npz (iterative_inference.py:136,157) -- give bit-identical results
    to passing the same arrays in memory."""
    from iterative_inference_segm_b200.iterative_inference import inference
    from iterative_inference_segm_b200.helpers import build_experiment_name
    from iterative_inference_segm_b200.iterative_inference import DAE_DICT_DEFAULTS
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    dd = dict(DAE_DICT_DEFAULTS); dd.update(DAE_DICT)
    exp = build_experiment_name('fcn8', data_aug=False, ae_h=False, **dd)
    wdir, ldir = tmp_path / 'w', tmp_path / 'l'
    os.makedirs(wdir / 'camvid'); os.makedirs(ldir / 'camvid' / exp)
    weights.save_npz(str(wdir / 'camvid' / 'fcn8_model.npz'), pf)
    weights.save_npz(str(ldir / 'camvid' / exp / 'dae_model_best.npz'), pd)
    kw = dict(dae_dict_updates=DAE_DICT, savepath=str(tmp_path / 'out'), verbose=False)
    a = inference('camvid', 'fcn8', 0.05, 2, data_iter=_tiny_iter(2, 2), weights_path=str(wdir), loadpath=str(ldir), **kw)
    b = inference('camvid', 'fcn8', 0.05, 2, data_iter=_tiny_iter(2, 2), fcn_params=pf, dae_params=pd, loadpath=str(ldir), **kw)
    assert np.array_equal(a['cm'], b['cm']) and np.array_equal(a['jacc_tot'], b['jacc_tot']) and a['iterative'] == b['iterative']
    assert a['n_exec'] == b['n_exec'] == [2, 2]


# ---------------------------------------------------------------------------
# SURVEY 8(f4), first slice: the other unpool types of models/fcn_up.py
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['mixed', 'bf16'])
def test_unpool_type_standard_deconv_vs_oracle(cuda, precision):
    """unpool_type='standard' (models/fcn_up.py:37-63): Deconv2DLayer(4, stride=2, crop='valid') per level + centre-cropped
    skip-sum, no convolution on the way up.  One application and a 3-step loop against the oracle, odd and even sizes."""
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_dae, IterativeInference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=4, out_gain=0.1, unpool_type='standard')
    assert tuple(pd[12].shape) == (2048, 1024, 4, 4) and tuple(pd[22].shape) == (64, NCLS, 4, 4)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64,
                   conv_before_pool=1, additional_pool=2, skip=True, unpool_type='standard', params=pd, precision=precision)
    tol, agree = (TOL_F32, MIN_ARGMAX_F32) if precision == 'mixed' else (TOL_DAE_P, MIN_ARGMAX)
    for (H, W) in ((32, 40), (37, 45)):
        X, L, lab = weights.synthetic_batch(2, H, W, NCLS, seed=13)
        h, y0 = nets.fcn8_forward(pf, X, NCLS)
        p_o = nets.dae_forward(pd, y0, h, 100, unpool_type='standard')
        p_d = function_pred_dae(dae)(h.numpy(), y0.numpy())
        assert p_d.shape == tuple(p_o.shape)
        assert float(np.abs(p_d - p_o.numpy()).max()) < tol, float(np.abs(p_d - p_o.numpy()).max())
        y_o = y0.clone()
        for _ in range(3):
            y_o = torch.clamp(y_o - 0.05 * (y_o - nets.dae_forward(pd, y_o, h, 100, unpool_type='standard')), 0, 1)
        res = IterativeInference(dae, NCLS, [NCLS]).run(h.to(cuda), y0.to(cuda), 0.05, 3, eps=0.0, labels=lab.to(torch.int32).to(cuda))
        y = res['y'].cpu()
        assert float((y - y_o).abs().max()) < tol and float((y.argmax(1) == y_o.argmax(1)).float().mean()) >= agree
        onehot = np.eye(NCLS + 1, dtype=np.float32)[lab.numpy()].transpose(0, 3, 1, 2)
        assert np.array_equal(res['cm'].sum(0).cpu().numpy().reshape(NCLS, NCLS), M.confusion_matrix(y.numpy(), onehot, NCLS))


def test_unpool_type_inverse_is_the_tie_mask_unpool(cuda, built):
    """unpool_type='inverse' (models/fcn_up.py:76-79): lasagne's InverseLayer of the max-pool = the pool's input-gradient with
    the incoming map as upstream gradient; Theano's CPU MaxPoolGrad sends it to EVERY tied maximum, which is DePool2D's
    repeat * tie-mask.  Checked in the oracle with autograd-free arithmetic (the gradient rule written out), and on the device
    as bit-identical results of the two unpool types."""
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_dae
    from oracle import lasagne_semantics as LS
    torch.manual_seed(0)
    x = torch.relu(torch.randn(2, 8, 9, 11)).mul(2).round().div(2)          # many ties
    u = torch.randn(2, 8, 4, 5)
    # MaxPoolGrad written out: gx[i, j] += gz[i // 2, j // 2] if x[i, j] == max of its window
    gx = torch.zeros_like(x)
    pooled = LS.maxpool2(x)
    for i in range(8):
        for j in range(10):
            gx[:, :, i, j] = torch.where(x[:, :, i, j] == pooled[:, :, i // 2, j // 2], u[:, :, i // 2, j // 2], torch.zeros(()))
    assert torch.equal(gx, LS.depool2d(u, x))
    pf, pd, fcn, dae = built
    dae_inv = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64,
                       conv_before_pool=1, additional_pool=2, skip=True, unpool_type='inverse', params=pd)
    g = np.load(os.path.join(GOLD, 'dae_32x40.npz'))
    assert np.array_equal(function_pred_dae(dae_inv)(g['h'], g['y']), function_pred_dae(dae)(g['h'], g['y']))
    assert float(np.abs(function_pred_dae(dae_inv)(g['h'], g['y']) - nets.dae_forward(pd, torch.from_numpy(g['y']), torch.from_numpy(g['h']), 100,
                                                                                       unpool_type='inverse').numpy()).max()) < TOL_DAE_P


def _with_bn(pd, n_levels, unpool_type, seed=21):
    """Insert BatchNormLayer parameters (beta, gamma, mean, inv_std -- lasagne's order) behind every conv of a DAE_h
    checkpoint, as bn=1 saves them; a few negative gammas so that nothing relies on a sign."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    k_up = 1 if unpool_type == 'standard' else 2
    for i in range(2 * n_levels):
        W, b = pd[2 * i], pd[2 * i + 1]
        out += [W, b]
        if i >= n_levels and unpool_type == 'standard':
            continue
        c = b.shape[0]
        gamma = 0.5 + torch.rand(c, generator=gen)
        gamma[::7] *= -1.0
        out += [0.1 * torch.randn(c, generator=gen), gamma, 0.05 * torch.randn(c, generator=gen), 0.5 + 1.5 * torch.rand(c, generator=gen)]
    return out


@pytest.mark.parametrize('unpool_type', ['trackind', 'standard'])
def test_dae_with_batchnorm_vs_oracle(cuda, unpool_type):
    """bn=1 (models/fcn_down.py:113-115, models/fcn_up.py:91-93) at inference: BatchNormLayer on its stored averages.  The
    contracting path applies it after the rectifier and BEFORE the pool, so the pool maxima and DePool2D's tie masks are those of
    the normalised maps (negative gammas included); the expanding path's layer is folded into its conv."""
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_dae
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = _with_bn(weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1, unpool_type=unpool_type), 6, unpool_type)
    assert len(pd) == (72 if unpool_type == 'trackind' else 48)
    X, _, _ = weights.synthetic_batch(2, 37, 45, NCLS, seed=17)
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    p_o = nets.dae_forward(pd, y0, h, 100, unpool_type=unpool_type, bn=True)
    import warnings
    # Both unpool types hold the fp32 bar.  With DePool2D the masks come from the reference's batch-statistics mask pass
    # (oracle/nets.py:batchnorm_batch_stats, found by executing the reference: tests/golden/ref_bn.npz): measured 1.3e-3
    # (fp32x3) / 1.4e-3 (mixed) on this test's harsh gains (gamma * inv_std up to 3 per level), 8e-4 on the reference's own run.
    # History: ~4e-3 with the stored-average masks of round 2's first build, 6.6e-3 when the batch statistics were taken from
    # bf16-rounded activations; the ties are decided on the fp32 accumulators and the layer evaluated in lasagne's operation order.
    tol_f32 = TOL_F32
    tol_bf16 = 2e-2 if unpool_type == 'standard' else 6e-2         # measured 4.2e-2 with DePool2D (tie flips x BN gains)
    for precision, tol in (('fp32x3', tol_f32), ('mixed', tol_f32), ('bf16', tol_bf16)):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64,
                           conv_before_pool=1, additional_pool=2, skip=True, unpool_type=unpool_type, bn=1, params=pd, precision=precision)
        p_d = function_pred_dae(dae)(h.numpy(), y0.numpy())
        err = float(np.abs(p_d - p_o.numpy()).max())
        print('bn=1 %s %s: p max-abs %.3e' % (unpool_type, precision, err))
        assert err < tol, (precision, err)


def test_stochastic_mask_subgraph_opt_in(cuda):
    """The reference's DePool2D builds its masks with get_output(...) WITHOUT deterministic=True (layers/mylayers.py:91-93): for a
    DAE built with noise > 0 (the README's 0.5) the tie masks come from a separate pass of the contracting path on
    y + N(0, noise^2), even at inference.  `buildDAE(..., noise > 0)` does the same (`stochastic_masks=False` asks for the
    deterministic graph and warns): with the SAME noise tensor the device application tracks the oracle's `mask_source_y` form
    (here one shared draw; test_noised_mask_subgraphs_vs_reference_run covers the reference's one draw per DePool2D)."""
    import warnings
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import IterativeInference
    from iterative_inference_segm_b200 import _kernels as K
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1)
    kw = dict(padding=100, concat_h=['pool4'], n_filters=64, conv_before_pool=1, additional_pool=2, skip=True, unpool_type='trackind',
              params=pd, precision='mixed')
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        det = buildDAE([None], None, NCLS, nb_features_to_concat=512, noise=0.5, stochastic_masks=False, **kw)
        assert any('deterministic masks' in str(x.message) for x in w) and det.net.mask_noise == 0.0
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, noise=0.5, **kw)          # default: as the reference
    net = dae.net
    assert net.mask_noise == 0.5
    X, L, lab = weights.synthetic_batch(2, 32, 40, NCLS, seed=23)
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    noise = torch.randn(y0.shape, generator=torch.Generator().manual_seed(5))
    p_o = nets.dae_forward(pd, y0, h, 100, mask_source_y=y0 + 0.5 * noise)
    p_det = nets.dae_forward(pd, y0, h, 100)
    assert float((p_o - p_det).abs().max()) > 2 * TOL_F32               # the noised masks matter (7.7e-3 here)
    yd = y0.to(cuda)
    logits = net.logits(K.pack_nchw(h.to(cuda), net.h_pad, split=True), K.pack_nchw(yd, net.y_cpad, split=True), y_f32=yd,
                        noise=noise.to(cuda))
    p_d = torch.empty_like(yd)
    K.softmax_nchw(logits, NCLS, p_d)
    assert float((p_d.cpu() - p_o).abs().max()) < TOL_F32, float((p_d.cpu() - p_o).abs().max())
    # the loop draws its own noise every iteration (inside the captured graph): runs, stays a distribution, differs from the
    # deterministic loop, and two replays differ from each other (fresh noise)
    ii = IterativeInference(dae, NCLS, [NCLS])
    r1 = ii.run(h.to(cuda), yd, 0.05, 3, eps=0.0, labels=lab.to(torch.int32).to(cuda))['y'].clone()
    r2 = ii.run(h.to(cuda), yd, 0.05, 3, eps=0.0, labels=lab.to(torch.int32).to(cuda))['y'].clone()
    assert bool(torch.isfinite(r1).all()) and float(r1.min()) >= 0 and float(r1.max()) <= 1
    assert not torch.equal(r1, r2)


@pytest.mark.parametrize('unpool_type', ['trackind', 'standard'])
def test_dae_without_skip_connections_and_ae_h_flag(cuda, unpool_type):
    """skip=False (models/fcn_up.py:103-113): the up-conv of each level is only centre-cropped to the size of pool_{p-1}, nothing is
    added.  ae_h=True only renames layers / freezes parameters for the training loss: same inference graph."""
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_dae
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=1, out_gain=0.1, unpool_type=unpool_type)
    X, _, _ = weights.synthetic_batch(2, 37, 45, NCLS, seed=19)
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    p_o = nets.dae_forward(pd, y0, h, 100, unpool_type=unpool_type, skip=False)
    assert float((p_o - nets.dae_forward(pd, y0, h, 100, unpool_type=unpool_type)).abs().max()) > 1e-3      # the skips matter
    kw = dict(nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64, conv_before_pool=1, additional_pool=2,
              unpool_type=unpool_type, params=pd, precision='mixed')
    dae = buildDAE([None], None, NCLS, skip=False, **kw)
    p_d = function_pred_dae(dae)(h.numpy(), y0.numpy())
    assert float(np.abs(p_d - p_o.numpy()).max()) < TOL_F32, float(np.abs(p_d - p_o.numpy()).max())
    dae_ae = buildDAE([None], None, NCLS, skip=False, ae_h=True, **kw)
    assert np.array_equal(function_pred_dae(dae_ae)(h.numpy(), y0.numpy()), p_d)


def test_dae_with_two_convs_before_each_pool(cuda):
    """conv_before_pool=2 (models/fcn_down.py:83-104): conv_p_1, conv_p_2 (rectified, 'same' after the pad-100 first one), then the
    pool; 36 parameter arrays.  One application and a 3-step loop against the oracle at the fp32 bar."""
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_dae, IterativeInference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 512, seed=6, out_gain=0.1, conv_before_pool=2)
    assert len(pd) == 36
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=0.0, n_filters=64,
                   conv_before_pool=2, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision='mixed')
    X, L, lab = weights.synthetic_batch(2, 37, 45, NCLS, seed=29)
    h, y0 = nets.fcn8_forward(pf, X, NCLS)
    p_o = nets.dae_forward(pd, y0, h, 100, conv_before_pool=2)
    p_d = function_pred_dae(dae)(h.numpy(), y0.numpy())
    assert float(np.abs(p_d - p_o.numpy()).max()) < TOL_F32, float(np.abs(p_d - p_o.numpy()).max())
    y_o = y0.clone()
    for _ in range(3):
        y_o = torch.clamp(y_o - 0.05 * (y_o - nets.dae_forward(pd, y_o, h, 100, conv_before_pool=2)), 0, 1)
    y = IterativeInference(dae, NCLS, [NCLS]).run(h.to(cuda), y0.to(cuda), 0.05, 3, eps=0.0)['y'].cpu()
    assert float((y - y_o).abs().max()) < TOL_F32 and float((y.argmax(1) == y_o.argmax(1)).float().mean()) >= MIN_ARGMAX_F32


@pytest.mark.parametrize('precision', ['mixed', 'bf16'])
def test_dae_conditioned_on_the_input_image(cuda, precision):
    """concat_h=['input'] (models/model_helpers.py:86-96, models/fcn_down.py:69-74,90-95): the conditioning tensor is the image
    (the segmentation net's 'input' layer, 3 channels), concatenated BEFORE y at the DAE's input; no pool precedes it, so the
    DAE has `additional_pool` levels and its first conv is 'same'-padded (no pad 100).  The iteration-invariant image half of
    conv1_1 is hoisted like the h half of conv5_1.  Through the drop-in callables: buildFCN8(layer=['input', ...]) ->
    pred_fcn_fn -> the loop, against the oracle."""
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200.functions import function_pred_fcn, function_pred_dae, IterativeInference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pd = weights.synthetic_dae_params(NCLS, 3, seed=8, out_gain=0.1, concat_h=('input',), additional_pool=3)
    assert len(pd) == 12 and tuple(pd[0].shape) == (64, 3 + NCLS, 3, 3)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['input', 'probs_dimshuffle'], params=pf, precision=precision)
    assert fcn[0].output_shape[1] == 3
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=fcn[0].output_shape[1], padding=100, concat_h=['input'], noise=0.0,
                   n_filters=64, conv_before_pool=1, additional_pool=3, skip=True, unpool_type='trackind', params=pd, precision=precision)
    tol, agree = (TOL_F32, MIN_ARGMAX_F32) if precision == 'mixed' else (TOL_DAE_P, MIN_ARGMAX)
    X, L, lab = weights.synthetic_batch(2, 40, 56, NCLS, seed=31)
    _, y0 = nets.fcn8_forward(pf, X, NCLS)
    h_d, y0_d = function_pred_fcn(fcn)(X.numpy())
    assert np.array_equal(h_d, X.numpy())
    kw = dict(concat_h=('input',), additional_pool=3)
    p_o = nets.dae_forward(pd, y0, X, 100, **kw)
    p_d = function_pred_dae(dae)(X.numpy(), y0.numpy())
    assert float(np.abs(p_d - p_o.numpy()).max()) < tol, float(np.abs(p_d - p_o.numpy()).max())
    y_o = y0.clone()
    for _ in range(3):
        y_o = torch.clamp(y_o - 0.05 * (y_o - nets.dae_forward(pd, y_o, X, 100, **kw)), 0, 1)
    res = IterativeInference(dae, NCLS, [NCLS]).run(X.to(cuda), y0.to(cuda), 0.05, 3, eps=0.0, labels=lab.to(torch.int32).to(cuda))
    y = res['y'].cpu()
    assert float((y - y_o).abs().max()) < tol and float((y.argmax(1) == y_o.argmax(1)).float().mean()) >= agree


def test_ctx_conv_kernel_vs_torch(cuda):
    """csrc/contextmod.cu against conv2d on the same values: dilations 1..16, 'same' zero padding, the output window inside
    a larger tensor (PadLayer interior), the hoisted addend, frozen images, the padded <16,16> instantiation and the fused
    1x1 tail.  fp32 FMA chains in a different order than the CPU's: 1e-5 of the output scale."""
    from iterative_inference_segm_b200 import _kernels as K
    import torch.nn.functional as Fn
    g = torch.Generator().manual_seed(5)
    for (cin, cout, dil, H, W) in [(11, 11, 1, 37, 70), (11, 11, 16, 40, 97), (11, 11, 8, 50, 130), (11, 11, 2, 16, 65),
                                   (3, 11, 1, 21, 33), (5, 7, 4, 30, 65), (16, 16, 2, 19, 40)]:
        x = torch.randn(3, cin, H, W, generator=g)
        Wt = torch.randn(cin, cout, 3, 3, generator=g) / (cin * 9) ** 0.5          # DilatedConv2DLayer layout (in, out, r, s)
        b = torch.randn(cout, generator=g)
        wk = np.ascontiguousarray(Wt.permute(0, 2, 3, 1).numpy())
        ref = torch.relu(Fn.conv2d(x, Wt.permute(1, 0, 2, 3), b, dilation=dil))
        OH, OW = H - 2 * dil, W - 2 * dil
        if OH >= 1 and OW >= 1:
            out = torch.full((3, cout, OH + 5, OW + 9), -7.0, device=cuda)
            act = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
            K.ctx_conv(x.to(cuda), wk, b.numpy(), dil, out, relu=True, out_origin=(2, 3), size=(OH, OW), active=act)
            got = out.cpu()
            assert float((got[[0, 2], :, 2:2 + OH, 3:3 + OW] - ref[[0, 2]]).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
            assert float((got[1] + 7.0).abs().max()) == 0.0                        # frozen image untouched
            got[:, :, 2:2 + OH, 3:3 + OW] = -7.0
            assert float((got + 7.0).abs().max()) == 0.0                           # nothing outside the window
        # 'same' zero padding + addend, linear
        add = torch.randn(3, cout, H, W, generator=g)
        ref = Fn.conv2d(x, Wt.permute(1, 0, 2, 3), b, dilation=dil, padding=dil) + add
        out = torch.empty((3, cout, H, W), device=cuda)
        K.ctx_conv(x.to(cuda), wk, b.numpy(), dil, out, relu=False, origin=(-dil, -dil), check=True, addend=add.to(cuda))
        assert float((out.cpu() - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
        if cin != 3:
            # fused 1x1 tail -> NHWC16 logits rows, zero beyond C2
            c2 = min(cout, 11)
            W2 = torch.randn(cout, c2, generator=g)
            b2 = torch.randn(c2, generator=g)
            ref = torch.einsum('nfhw,fg->nhwg', torch.relu(Fn.conv2d(x, Wt.permute(1, 0, 2, 3), b, dilation=dil, padding=dil)), W2) + b2
            lg = torch.full((3, H, W, 16), 9.0, device=cuda)
            K.ctx_conv(x.to(cuda), wk, b.numpy(), dil, lg, relu=True, origin=(-dil, -dil), check=True,
                       tail=(np.ascontiguousarray(W2.numpy()), b2.numpy()))
            assert float((lg[..., :c2].cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
            assert float(lg[..., c2:].abs().max()) == 0.0
        # channels-last tensors (the module's intermediate layout): [N,H,W,pad4(C)], pad channels of the input hold garbage
        # (they have no weights), pad channels of the output are written as zero
        if cin >= 4 and OH >= 1 and OW >= 1:
            ci4, co4 = (cin + 3) & ~3, (cout + 3) & ~3
            xcl = torch.full((3, H, W, ci4), 123.0)
            xcl[..., :cin] = x.permute(0, 2, 3, 1)
            ref = torch.relu(Fn.conv2d(x, Wt.permute(1, 0, 2, 3), b, dilation=dil))
            add = torch.randn(3, OH, OW, co4, generator=g)
            add[..., cout:] = 0
            refa = torch.relu(Fn.conv2d(x, Wt.permute(1, 0, 2, 3), b, dilation=dil) + add[..., :cout].permute(0, 3, 1, 2))
            ocl = torch.full((3, OH + 4, OW + 6, co4), -7.0, device=cuda)
            K.ctx_conv(xcl.to(cuda), wk, b.numpy(), dil, ocl, relu=True, out_origin=(1, 2), size=(OH, OW), in_nhwc=True, out_nhwc=True,
                       addend=add.to(cuda))
            got = ocl.cpu()[:, 1:1 + OH, 2:2 + OW]
            assert float((got[..., :cout].permute(0, 3, 1, 2) - refa).abs().max()) < 1e-5 * max(1.0, float(refa.abs().max()))
            assert co4 == cout or float(got[..., cout:].abs().max()) == 0.0
            opl = torch.empty((3, cout, OH, OW), device=cuda)           # channels-last in, planar out
            K.ctx_conv(xcl.to(cuda), wk, b.numpy(), dil, opl, relu=True, in_nhwc=True)
            assert float((opl.cpu() - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
            ocl2 = torch.empty((3, OH, OW, co4), device=cuda)           # planar in, channels-last out
            K.ctx_conv(x.to(cuda), wk, b.numpy(), dil, ocl2, relu=True, out_nhwc=True)
            assert float((ocl2.cpu()[..., :cout].permute(0, 3, 1, 2) - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
            c2 = min(cout, 11)
            W2 = torch.randn(cout, c2, generator=g)
            b2 = torch.randn(c2, generator=g)
            reft = torch.einsum('nfhw,fg->nhwg', ref, W2) + b2
            lg = torch.empty((3, OH, OW, 16), device=cuda)
            K.ctx_conv(xcl.to(cuda), wk, b.numpy(), dil, lg, relu=True, in_nhwc=True, tail=(np.ascontiguousarray(W2.numpy()), b2.numpy()))
            assert float((lg[..., :c2].cpu() - reft).abs().max()) < 2e-5 * max(1.0, float(reft.abs().max()))


def test_contextmod_dae_vs_oracle(cuda):
    """kind='contextmod' (models/contextmod_dae.py:19-138, the reference CLI's default DAE) through the drop-in callables:
    buildFCN8(layer=['input', 'probs_dimshuffle']) -> pred_fcn_fn -> pred_dae_fn / de_fn / the device loop with early exit
    and metrics, against the oracle.  The module computes in fp32 on the device as in the reference: 2e-5 on probabilities
    (fp32 summation order), argmax identical away from exact ties."""
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.contextmod_dae import buildDAE_contextmod
    from iterative_inference_segm_b200.functions import function_pred_fcn, function_pred_dae, function_de, IterativeInference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    pc = weights.synthetic_contextmod_params(NCLS, 3, seed=3)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=['input', 'probs_dimshuffle'], params=pf, precision='mixed')
    dae = buildDAE_contextmod([None], None, NCLS, concat_h=['input'], noise=0.0, params=pc,
                              nb_features_to_concat=fcn[0].output_shape[1])
    X, L, lab = weights.synthetic_batch(3, 45, 70, NCLS, seed=12)
    _, y0 = nets.fcn8_forward(pf, X, NCLS)
    p_o = nets.contextmod_forward(pc, y0, X)
    p_d = function_pred_dae(dae)(X.numpy(), y0.numpy())
    e = float(np.abs(p_d - p_o.numpy()).max())
    assert e < 2e-5, e
    g_d = function_de(dae)(X.numpy(), y0.numpy())
    assert float(np.abs(g_d - (y0 - p_o).numpy()).max()) < 2e-5
    # the loop: 6 iterations, step 0.5, an eps that freezes some images early (oracle.loop semantics)
    from oracle import loop as oloop
    step, n_it = 0.5, 6
    ys, norms = [], []
    for b in range(3):
        y = y0[b:b + 1].clone()
        nb = []
        for _ in range(n_it):
            gr = y - nets.contextmod_forward(pc, y, X[b:b + 1])
            y = torch.clamp(y - step * gr, 0, 1)
            nb.append(float(torch.linalg.vector_norm(gr, dim=1).mean()))
        ys.append(y); norms.append(nb)
    eps = 0.5 * (sorted(n[2] for n in norms)[0] + sorted(n[2] for n in norms)[1])      # one image stops after 3 iterations
    res = IterativeInference(dae, NCLS, [NCLS]).run(X.to(cuda), y0.to(cuda), step, n_it, eps=eps, labels=lab.to(torch.int32).to(cuda))
    n_exec = res['n_exec'].cpu().tolist()
    exp_exec = [next((k + 1 for k, v in enumerate(nb) if v < eps), n_it) for nb in norms]
    assert n_exec == exp_exec and min(n_exec) < n_it, (n_exec, exp_exec)
    for b in range(3):
        y = y0[b:b + 1].clone()
        for _ in range(n_exec[b]):
            y = torch.clamp(y - step * (y - nets.contextmod_forward(pc, y, X[b:b + 1])), 0, 1)
        assert float((res['y'][b:b + 1].cpu() - y).abs().max()) < 5e-5


@pytest.mark.parametrize('kind', ['contextmod', 'fcn8'])
def test_inference_script_with_the_other_dae_kinds(cuda, tmp_path, kind):
    """inference(...) with the reference CLI's default dae_dict (iterative_inference.py:355-362: kind='contextmod',
    concat_h=['input'], step 1.0 from :341) and with inference()'s own default kind='fcn8' (:64, here with concat_h=['pool4']):
    the fused device loop equals the literal per-image loop over the callables and the saved batches track the oracle."""
    from iterative_inference_segm_b200.iterative_inference import inference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    if kind == 'contextmod':
        pc = weights.synthetic_contextmod_params(NCLS, 3, seed=3)
        dd = dict(DAE_DICT, kind='contextmod', concat_h=['input'])
    else:
        pc = weights.synthetic_fcn8_params(NCLS, NCLS, seed=6, logit_gain=10.0, concat=('pool4', 512))
        dd = dict(DAE_DICT, kind='fcn8', concat_h=['pool4'])
    kw = dict(dae_dict_updates=dd, savepath=str(tmp_path), loadpath=str(tmp_path), fcn_params=pf, dae_params=pc, verbose=False,
              precision='mixed')
    a = inference('camvid', 'fcn8', 1.0, 3, data_iter=_tiny_iter(), fused=True, save_batches=True, **kw)
    b = inference('camvid', 'fcn8', 1.0, 3, data_iter=_tiny_iter(), fused=False, **kw)
    assert a['n_exec'] == b['n_exec'] == [3, 3, 3]
    assert np.array_equal(a['jacc_tot'], b['jacc_tot'])
    it = _tiny_iter()
    for i in range(it.nbatches):
        X, L = it.next()
        Xt = torch.from_numpy(X)
        h, y = nets.fcn8_forward(pf, Xt, NCLS)
        for _ in range(3):
            p = nets.contextmod_forward(pc, y, Xt) if kind == 'contextmod' else nets.fcn8_dae_forward(pc, y, h, NCLS, concat_h=('pool4',))
            y = torch.clamp(y - 1.0 * (y - p), 0, 1)
        saved = np.load(os.path.join(a['savepath'], 'batch%d.npz' % i))
        assert float(np.abs(saved['Y_ii'] - y.numpy()).max()) < TOL_F32
        assert float((saved['Y_ii'].argmax(1) == y.numpy().argmax(1)).mean()) >= MIN_ARGMAX_F32


@pytest.mark.parametrize('concat_h,precision', [('pool4', 'mixed'), ('pool4', 'bf16'), ('input', 'mixed'), ('pool2', 'mixed')])
def test_fcn8_shaped_dae_vs_oracle(cuda, concat_h, precision):
    """kind='fcn8' (models/fcn8_dae.py:19-271): the FCN8 graph on y with h concatenated at 'input' / 'poolN', through the
    drop-in callables and the device loop, against the oracle.  The h half of the widened conv is hoisted (computed with the
    first iteration, reused by the next ones)."""
    from iterative_inference_segm_b200.models.fcn8 import buildFCN8
    from iterative_inference_segm_b200.models.fcn8_dae import buildFCN8_DAE
    from iterative_inference_segm_b200.functions import function_pred_fcn, function_pred_dae, IterativeInference
    pf = weights.synthetic_fcn8_params(3, NCLS, seed=0, logit_gain=10.0)
    fcn = buildFCN8(3, None, n_classes=NCLS, layer=[concat_h, 'probs_dimshuffle'], params=pf, precision=precision)
    nb_h = fcn[0].output_shape[1]
    pdae = weights.synthetic_fcn8_params(NCLS, NCLS, seed=6, logit_gain=10.0, concat=(concat_h, nb_h))
    dae = buildFCN8_DAE([None], None, NCLS, nb_in_channels=NCLS, concat_h=[concat_h], noise=0.0, params=pdae, precision=precision,
                        nb_features_to_concat=nb_h)
    tol, agree = (TOL_F32, MIN_ARGMAX_F32) if precision == 'mixed' else (TOL_FCN_PROBS, 0.98)
    X, L, lab = weights.synthetic_batch(2, 32, 40, NCLS, seed=17)
    if concat_h == 'input':
        h_o, y0 = X, nets.fcn8_forward(pf, X, NCLS, layer=('probs_dimshuffle',))[0]
    else:
        h_o, y0 = nets.fcn8_forward(pf, X, NCLS, layer=(concat_h, 'probs_dimshuffle'))
    p_o = nets.fcn8_dae_forward(pdae, y0, h_o, NCLS, concat_h=(concat_h,))
    p_d = function_pred_dae(dae)(h_o.numpy(), y0.numpy())
    assert float(np.abs(p_d - p_o.numpy()).max()) < tol, float(np.abs(p_d - p_o.numpy()).max())
    y_o = y0.clone()
    for _ in range(3):
        y_o = torch.clamp(y_o - 0.5 * (y_o - nets.fcn8_dae_forward(pdae, y_o, h_o, NCLS, concat_h=(concat_h,))), 0, 1)
    res = IterativeInference(dae, NCLS, [NCLS]).run(h_o.to(cuda), y0.to(cuda), 0.5, 3, eps=0.0, labels=lab.to(torch.int32).to(cuda))
    y = res['y'].cpu()
    print('fcn8 dae %s %s: p %.2e  y %.2e  argmax %.4f' % (concat_h, precision, float(np.abs(p_d - p_o.numpy()).max()),
                                                         float((y - y_o).abs().max()), float((y.argmax(1) == y_o.argmax(1)).float().mean())))
    assert float((y - y_o).abs().max()) < tol and float((y.argmax(1) == y_o.argmax(1)).float().mean()) >= agree


# ---- the CUDA path against OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_*.npz) ------------------------------------------
# The fixtures hold what the reference's own iterative_inference.py:inference / iterative_inference_valid.py:inference wrote
# and printed when executed through oracle/refrun (tests/golden/make_reference_golden.py).  These tests call this package's
# drop-ins for the same two functions, with the same arguments, the same checkpoints on disk in the reference's directory
# layout and the same data iterator, and compare what they save and return.
from tests import reference_fixtures as RF  # noqa: E402

# ref_bn: with bn=1 the reference's DePool2D mask sub-graph normalises with BATCH statistics (lasagne BatchNormLayer under
# deterministic=False, layers/mylayers.py:91-93); see test_dae_with_batchnorm_vs_oracle and DESIGN.md 3.10.
_REF_PRECISIONS = [('mixed', 2e-3, 0.999, 2e-3), ('bf16', 3e-2, 0.99, 2e-2)]      # (precision, max-abs on y, argmax agreement, metric tolerance)


def _reference_layout(tmp_path, case, exp_name):
    """Checkpoints on disk where the reference looks for them: WEIGHTS_PATH/<dataset>/fcn8_model.npz (iterative_inference.py:137)
    and LOADPATH/<dataset>/<exp_name>/dae_model_best.npz (:93, :150)."""
    wdir = tmp_path / 'weights' / 'camvid'
    wdir.mkdir(parents=True)
    weights.save_npz(str(wdir / 'fcn8_model.npz'), weights.synthetic_fcn8_params(3, NCLS, **RF.G.FCN8_WEIGHTS))
    if case.get('segm_net') == 'densenet':          # <weights_path>/<dataset>/DenseNet103/weights/FC-DenseNet103_weights.npz (models/FCDenseNet.py:198)
        ddir = wdir / 'DenseNet103' / 'weights'
        ddir.mkdir(parents=True)
        weights.save_npz(str(ddir / 'FC-DenseNet103_weights.npz'), RF.G.case_densenet_params(case))
    ldir = tmp_path / 'load' / 'camvid' / exp_name
    ldir.mkdir(parents=True)
    weights.save_npz(str(ldir / 'dae_model_best.npz'), RF.G.case_dae_params(case))
    return str(tmp_path / 'weights'), str(tmp_path / 'load'), str(tmp_path / 'save')


# (ref_noise draws random numbers: test_noised_mask_subgraphs_vs_reference_run replays it with the logged draws)
@pytest.mark.parametrize('name', [n for n in RF.LOOP_CASES if RF.G.CASES[n]['script'] == 'inference' and not RF.G.CASES[n]['dae']['noise']])
def test_inference_dropin_vs_reference_run(cuda, tmp_path, name):
    """iterative_inference.inference() of this package against the reference's own run of iterative_inference.py:inference():
    the saved batch<i>.npz (Y_fcn, Y_ii) and the three print_results blocks of the summary."""
    import warnings
    from iterative_inference_segm_b200.iterative_inference import inference
    from iterative_inference_segm_b200.helpers import build_experiment_name
    fx, case = RF.load(name)
    _, blocks = RF.parse_stdout(str(fx['stdout']))
    d = dict(case['dae'], concat_h=list(case['dae']['concat_h']))
    segm_net = case.get('segm_net', 'fcn8')
    exp_name = build_experiment_name(segm_net, data_aug=False, ae_h=False, **dict(list(d.items()) + list(RF.G.TRAINING_DICT.items())))
    for precision, tol, agree, mtol in _REF_PRECISIONS:
        if 'sub' in case and precision == 'bf16':
            agree = 0.985          # full size, 50 iterations: the bf16 variant's measured 98.9 %
        if segm_net == 'densenet' and precision == 'bf16':
            tol, agree, mtol = 0.15, 0.97, 5e-2          # the bf16 DenseNet's own tolerance (tests/test_densenet_gpu.py)
        if case['dae']['kind'] == 'contextmod' and precision == 'bf16':
            continue          # the context module is fp32 either way
        root = tmp_path / precision
        root.mkdir()
        wpath, lpath, spath = _reference_layout(root, case, exp_name)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = inference('camvid', segm_net, learn_step=case['step'], num_iter=case['num_iter'],
                            dae_dict_updates=dict(case['dae'], concat_h=list(case['dae']['concat_h'])), training_dict=dict(RF.G.TRAINING_DICT),
                            data_augmentation=False, which_set='test', ae_h=False, savepath=spath, loadpath=lpath, weights_path=wpath,
                            data_iter=RF.G.SyntheticCamvidIterator(case), save_batches=True, verbose=False, precision=precision)
        worst = {'Y_fcn': 0.0, 'Y_ii': 0.0}
        for i in range(case['nbatches']):
            with np.load(os.path.join(out['savepath'], 'batch%d.npz' % i)) as f:
                k = case.get('sub', 1)      # ref_full_size keeps every k-th pixel of the probabilities and all argmax labels
                for key in ('Y_fcn', 'Y_ii'):
                    ref = fx['%s_%d' % (key, i)]
                    got = f[key][:, :, ::k, ::k]
                    err = float(np.abs(got - ref).max())
                    worst[key] = max(worst[key], err)
                    assert got.shape == ref.shape and err < tol, (name, precision, key, i, err)
                    lab_ref = fx['labels_%s_%d' % (key[2:], i)] if k > 1 else ref.argmax(1)
                    assert float((f[key].argmax(1) == lab_ref).mean()) >= agree, (name, precision, key, i)
        # the summary blocks the reference printed last: FCN, FCN+DAE, ITERATIVE INFERENCE (loss, accuracy, mean Jaccard)
        for (title, loss_r, acc_r, jacc_r), key in zip(blocks[-3:], ('fcn', 'fcn_dae', 'iterative')):
            loss, acc, jacc = [float(v) for v in out[key]]
            assert abs(loss - loss_r) < mtol and abs(acc - acc_r) < mtol and abs(jacc - jacc_r) < mtol, (name, precision, title, out[key], (loss_r, acc_r, jacc_r))
        print('%s %s: max-abs vs the reference run: Y_fcn %.2e  Y_ii %.2e' % (name, precision, worst['Y_fcn'], worst['Y_ii']))


@pytest.mark.parametrize('name', [n for n in RF.LOOP_CASES if RF.G.CASES[n]['script'] == 'valid'])
def test_valid_dropin_vs_reference_run(cuda, tmp_path, name):
    """iterative_inference_valid.inference() against the reference's own run: the saved iterations<step>.npz (valid_mat, the
    per-iteration Jaccard numerators / denominators summed over images -- zero columns where every image had converged) and
    the returned per-iteration mean Jaccard."""
    import warnings
    from iterative_inference_segm_b200.iterative_inference_valid import inference
    from iterative_inference_segm_b200.helpers import build_experiment_name
    fx, case = RF.load(name)
    d = dict(case['dae'], concat_h=list(case['dae']['concat_h']))
    exp_name = build_experiment_name('fcn8', data_aug=False, ae_h=False, **dict(list(d.items()) + list(RF.G.TRAINING_DICT.items())))
    wpath, lpath, spath = _reference_layout(tmp_path, case, exp_name)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = inference('camvid', 'fcn8', learn_step=case['step'], num_iter=case['num_iter'], dae_dict_updates=d,
                        training_dict=dict(RF.G.TRAINING_DICT), which_set='test', savepath=spath, loadpath=lpath,
                        weights_path=wpath, data_iter=RF.G.SyntheticCamvidIterator(case), verbose=False, precision='mixed')
    with np.load(os.path.join(spath, 'camvid', exp_name, 'img_plots', str(case['step']), 'test', 'iterations%s.npz' % str(case['step']))) as f:
        vm = f['arr_0']
    ref = fx['valid_mat']
    assert vm.shape == ref.shape
    assert np.array_equal(vm.sum(axis=(0, 1)) == 0, ref.sum(axis=(0, 1)) == 0), 'iterations reached (early exit) differ'
    n_pix = case['B'] * case['nbatches'] * case['H'] * case['W']
    assert float(np.abs(vm - ref).max()) <= 2e-3 * n_pix, float(np.abs(vm - ref).max())          # a handful of argmax flips at most
    assert np.allclose(res, fx['res'], atol=1e-3, equal_nan=True), (res, fx['res'])


def test_noised_mask_subgraphs_vs_reference_run(cuda):
    """noise > 0 at inference (the valid script's default is 0.5): in the reference every DePool2D's mask sub-graph is noised
    with ITS OWN draw (layers/mylayers.py:91-93; one independent random stream per symbolic call).  tests/golden/ref_noise.npz
    is the reference's own run with the stand-in's logged, reproducible draws; feeding the same numbers to
    `buildDAE(..., stochastic_masks=True)` -- level p's mask from a pass on y + 0.5 * N_p -- reproduces its loop at the fp32 bar."""
    from tests.test_oracle import fx_noise_rows
    from iterative_inference_segm_b200.models.DAE_h import buildDAE
    from iterative_inference_segm_b200 import _kernels as K
    G = RF.G
    fx, case = RF.load('ref_noise')
    pf = weights.synthetic_fcn8_params(3, NCLS, **G.FCN8_WEIGHTS)
    pd = G.case_dae_params(case)
    dae = buildDAE([None], None, NCLS, nb_features_to_concat=512, padding=100, concat_h=['pool4'], noise=case['dae']['noise'],
                   n_filters=64, conv_before_pool=1, additional_pool=2, skip=True, unpool_type='trackind', params=pd, precision='mixed',
                   stochastic_masks=True)
    net = dae.net
    X, _ = G.case_batch(case, 0)
    h, y0 = nets.fcn8_forward(pf, torch.from_numpy(X), NCLS)
    rows = list(fx_noise_rows(fx))
    c, outs, shared = 1, [], []          # row 0 is the batch-of-2 pred_dae_fn call; then de_fn per image and iteration
    for im in range(case['B']):
        hd = K.pack_nchw(h[im:im + 1].to(cuda), net.h_pad, split=True)
        y = y0[im:im + 1].to(cuda)
        ys = y.clone()
        for it in range(case['num_iter']):
            noise = [n.to(cuda) for n in rows[c](y.shape)]
            c += 1
            for yy, nz, dst in ((y, noise, 'y'), (ys, noise[0], 'ys')):          # per-DePool2D draws / one shared draw
                logits = net.logits(hd, K.pack_nchw(yy, net.y_cpad, split=True), y_f32=yy, noise=nz)
                p = torch.empty_like(yy)
                K.softmax_nchw(logits, NCLS, p)
                if dst == 'y':
                    y = torch.clamp(y - case['step'] * (y - p), 0.0, 1.0)
                else:
                    ys = torch.clamp(ys - case['step'] * (ys - p), 0.0, 1.0)
        outs.append(y.cpu())
        shared.append(ys.cpu())
    err = float((torch.cat(outs).numpy() - fx['Y_ii_0']).__abs__().max())
    err_shared = float((torch.cat(shared).numpy() - fx['Y_ii_0']).__abs__().max())
    print('noise 0.5, per-DePool2D draws: y max-abs vs the reference run %.2e (one shared draw: %.2e)' % (err, err_shared))
    assert err < TOL_F32, err
    assert err_shared > 3 * err          # the joint distribution matters: a shared draw is not what the reference computes
