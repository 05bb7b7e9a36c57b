"""Pins the CPU oracle: the reference ships no tests or fixtures ("parity unpinned"), so the
oracle is checked against what CAN be derived from the reference code by hand -- the shape
tables, the Lasagne layer identities, hand-computed metric examples -- against an fp64 re-run
of itself, against the committed golden vectors (tests/golden/, made by make_golden.py) -- and, since round 2, against
OUTPUTS OF THE REFERENCE ITSELF: tests/golden/ref_*.npz hold what the reference's own drivers, model builders and metrics
computed when executed in the build container through oracle/refrun (tests/golden/make_reference_golden.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lasagne_semantics as L, loop, metrics as M, nets, weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


# ---- shapes (derived from models/fcn_down.py, fcn_up.py, fcn8.py) -------
def test_dae_param_shapes_benchmark_config():
    sh = nets.dae_param_shapes(11, 512, n_filters=64, concat_h=('pool4',), additional_pool=2)
    names = [s[0] for s in sh]
    assert names == ['conv%d_1' % i for i in range(1, 7)] + ['up_conv%d' % i for i in range(6, 0, -1)]
    W = {s[0]: s[1] for s in sh}
    assert W['conv1_1'] == (64, 11, 3, 3)
    assert W['conv5_1'] == (1024, 1024, 3, 3)      # (512 h + 512) -> 1024
    assert W['conv6_1'] == (2048, 1024, 3, 3)
    assert W['up_conv6'] == (1024, 2048, 3, 3)
    assert W['up_conv5'] == (512, 1024, 3, 3)      # skip-sum partner is the UN-concatenated pool4
    assert W['up_conv1'] == (11, 64, 3, 3)
    n_params = sum(int(np.prod(s[1])) + s[2][0] for s in sh)
    assert abs(n_params - 55.0e6) < 0.1e6           # "DAE parameters: 55.0 M"


def test_fcn8_param_shapes():
    sh = nets.fcn8_param_shapes(3, 11)
    assert len(sh) == 21                            # 42 arrays
    assert sh[13][0] == 'fc6' and sh[13][1] == (4096, 512, 7, 7)
    assert sh[-1][0] == 'upsample' and sh[-1][1] == (11, 11, 16, 16)


def test_fcn8_shapes_small():
    """pad=100 geometry: pool4 of a HxW image is ((H+198)//16, (W+198)//16); probs has the input size."""
    X, _, _ = weights.synthetic_batch(1, 32, 48)
    p = weights.synthetic_fcn8_params(3, 11)
    h, y0, s2, s4, up = nets.fcn8_forward(p, X, 11, layer=('pool4', 'probs_dimshuffle', 'score2', 'score4', 'upsample'))
    assert h.shape == (1, 512, (32 + 198) // 16, (48 + 198) // 16)
    assert y0.shape == (1, 11, 32, 48)
    assert torch.allclose(y0.sum(1), torch.ones(1, 32, 48), atol=1e-5)


@pytest.mark.parametrize('H,W', [(360, 480), (224, 224)])
def test_fcn8_geometry_table(H, W):
    """SURVEY.md App. B.2 by arithmetic only (no forward pass)."""
    s = [H + 198, W + 198]
    sizes = []
    for _ in range(5):
        s = [s[0] // 2, s[1] // 2]
        sizes.append(tuple(s))
    if (H, W) == (360, 480):
        assert sizes[3] == (34, 42) and sizes[4] == (17, 21)
        fc6 = (sizes[4][0] - 6, sizes[4][1] - 6)
        assert fc6 == (11, 15)
        score2 = ((fc6[0] - 1) * 2 + 4, (fc6[1] - 1) * 2 + 4)
        assert score2 == (24, 32) and (sizes[3][0] - 24) // 2 == 5
        score4 = ((score2[0] - 1) * 2 + 4, (score2[1] - 1) * 2 + 4)
        assert score4 == (50, 66) and (sizes[2][0] - 50) // 2 == 9
        up = ((score4[0] - 1) * 8 + 16, (score4[1] - 1) * 8 + 16)
        assert up == (408, 536) and ((up[0] - H) // 2, (up[1] - W) // 2) == (24, 28)
    else:
        assert sizes[3] == (26, 26)


def test_dae_level_sizes_360x480():
    s = (360 + 198, 480 + 198)
    expect = [(558, 678), (279, 339), (139, 169), (69, 84), (34, 42), (17, 21)]
    for e in expect:
        assert s == e
        s = (s[0] // 2, s[1] // 2)
    assert s == (8, 10)
    assert ((558 - 360) // 2, (678 - 480) // 2) == (99, 99)


# ---- Lasagne identities ------------------------------------------------------------------
def test_tie_inclusive_mask_and_depool():
    x = torch.tensor([[[[1., 1., 0.], [0., 1., 5.], [2., 3., 4.]]]])      # 3x3: trailing row/col dropped
    m = L.tie_mask(x)
    assert m.tolist() == [[[[1., 1., 0.], [0., 1., 0.], [0., 0., 0.]]]]   # all three 1s tie
    u = torch.tensor([[[[7.]]]])
    assert L.depool2d(u, x).tolist() == [[[[7., 7., 0.], [0., 7., 0.], [0., 0., 0.]]]]
    # torch's own unpool routes to ONE element only -> must not be used for the reference semantics
    _, idx = F.max_pool2d(x, 2, 2, return_indices=True)
    first_only = F.max_unpool2d(u, idx, 2, 2, output_size=(3, 3))
    assert float(first_only.sum()) == 7.0


def test_tie_mask_equals_theano_maxpoolgrad_rule():
    """MaxPoolGrad on CPU: for each window element, gx += gz if x == max."""
    torch.manual_seed(0)
    x = torch.relu(torch.randn(2, 3, 7, 9)).round()            # many exact ties and zeros
    m = L.tie_mask(x)
    ref = torch.zeros_like(x)
    for i in range(3):
        for j in range(4):
            win = x[:, :, 2 * i:2 * i + 2, 2 * j:2 * j + 2]
            mx = win.amax((2, 3), keepdim=True)
            ref[:, :, 2 * i:2 * i + 2, 2 * j:2 * j + 2] = (win == mx).float()
    assert torch.equal(m, ref)


def test_deconv_is_flipped_conv_transpose():
    """Deconv2DLayer(flip_filters=False) = input-gradient of a TRUE convolution."""
    torch.manual_seed(1)
    x = torch.randn(1, 3, 5, 6)
    W = torch.randn(3, 4, 4, 4)     # (in, out, k, k)
    b = torch.randn(4)
    y = L.deconv2d(x, W, b, 2)
    # gradient of sum(conv_true(z, W') * x) wrt z, conv_true = correlation with flipped kernel
    z = torch.zeros(1, 4, y.shape[2], y.shape[3], requires_grad=True)
    out = F.conv2d(z, W.flip(2, 3), stride=2)       # true convolution of z with W (in=4 -> out=3)
    (out * x).sum().backward()
    assert torch.allclose(y - b.view(1, -1, 1, 1), z.grad, atol=1e-5)


def test_center_crop_offsets():
    a = torch.arange(7 * 9.).view(1, 1, 7, 9)
    c = L.center_crop_to(a, 4, 4)
    assert c[0, 0, 0, 0] == a[0, 0, 1, 2]       # offsets (7-4)//2 = 1, (9-4)//2 = 2


# ---- metrics: hand-computed examples (metrics.py) ----------------------------------------------
def _onehot(lab, C):
    return np.eye(C, dtype=np.float32)[lab].transpose(0, 3, 1, 2)


def test_metrics_hand_example():
    # 1 image, 2x2 pixels, 3 classes + void(3).  truth: [[0,1],[2,void]]  pred: [[0,2],[2,1]]
    lab = np.array([[[0, 1], [2, 3]]])
    t = _onehot(lab, 4)
    y = np.zeros((1, 3, 2, 2), np.float32)
    for (i, j), c in {(0, 0): 0, (0, 1): 2, (1, 0): 2, (1, 1): 1}.items():
        y[0, :, i, j] = 0.1
        y[0, c, i, j] = 0.8
    cm = M.confusion_matrix(y, t, 3)
    assert cm.tolist() == [[1, 0, 0], [0, 0, 0], [0, 1, 1]]     # rows = prediction; void pixel dropped
    jac = M.jaccard(y, t, 3)
    assert jac.dtype == np.float32
    assert jac.tolist() == [[1, 0, 1], [1, 1, 2]]               # [TP; TP+FP+FN]
    assert M.accuracy(y, t, [3]) == np.float32(2. / 3.)
    # squared error: mean over 3 channels, masked by non-void, over 3 non-void pixels
    per_pix = [((0.8 - 1) ** 2 + 2 * 0.1 ** 2) / 3, (0.1 ** 2 + (0.1 - 1) ** 2 + 0.8 ** 2) / 3,
               ((0.8 - 1) ** 2 + 2 * 0.1 ** 2) / 3]
    assert abs(float(M.squared_error(y, t, 3)) - sum(per_pix) / 3) < 1e-6


def test_argmax_tie_is_first_index():
    y = np.full((1, 3, 1, 1), 1 / 3., np.float32)
    t = _onehot(np.array([[[0]]]), 4)
    assert M.confusion_matrix(y, t, 3)[0, 0] == 1


# ---- the loop (iterative_inference.py:258-291) ---------------------------------------------------
def _tiny_dae(seed=3):
    """A 3-level DAE (concat at pool1, additional_pool=2) small enough for CPU unit tests."""
    gen = torch.Generator().manual_seed(seed)
    sh = nets.dae_param_shapes(4, 8, n_filters=8, concat_h=('pool1',), additional_pool=2)
    params = []
    for name, ws, bs in sh:
        params += [weights.glorot_uniform(ws, gen), torch.zeros(bs)]
    return params


def test_loop_update_order_and_early_exit():
    params = _tiny_dae()
    torch.manual_seed(0)
    y0 = torch.softmax(torch.randn(1, 4, 12, 14), 1)
    h = torch.relu(torch.randn(1, 8, 6, 7))
    kw = dict(concat_h=('pool1',), additional_pool=2)
    y, n_exec, per_iter, tr = loop.iterate_image(params, h, y0, 0.5, 5, 0, eps=0.0, record=True, **kw)
    assert n_exec == 5
    p0 = nets.dae_forward(params, y0, h, 0, **kw)
    y1 = torch.clamp(y0 - 0.5 * (y0 - p0), 0, 1)
    assert torch.allclose(tr[0]['y'], y1)
    # huge eps: exactly one update happens, then break BEFORE the metrics of that iteration
    L_ = torch.zeros(1, 5, 12, 14); L_[:, 0] = 1
    y, n_exec, per_iter, _ = loop.iterate_image(params, h, y0, 0.5, 5, 0, eps=1e9, t_im=L_.numpy(), n_classes=4,
                                                void_labels=[4], **kw)
    assert n_exec == 1 and per_iter == [] and torch.allclose(y, y1)


def test_fp32_vs_fp64_single_application():
    params = _tiny_dae()
    torch.manual_seed(1)
    y0 = torch.softmax(torch.randn(1, 4, 12, 14), 1)
    h = torch.relu(torch.randn(1, 8, 6, 7))
    kw = dict(concat_h=('pool1',), additional_pool=2)
    p32 = nets.dae_forward(params, y0, h, 0, **kw)
    p64 = nets.dae_forward([p.double() for p in params], y0.double(), h.double(), 0, **kw)
    assert float((p32.double() - p64).abs().max()) < 1e-5


# ---- golden vectors ----------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['dae_32x40', 'fcn8_32x40', 'loop_32x40'])
def test_golden(name):
    from tests.golden import make_golden
    g = np.load(os.path.join(GOLD, name + '.npz'))
    fresh = make_golden.CASES[name]()
    for k in g.files:
        a, b = g[k], fresh[k]
        assert a.shape == b.shape, k
        if a.dtype.kind in 'iu':
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, atol=5e-5, rtol=1e-4), (k, float(np.abs(a - b).max()))


def test_dilated_conv_is_lasagne_dilatedconv2dlayer():
    """DilatedConv2DLayer (models/contextmod_dae.py:76-103): W is (in, out, kh, kw), unflipped, 'valid';
    out[n, f, i, j] = b[f] + sum_{c, r, s} W[c, f, r, s] x[n, c, i + r*d, j + s*d] -- written out as loops."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 9, 11, generator=g)
    W = torch.randn(3, 4, 3, 3, generator=g)
    b = torch.randn(4, generator=g)
    for d in (1, 2, 3):
        out = L.dilated_conv2d(x, W, b, d, relu=False)
        assert tuple(out.shape) == (2, 4, 9 - 2 * d, 11 - 2 * d)
        ref = torch.zeros_like(out)
        for f in range(4):
            for c in range(3):
                for r in range(3):
                    for s in range(3):
                        ref[:, f] += W[c, f, r, s] * x[:, c, r * d:r * d + out.shape[2], s * d:s * d + out.shape[3]]
            ref[:, f] += b[f]
        assert torch.allclose(out, ref, atol=1e-5)
    assert float(L.dilated_conv2d(x, W, b, 2, relu=True).min()) >= 0.0


def test_contextmod_with_identity_init_returns_its_first_conv():
    """The reference initialises dilconv1..7 with IdentityInit (models/contextmod_dae.py:61-71, centre tap = identity, zero
    bias): every dilated conv then returns the centre crop of its (non-negative) input, PadLayer(32) is exactly the total
    shrink 2 * (1 + 2 + 4 + 8 + 16 + 1) = 64, and the module's logits are relu(conv1([h | y])) at the input size."""
    C, nb_h = 5, 3
    shapes = nets.contextmod_param_shapes(C, nb_h)
    assert [s[0] for s in shapes] == ['conv1'] + ['dilconv%d' % i for i in range(1, 8)]
    assert shapes[0][1] == (C, nb_h + C, 3, 3) and shapes[1][1] == (C, C, 3, 3) and shapes[7][1] == (C, C, 1, 1)
    g = torch.Generator().manual_seed(1)
    params = [torch.randn(shapes[0][1], generator=g) * 0.3, torch.randn(C, generator=g) * 0.1]
    for name, ws, bs in shapes[1:]:
        W = torch.zeros(ws)
        for i in range(C):
            W[i, i, ws[2] // 2, ws[3] // 2] = 1.0
        params += [W, torch.zeros(bs)]
    y = torch.softmax(torch.randn(2, C, 13, 17, generator=g), 1)
    h = torch.rand(2, nb_h, 13, 17, generator=g)
    logits = nets.contextmod_forward(params, y, h, return_logits=True)
    ref = torch.relu(F.conv2d(torch.cat([h, y], 1), params[0], params[1], padding=1))
    assert torch.equal(logits, ref)
    p = nets.contextmod_forward(params, y, h)
    assert torch.allclose(p.sum(1), torch.ones(2, 13, 17), atol=1e-6)


@pytest.mark.parametrize('concat', ['input', 'pool2'])
def test_fcn8_shaped_dae_concatenates_h_in_front(concat):
    """models/fcn8_dae.py:46-48,60-115 + models/model_helpers.py:91-93: h joins IN FRONT of the layer's channels, so the widened
    conv's W[:, :nb_h] acts on h.  With those weights zeroed the DAE equals the plain FCN8 on y; with the y part zeroed its
    output does not depend on y's channels at that conv (checked through the parameter shapes)."""
    C = 4
    nb_h = 3 if concat == 'input' else 128
    shapes = nets.fcn8_param_shapes(C, C, concat=(concat, nb_h))
    widened = 'conv1_1' if concat == 'input' else 'conv3_1'
    base = dict((n, ws) for n, ws, _ in nets.fcn8_param_shapes(C, C))
    for n, ws, _ in shapes:
        assert ws[1] == base[n][1] + (nb_h if n == widened else 0), (n, ws)
    pdae = weights.synthetic_fcn8_params(C, C, seed=2, logit_gain=5.0, concat=(concat, nb_h))
    idx = [n for n, _, _ in shapes].index(widened)
    pdae[2 * idx][:, :nb_h] = 0.0
    plain = [p.clone() for p in pdae]
    plain[2 * idx] = pdae[2 * idx][:, nb_h:].clone()
    g = torch.Generator().manual_seed(3)
    y = torch.softmax(torch.randn(1, C, 32, 40, generator=g), 1)
    hs = (32, 40) if concat == 'input' else ((32 + 198) // 4, (40 + 198) // 4)
    h = torch.randn(1, nb_h, *hs, generator=g)
    a = nets.fcn8_dae_forward(pdae, y, h, C, concat_h=(concat,))
    b = nets.fcn8_forward(plain, y, C, layer=('probs_dimshuffle',))[0]
    assert torch.allclose(a, b, atol=1e-6)


# ---- the oracle against outputs of the reference's own code (tests/golden/ref_*.npz) ------------------------------------
from tests import reference_fixtures as RF  # noqa: E402

REF_TOL = 2e-6          # float32 CPU arithmetic on both sides, different conv algorithms (measured: <= 3e-7)


def fx_noise_rows(fx):
    """Per pred_dae_fn / de_fn call of the reference run, in order: shape -> [N(0,1) tensor of level 1, ..., level P]."""
    from oracle.refrun import py2import
    py2import.install_stubs()
    from theano.sandbox import rng_mrg
    for ks, nb in zip(fx['noise_k'], fx['noise_batch']):
        def make(shape, ks=ks, nb=nb):
            assert shape[0] == nb, (tuple(shape), nb)
            return [rng_mrg.draw(int(k), shape) for k in ks]
        yield make


def _oracle_replay(case):
    """The oracle's version of what the reference's driver did for `case`: per batch (Y_fcn, Y_ii, n_exec, batch metrics,
    FCN metrics, FCN+DAE metrics, per-image per-iteration metrics) and the accumulated valid_mat."""
    G = RF.G
    pf = weights.synthetic_fcn8_params(3, RF.NCLS, **G.FCN8_WEIGHTS)
    pd = G.case_dae_params(case)
    d, kw, forward = case['dae'], {}, None
    if d['kind'] == 'standard':
        kw = RF.dae_kwargs(case)
    elif d['kind'] == 'contextmod':
        forward = lambda y, h: nets.contextmod_forward(pd, y, h)          # noqa: E731
    else:
        forward = lambda y, h: nets.fcn8_dae_forward(pd, y, h, RF.NCLS, concat_h=tuple(d['concat_h']))          # noqa: E731
    out, valid_mat = [], np.zeros((2, RF.NCLS, case['num_iter']))
    densenet = case.get('segm_net') == 'densenet'
    padding = 0 if densenet else 100          # iterative_inference.py:140,145
    if d['kind'] == 'standard' and d['noise'] > 0:
        # the DePool2D mask sub-graphs are noised even at inference: one independent draw per DePool2D and per function call.
        # The fixture logged which draw of the stand-in's stream each one consumed; regenerate the same numbers.
        noise_rows = iter(fx_noise_rows(case['_fixture']))
        forward = lambda y, h: nets.dae_forward(pd, y, h, padding, mask_source_y=[y + d['noise'] * n for n in next(noise_rows)(y.shape)], **kw)          # noqa: E731
    if densenet:
        from oracle import densenet as OD
        pdn = G.case_densenet_params(case)
    for i in range(case['nbatches']):
        X, Lb = G.case_batch(case, i)
        Xt = torch.from_numpy(X)
        if densenet:
            h, y0 = OD.densenet_forward(pdn, Xt, RF.NCLS, layer=list(d['concat_h']))
        elif d['concat_h'][0] == 'input':          # layer=['input', ...]: h is the image itself (iterative_inference.py:139)
            h, y0 = Xt, nets.fcn8_forward(pf, Xt, RF.NCLS, layer=('probs_dimshuffle',))[0]
        else:
            h, y0 = nets.fcn8_forward(pf, Xt, RF.NCLS, layer=(d['concat_h'][0], 'probs_dimshuffle'))
        m_fcn = M.val_fn(y0.numpy(), Lb, RF.NCLS, [RF.NCLS])
        p = forward(y0, h) if forward else nets.dae_forward(pd, y0, h, padding, **kw)
        m_dae = M.val_fn(p.numpy(), Lb, RF.NCLS, [RF.NCLS])
        per_image, ys, n_exec = [], [], []
        for im in range(X.shape[0]):          # iterative_inference.py:258-284; valid_mat: iterative_inference_valid.py:288
            y, n, per_iter, _ = loop.iterate_image(pd, h[im:im + 1], y0[im:im + 1], case['step'], case['num_iter'], padding,
                                                   t_im=Lb[im:im + 1], n_classes=RF.NCLS, void_labels=[RF.NCLS], forward=forward, **kw)
            per_image.append(per_iter)
            ys.append(y)
            n_exec.append(n)
            for it, (_, jacc_iter, _) in enumerate(per_iter):
                valid_mat[:, :, it] += jacc_iter
        Y = torch.cat(ys, dim=0)
        bm = M.val_fn(Y.numpy(), Lb, RF.NCLS, [RF.NCLS])
        out.append(dict(Y_fcn=y0.numpy(), Y_ii=Y.numpy(), n_exec=n_exec, m_ii=bm, m_fcn=m_fcn, m_dae=m_dae, per_image=per_image))
    return out, valid_mat


@pytest.mark.parametrize('name', RF.LOOP_CASES)
def test_oracle_vs_reference_run(name):
    """The oracle restatement against what the reference's own code computed (iterative_inference.py:inference /
    iterative_inference_valid.py:inference executed through oracle/refrun): FCN8 probabilities, the loop's final y, the
    per-iteration and per-batch metrics it printed, valid_mat."""
    fx, case = RF.load(name)
    got, valid_mat = _oracle_replay(dict(case, _fixture=fx))
    per_iter_ref, blocks = RF.parse_stdout(str(fx['stdout']))
    rel = lambda a, b: abs(a - b) <= 2e-6 * max(1.0, abs(b)) or (np.isnan(a) and np.isnan(b))          # noqa: E731
    # accuracy / Jaccard are argmax counts: probabilities that agree to 2e-7 can still break one near-tie differently
    # (ref_bn: one pixel of 3330 in the plain DAE pass, 4.6e-6 on the mean Jaccard)
    cnt = lambda a, b: abs(a - b) <= 2e-5 or (np.isnan(a) and np.isnan(b))          # noqa: E731
    # per-image, per-iteration `rec acc jaccard` lines: same COUNT (the early exit) and same values
    flat = [pi for g in got for pi in g['per_image']]
    assert [len(p) for p in flat] == [len(p) for p in per_iter_ref], 'iterations that reached val_fn differ (early exit)'
    for po, pr in zip(flat, per_iter_ref):
        for (acc, jacc, mse), (rec_r, acc_r, jm_r) in zip(po, pr):
            with np.errstate(divide='ignore', invalid='ignore'):
                jm = float(np.nanmean(jacc[0] / jacc[1]))
            assert rel(float(mse), rec_r) and cnt(float(acc), acc_r) and cnt(jm, jm_r), ((mse, acc, jm), (rec_r, acc_r, jm_r))
    # print_results blocks: running totals / (i + 1) of FCN, FCN+DAE and (inference script) ITERATIVE INFERENCE
    tot = {k: [0.0, 0.0, np.zeros((2, RF.NCLS))] for k in ('m_fcn', 'm_dae', 'm_ii')}
    expect = []
    for i, g in enumerate(got):
        for key, title in (('m_fcn', 'FCN'), ('m_dae', 'FCN+DAE'), ('m_ii', 'ITERATIVE INFERENCE')):
            acc, jacc, mse = g[key]
            tot[key][0] += float(mse)
            tot[key][1] += float(acc)
            tot[key][2] = tot[key][2] + jacc
            if key != 'm_ii' or case['script'] == 'inference':
                expect.append((title,) + tuple(float(v) for v in M.print_results_values(tot[key][0], tot[key][1], tot[key][2], i + 1)))
    n_summary = 3 if case['script'] == 'inference' else 2
    assert len(blocks) == len(expect) + n_summary
    for (t_r, l_r, a_r, j_r), (t_o, l_o, a_o, j_o) in zip(blocks, expect):
        assert t_r == t_o and rel(l_o, l_r) and cnt(a_o, a_r) and cnt(j_o, j_r), ((t_r, l_r, a_r, j_r), (t_o, l_o, a_o, j_o))
    if case['script'] == 'inference':
        k = case.get('sub', 1)          # ref_full_size keeps every k-th pixel of the probabilities and all argmax labels
        for i, g in enumerate(got):
            assert float(np.abs(g['Y_fcn'][:, :, ::k, ::k] - fx['Y_fcn_%d' % i]).max()) < REF_TOL
            # 50 free-running iterations at 360x480: two float32 CPU evaluations with different summation orders drift apart
            # through pool ties (measured 9.9e-5 here; the oracle's own fp32-vs-fp64 drift is 1.2e-4, oracle/validate_recipe.py)
            assert float(np.abs(g['Y_ii'][:, :, ::k, ::k] - fx['Y_ii_%d' % i]).max()) < (REF_TOL if k == 1 else 5e-4)
            if k > 1:
                assert float((g['Y_fcn'].argmax(1) == fx['labels_fcn_%d' % i]).mean()) >= 0.9999
                assert float((g['Y_ii'].argmax(1) == fx['labels_ii_%d' % i]).mean()) >= 0.9995          # measured 0.99985 (26 pixels)
    else:
        assert np.array_equal(valid_mat, fx['valid_mat'])          # integer counts: exact
        with np.errstate(divide='ignore', invalid='ignore'):
            res = np.nanmean(valid_mat[0] / valid_mat[1], axis=0)
        assert np.allclose(res, fx['res'], rtol=1e-12, atol=0, equal_nan=True)


def test_oracle_temperature_vs_reference_run():
    """models/fcn8.py:193-198 executed by the reference's buildFCN8 with temperature = 2.5."""
    fx, case = RF.load('ref_temperature')
    pf = weights.synthetic_fcn8_params(3, RF.NCLS, **RF.G.FCN8_WEIGHTS)
    X, _ = RF.G.case_batch(case, 0)
    h, y = nets.fcn8_forward(pf, torch.from_numpy(X), RF.NCLS, temperature=case['temperature'])
    assert float(np.abs(h.numpy() - fx['pool4']).max()) < REF_TOL * max(1.0, float(np.abs(fx['pool4']).max()))
    assert float(np.abs(y.numpy() - fx['Y_fcn']).max()) < REF_TOL
    y1 = nets.fcn8_forward(pf, torch.from_numpy(X), RF.NCLS)[1]
    assert float(np.abs(y1.numpy() - fx['Y_fcn']).max()) > 1e-2          # the temperature does something


@pytest.mark.parametrize('name', ['ref_train', 'ref_train_noise', 'ref_train_adam', 'ref_train_dice', 'ref_train_aeh'])
def test_oracle_train_step_vs_reference_run(name):
    """oracle/train.py against the reference's own train_dae.py:train() (two epochs of two rmsprop steps, the learning rate
    annealed in between, validation after each epoch; tests/golden/ref_train.npz, and ref_train_noise.npz with noise = 0.5 and
    the logged draws): per-epoch training / validation cost,
    validation Jaccard and squared error, and the parameters the reference saved after the fourth step (digest: 4096 strided
    samples per array + the sum / norm / max of each array's change)."""
    from oracle import train as T_
    G = RF.G
    fx, case = RF.load(name)
    sigma = case['dae']['noise']
    if sigma > 0:          # ref_train_noise: regenerate the numbers the reference consumed (main draw + one per DePool2D, also in validation)
        from oracle.refrun import py2import
        py2import.install_stubs()
        from theano.sandbox import rng_mrg
        train_k, val_k = iter(fx['train_k']), iter(fx['val_k'])
    pf = weights.synthetic_fcn8_params(3, RF.NCLS, **G.FCN8_WEIGHTS)
    init = G.case_dae_params(case)
    params, accus = [p.clone() for p in init], [torch.zeros_like(p) for p in init]
    moms, t_adam = [torch.zeros_like(p) for p in init], 0.0
    lr = np.float32(case['learning_rate'])
    tl = case['training_loss']          # ref_train_dice: crossentropy + dice_loss + squared_error (train_dae.py:278-294)
    terms = dict(use_ce='crossentropy' in tl, use_mse='squared_error' in tl, use_dice='dice' in tl)
    err_train, err_valid, jacc_val, mse_val = [], [], [], []
    for epoch in range(case['num_epochs']):
        tot = 0.0
        for i in range(case['nbatches']):                      # train_dae.py:356-383
            X, Lb = G.case_batch(case, i, 'train')
            h, y = nets.fcn8_forward(pf, torch.from_numpy(X), RF.NCLS)
            nkw = {}
            if sigma > 0:
                ks = next(train_k)
                nkw = dict(noise_main=sigma * rng_mrg.draw(int(ks[0]), y.shape), noise_mask=[sigma * rng_mrg.draw(int(k), y.shape) for k in ks[1:]])
            loss, grads, p_rms, a_rms = T_.train_step(params, accus, y, h, torch.from_numpy(Lb), RF.NCLS, 100, float(lr), lmb=case['lmb'], loss_terms=terms, ae_h=case.get('ae_h', False), **nkw)
            if case.get('optimizer') == 'adam':          # lasagne.updates.adam (train_dae.py:328-329)
                params, moms, accus, t_adam = T_.adam_update(params, moms, accus, grads, t_adam, float(lr))
            else:
                params, accus = p_rms, a_rms
            tot += loss
        err_train.append(tot / case['nbatches'])
        cost, jacc, mse = 0.0, 0.0, 0.0
        for i in range(case['val_nbatches']):                  # train_dae.py:387-411
            X, Lb = G.case_batch(case, i, 'val')
            h, y = nets.fcn8_forward(pf, torch.from_numpy(X), RF.NCLS)
            with torch.no_grad():
                # validation: deterministic=True switches the main noise off, but the DePool2D sub-graphs stay noised
                msk = [y + sigma * rng_mrg.draw(int(k), y.shape) for k in next(val_k)] if sigma > 0 else None
                ae = {} if case.get('ae_h') else None          # ref_train_aeh: + squared_error(h, h_hat).mean() (train_dae.py:317-319)
                logits = T_.dae_forward_train(params, y, h, 100, mask_source_y=msk, ae_out=ae)
                cost += float(T_.loss_fn(logits, torch.from_numpy(Lb), RF.NCLS, lmb=case['lmb'], **terms))
                if ae is not None:
                    cost += float(T_.ae_h_loss(ae))
                p = torch.softmax(logits, dim=1).numpy()
            jacc = jacc + M.jaccard(p, Lb, RF.NCLS)
            mse += float(M.squared_error(p, Lb, RF.NCLS))
        err_valid.append(cost / case['val_nbatches'])
        jacc_val.append(float(np.mean(jacc[0] / jacc[1])))
        mse_val.append(mse / case['val_nbatches'])
        lr = np.float32(float(lr) * case['lr_anneal'])          # train_dae.py:424: lr.set_value(float(lr.get_value() * lr_anneal)), a float32 shared variable
    assert np.allclose(err_train, fx['err_train'], rtol=2e-6, atol=0), (err_train, fx['err_train'])
    assert np.allclose(err_valid, fx['err_valid'], rtol=2e-6, atol=0), (err_valid, fx['err_valid'])
    assert np.allclose(mse_val, fx['mse_val'], rtol=2e-6, atol=0)
    assert np.allclose(jacc_val, fx['jacc_val'], rtol=0, atol=2e-5)
    assert str(fx['saved_as']) == ('dae_model_best.npz' if err_valid[1] < err_valid[0] else 'dae_model_last.npz')
    worst = 0.0
    for i, (p_new, p_old) in enumerate(zip(params, init)):
        dig = G.param_digest(i, p_new.numpy(), p_old.numpy())
        step = float(fx['p%d_delta' % i][2])                      # largest change of any element of this array over the 4 steps
        err = float(np.abs(dig['p%d_sample' % i] - fx['p%d_sample' % i]).max())
        worst = max(worst, err / step)
        # rmsprop divides by sqrt(accumulated g^2): an element whose gradient is ~0 amplifies rounding, hence a tolerance
        # relative to the array's largest update rather than per element
        assert err <= 2e-3 * step, (i, err, step)
        assert np.allclose(dig['p%d_delta' % i][1], fx['p%d_delta' % i][1], rtol=1e-3), (i, dig['p%d_delta' % i], fx['p%d_delta' % i])
    print('trained parameters: worst sample error / largest update = %.2e' % worst)


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='the reference tree only exists in the build container')
def test_reference_executed_again_reproduces_the_fixtures():
    """Where /root/reference is present, EXECUTE THE REFERENCE (its own inference() drivers, through oracle/refrun) again and
    require the committed fixtures bit for bit: the fixtures are what the reference computes, not hand-edited arrays.  Runs in a
    subprocess (the harness replaces modules such as data_loader); two small cases, ~40 s."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'tests.golden.make_reference_golden', '--check', 'ref_noskip', 'ref_noise'], cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('reproduces the committed fixture exactly') == 2
