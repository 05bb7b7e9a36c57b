"""The DAE training step of train_dae.py restated on torch CPU with autograd (test oracle).

Follows train_dae.py:243-335 (noise -> DAE forward without `deterministic`, masked crossentropy +
lmb * masked squared_error, lasagne.updates.rmsprop) and metrics.py:68-91,144-156, for the benchmark
configuration (kind='standard', unpool_type='trackind', skip=True, bn=0, dropout=0).  The parts of
Theano/Lasagne semantics that autograd's defaults would get wrong are custom functions:

  * Pool2DLayer backward = Theano CPU MaxPoolGrad: EVERY element equal to the window max receives
    the upstream gradient (torch routes it to one element);
  * DePool2D (layers/mylayers.py:88-115): out = repeat2(u) * mask, the mask being constant w.r.t.
    the parameters (T.grad of the pool w.r.t. its input, all-ones upstream); with noise > 0 the mask
    sub-graph is a SEPARATE stochastic forward of the contracting path (lasagne.layers.get_output is
    called without `deterministic`, layers/mylayers.py:91-93), so masks come from a second noise draw;
  * rectify = T.nnet.relu = 0.5 * (x + |x|): its gradient at exactly 0 is 0.5.

Noise tensors are explicit inputs (Theano's MRG31k3p stream cannot be reproduced): `noise_main`
for the path that produces values, `noise_mask` for the DePool2D mask sub-graph.
"""
import torch
import torch.nn.functional as F

from . import lasagne_semantics as L
from .nets import dae_levels

EPS_CE = 1e-7          # metrics.py:8 _EPSILON


class _Relu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * (0.5 * (1.0 + torch.sign(x)))       # d/dx 0.5*(x+|x|): 1, 0.5 at 0, 0


class _MaxPool2TieAll(torch.autograd.Function):
    """Pool2DLayer(2), ignore_border=True; backward to every tied maximum (Theano CPU MaxPoolGrad)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return F.max_pool2d(x, 2, 2)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        mask = L.tie_mask(x)
        up = torch.zeros_like(x)
        r = g.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        up[:, :, :r.shape[2], :r.shape[3]] = r
        return up * mask


def _q_bf16(x):
    """Round to bf16 with a straight-through gradient: emulates the storage format of the B200 path
    (weights, activations) so that pool ties -- a discontinuous function of the values -- fall the same way."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def dae_forward_train(params, y_noisy, h, padding, mask_source_y=None, concat_h=('pool4',), additional_pool=2,
                      emulate_bf16=False, tap=None, ae_out=None):
    """Differentiable DAE forward -> logits (before the softmax), cropped to the input size.
    `ae_out` (dict): receives the two layers the `ae_h` loss term compares (train_dae.py:238-239,317-319): 'h' = the layer named
    'h_to_recon', the DAE's own pool_{n_pool} (models/fcn_down.py:117-122), and 'h_hat' = the layer named 'h_hat',
    fused_up_{n_pool+1} = up_conv_{n_pool+1} + pool_{n_pool} (models/fcn_up.py:145-146).
    `mask_source_y`: input of the separate contracting-path forward the DePool2D masks are taken from
    (None: the same forward).  Masks are constants (detached).  `emulate_bf16`: weights, inputs and every
    stored activation are rounded to bf16 (straight-through), as on the B200 path; the arithmetic stays fp32.
    `tap` (dict): receives the discrete decisions of this pass -- 'masksA' (tie masks of the main pass: the pool
    backward routing), 'zero' (pre-rectifier value exactly 0: rectify'(0) = 0.5) and 'masksB' (the DePool2D masks) --
    so that a test can teacher-force them into the CUDA path and compare the arithmetic alone."""
    n_pool, total = dae_levels(concat_h, additional_pool)
    q = _q_bf16 if emulate_bf16 else (lambda t: t)
    Wd = [(q(params[2 * i]), params[2 * i + 1]) for i in range(total)]
    Wu = [(q(params[2 * (total + i)]), params[2 * (total + i) + 1]) for i in range(total)]
    y_noisy, h = q(y_noisy), q(h)
    if isinstance(mask_source_y, (list, tuple)):
        mask_source_y = [q(m) for m in mask_source_y]
    elif mask_source_y is not None:
        mask_source_y = q(mask_source_y)

    zero = []

    def down(x, differentiable, upto=total):
        pre, pools = [], []
        for p in range(upto):
            pad = padding if (p == 0 and padding > 0) else 1
            a = q(F.conv2d(x, Wd[p][0], Wd[p][1], padding=pad))
            if differentiable:
                zero.append((a.detach() == 0).to(a.dtype))
                if tap is not None and a.requires_grad:
                    a.retain_grad()
                    tap.setdefault('pre_act', []).append(a)
            r = _Relu.apply(a) if differentiable else torch.relu(a)
            pre.append(r)
            x = _MaxPool2TieAll.apply(r) if differentiable else F.max_pool2d(r, 2, 2)
            pools.append(x)
            if p + 1 == n_pool:
                x = torch.cat([h, x], dim=1)
        return pre, pools

    pre, pools = down(y_noisy, True)
    if mask_source_y is None:
        masks = [L.tie_mask(r.detach()) for r in pre]
    elif isinstance(mask_source_y, (list, tuple)):
        # the reference's graph: an independent noise draw per DePool2D, level p's mask from its own pass over levels 1..p
        assert len(mask_source_y) == total
        with torch.no_grad():
            masks = [L.tie_mask(down(mask_source_y[p], False, upto=p + 1)[0][p]) for p in range(total)]
    else:
        with torch.no_grad():
            pre_m, _ = down(mask_source_y, False)
        masks = [L.tie_mask(r) for r in pre_m]
    if tap is not None:
        tap['masksA'] = [L.tie_mask(r.detach()) for r in pre]
        tap['zero'] = zero
        tap['masksB'] = masks
    u = pools[-1]
    for i, p in enumerate(range(total, 0, -1)):
        m = masks[p - 1]
        up = torch.zeros_like(m)
        r = u.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        up[:, :, :r.shape[2], :r.shape[3]] = r
        v = up * m
        c = F.conv2d(v, Wu[i][0], Wu[i][1], padding=1)
        if p > 1:
            a, b = L.center_crop_pair(c, pools[p - 2])
            if ae_out is not None and p == n_pool + 1:
                ae_out['h'], ae_out['h_hat'] = pools[n_pool - 1], a + b
            u = q(a + b)
        else:
            u = L.center_crop_to(c, y_noisy.shape[2], y_noisy.shape[3])
    return u


def dice_loss(p, target, void_labels, class_for_dice=1):
    """metrics.py:93-113: -(2 sum(t p) + 1) / (sum t + sum p + 1) over the flattened channel `class_for_dice` of the softmax
    output p and of the one-hot target (cast to int32); entries whose TARGET VALUE equals a void label are dropped first
    (a comparison of 0/1 values with the label id, as the reference writes it -- not of the pixel's class)."""
    t = target[:, class_for_dice].reshape(-1).to(torch.int32)
    q = p[:, class_for_dice].reshape(-1)
    for v in void_labels:
        keep = t != int(v)
        q, t = q[keep], t[keep]
    inter = (t * q).sum()
    return -(2.0 * inter + 1) / (t.sum() + q.sum() + 1)


def ae_h_loss(ae):
    """train_dae.py:317-319: `squared_error_L(h, h_hat).mean()`.  NB h_hat = up_conv + h (skip sum), so the term is the mean square
    of up_conv_{n_pool+1}'s output up to fp32 rounding, and its gradient with respect to h cancels exactly (2 (h - h_hat) / n from
    `h`, the negative of it through the sum).  `freezeParameters(net['pool' + str(n_pool)])` (models/fcn_down.py:80-81, single=True)
    touches the pooling layer only, which has no parameters: nothing is frozen."""
    return ((ae['h'] - ae['h_hat']) ** 2).mean()


def loss_fn(logits, target, n_classes, lmb=1.0, use_ce=True, use_mse=True, use_dice=False):
    """crossentropy (metrics.py:68-91, one_hot=True, void = n_classes) [+ dice_loss (metrics.py:93-113, void_labels =
    [n_classes])] + lmb * squared_error (metrics.py:144-156, void int) of softmax(logits) against the one-hot target
    (B, C+1, H, W), in the order train_dae.py:278-294 adds them."""
    p = torch.softmax(logits, dim=1)
    true = target.argmax(dim=1)
    mask = (true != n_classes).to(p.dtype)
    loss = 0.0
    if use_ce:
        pc = torch.clamp(p, EPS_CE, 1.0 - EPS_CE)
        idx = (true * mask.long()).unsqueeze(1)                       # void pixels point at class 0, then get masked
        ce = -torch.log(pc.gather(1, idx)).squeeze(1)
        loss = loss + (ce * mask).sum() / mask.sum()
    if use_dice:
        loss = loss + dice_loss(p, target, [n_classes])
    if use_mse:
        t = target[:, :n_classes]
        se = ((p - t) ** 2).mean(dim=1)
        m2 = t.sum(dim=1)
        loss = loss + lmb * (se * m2).sum() / m2.sum()
    return loss


def adam_update(params, moms, vels, grads, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """lasagne.updates.adam with its defaults (train_dae.py:328-329): t <- t_prev + 1 (float32);
    a_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); m <- beta1 m + (1-beta1) g; v <- beta2 v + (1-beta2) g^2;
    p <- p - a_t m / (sqrt(v) + eps).  Returns (new_params, new_moms, new_vels, t)."""
    t = torch.tensor(float(t), dtype=torch.float32) + 1.0
    a_t = lr * torch.sqrt(1.0 - torch.tensor(beta2, dtype=torch.float32) ** t) / (1.0 - torch.tensor(beta1, dtype=torch.float32) ** t)
    new_p, new_m, new_v = [], [], []
    for p, m, v, g in zip(params, moms, vels, grads):
        m2 = beta1 * m + (1 - beta1) * g
        v2 = beta2 * v + (1 - beta2) * g * g
        new_m.append(m2)
        new_v.append(v2)
        new_p.append(p - a_t * m2 / (torch.sqrt(v2) + eps))
    return new_p, new_m, new_v, float(t)


def train_step(params, accus, y, h, target, n_classes, padding, lr, noise_main=None, noise_mask=None, lmb=1.0,
               rho=0.9, eps=1e-6, emulate_bf16=False, tap=None, loss_terms=None, ae_h=False, **dae_kw):
    """One train_fn call (train_dae.py:334-335): returns (loss, grads, new_params, new_accus).
    lasagne.updates.rmsprop: a <- rho*a + (1-rho)*g^2 ; p <- p - lr * g / sqrt(a + eps)."""
    ps = [p.clone().requires_grad_(True) for p in params]
    y_main = y if noise_main is None else y + noise_main
    if isinstance(noise_mask, (list, tuple)) or (noise_mask is not None and noise_mask.dim() == 5):
        y_mask = [y + n for n in noise_mask]          # one draw per DePool2D (level 1 first)
    else:
        y_mask = None if noise_mask is None else y + noise_mask
    ae = {} if ae_h else None
    logits = dae_forward_train(ps, y_main, h, padding, mask_source_y=y_mask, emulate_bf16=emulate_bf16, tap=tap, ae_out=ae, **dae_kw)
    loss = loss_fn(logits, target, n_classes, lmb=lmb, **(loss_terms or {}))
    if ae_h:
        loss = loss + ae_h_loss(ae)
    grads = torch.autograd.grad(loss, ps)
    new_p, new_a = [], []
    for p, a, g in zip(params, accus, grads):
        a2 = rho * a + (1 - rho) * g * g
        new_a.append(a2)
        new_p.append(p - lr * g / torch.sqrt(a2 + eps))
    return float(loss.detach()), [g.detach() for g in grads], new_p, new_a
