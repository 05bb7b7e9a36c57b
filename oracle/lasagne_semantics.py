"""Lasagne / Theano layer semantics restated on torch CPU tensors (test oracle).

Each function names the Lasagne layer it restates and the reference call site.
Theano ddafc3e2 / Lasagne 45bb5689 are not in /root/reference; semantics are
restated from their published behaviour.
"""
import torch
import torch.nn.functional as F


def conv2d(x, W, b, pad, relu):
    """lasagne.layers.Conv2DLayer(..., flip_filters=False): cross-correlation,
    W (out,in,kh,kw), symmetric integer pad or 'same' (= k//2), default
    nonlinearity rectify.  Reference: models/fcn_down.py:102-104,
    models/fcn_up.py:84-86, models/fcn8.py:34-85."""
    if pad == 'same':
        pad = W.shape[2] // 2
    elif pad == 'valid':
        pad = 0
    y = F.conv2d(x, W, b, padding=pad)
    return torch.relu(y) if relu else y


def deconv2d(x, W, b, stride):
    """lasagne.layers.Deconv2DLayer(crop='valid', flip_filters=False,
    nonlinearity=linear): the input-gradient of a TRUE convolution, i.e.
    conv_transpose2d with spatially flipped kernels; W (in,out,kh,kw).
    Reference: models/fcn8.py:90-91,100-101,109-110."""
    return F.conv_transpose2d(x, W.flip(2, 3), b, stride=stride)


def maxpool2(x):
    """lasagne Pool2DLayer(incoming, 2): max, stride 2, ignore_border=True
    (floor).  Reference: models/fcn_down.py:122, models/fcn8.py:38."""
    return F.max_pool2d(x, 2, 2)


def tie_mask(x):
    """0/1 mask of every element equal to its 2x2 window max (all ties set);
    trailing odd row/col is zero.  This is what T.grad(pool, all-ones) yields
    with Theano's CPU MaxPoolGrad (`if x == max: gx += gz`).
    Reference: layers/mylayers.py:111-112."""
    B, C, H, W = x.shape
    h2, w2 = H // 2, W // 2
    p = F.max_pool2d(x, 2, 2)
    up = p.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    m = torch.zeros_like(x)
    m[:, :, :2 * h2, :2 * w2] = (x[:, :, :2 * h2, :2 * w2] == up).to(x.dtype)
    return m


def depool2d(u, x_prepool):
    """layers/mylayers.py:88-115 DePool2D: nearest x2 upsample of `u`,
    zero-pad bottom/right to the pre-pool size, multiply by the tie mask of
    the pre-pool activations."""
    B, C, H, W = x_prepool.shape
    up = u.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    out = torch.zeros_like(x_prepool)
    out[:, :, :up.shape[2], :up.shape[3]] = up
    return out * tie_mask(x_prepool)


def center_crop_to(x, H, W):
    """lasagne autocrop with cropping 'center': offset (size-min)//2.
    Reference: layers/mylayers.py:36-57, models/fcn_up.py:106-113."""
    oh = (x.shape[2] - H) // 2
    ow = (x.shape[3] - W) // 2
    return x[:, :, oh:oh + H, ow:ow + W]


def center_crop_pair(a, b):
    """autocrop of two inputs to the per-axis minimum (ElemwiseSumLayer with
    cropping=[None,None,'center','center'], models/fcn8.py:94-97)."""
    H = min(a.shape[2], b.shape[2])
    W = min(a.shape[3], b.shape[3])
    return center_crop_to(a, H, W), center_crop_to(b, H, W)


def channel_softmax(x):
    """dimshuffle(0,2,3,1) -> reshape(N,C) -> softmax rows -> back to NCHW.
    Reference: models/fcn_up.py:154-169, models/fcn8.py:120-191."""
    return torch.softmax(x, dim=1)


def dilated_conv2d(x, W, b, dilation, relu):
    """lasagne.layers.DilatedConv2DLayer(num_filters, filter_size, dilation, pad=0, flip_filters=False): W has shape
    (num_input_channels, num_filters, kh, kw) -- "first two sizes are swapped compared to a forward convolution" -- and
    the layer computes, via the backward-pass-wrt-weights op with subsample = dilation,
        out[n, f, i, j] = b[f] + sum_{c, r, s} W[c, f, r, s] * x[n, c, i + r*d, j + s*d]          ('valid', unflipped)
    so the output shrinks by (k - 1) * d.  lasagne is an un-vendored dependency of the reference (README.md:17-22, Lasagne
    0.2.dev1): restated from its published layer; used at models/contextmod_dae.py:76-103."""
    out = F.conv2d(x, W.permute(1, 0, 2, 3).contiguous(), b, padding=0, dilation=dilation)
    return torch.relu(out) if relu else out


def pad_layer(x, width):
    """lasagne.layers.PadLayer(width, val=0, batch_ndim=2): zero border on the two spatial axes (models/contextmod_dae.py:75)."""
    return F.pad(x, (width, width, width, width))
