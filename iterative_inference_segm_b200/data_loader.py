"""Synthetic stand-in for data_loader.load_data (data_loader.py:7-121).

The reference wraps the external, un-vendored `dataset_loaders` package (README.md:22).  Only the
iterator's attribute contract matters to the path (iterative_inference.py:117-125, 233-234), so
this module provides a CamVid-shaped synthetic iterator with the same attributes:
`.next()` -> (X float32 (B,3,H,W) in [0,1), L one-hot float32 (B,C+1,H,W)), `.nbatches`,
`.non_void_nclasses`, `.void_labels`, `.data_shape`, `.cmap`, `.mask_labels`."""
import numpy as np


class SyntheticSegmentationIterator(object):
    def __init__(self, n_images=10, batch_size=10, height=360, width=480, n_classes=11, seed=0,
                 first_image=0):
        self.non_void_nclasses = n_classes
        self.void_labels = [n_classes]
        self.data_shape = (3, height, width)
        self.batch_size = batch_size
        self.n_images = n_images
        self.nbatches = (n_images + batch_size - 1) // batch_size
        self.cmap = np.linspace(0, 1, (n_classes + 1) * 3).reshape(n_classes + 1, 3)
        self.mask_labels = ['class%d' % i for i in range(n_classes)] + ['void']
        self._seed, self._first, self._pos = seed, first_image, 0

    def _image(self, idx):
        rng = np.random.RandomState(self._seed * 1000003 + idx)
        C, H, W = self.data_shape
        X = rng.rand(C, H, W).astype(np.float32)
        lab = rng.randint(0, self.non_void_nclasses + 1, size=(H, W))
        L = np.eye(self.non_void_nclasses + 1, dtype=np.float32)[lab].transpose(2, 0, 1)
        return X, L

    def next(self):
        if self._pos >= self.n_images:
            self._pos = 0
        idx = range(self._pos, min(self._pos + self.batch_size, self.n_images))
        self._pos += self.batch_size
        XL = [self._image(self._first + i) for i in idx]
        return np.stack([a for a, _ in XL]), np.stack([b for _, b in XL])

    __next__ = next


def load_data(dataset='camvid', data_augm_kwargs={}, one_hot=True, batch_size=[10, 10, 10], return_0_255=False,
              which_set='test', n_images=10, height=360, width=480, seed=0, first_image=0, **_):
    """Same call shape as data_loader.load_data(dataset, {}, one_hot=True, batch_size=[..], which_set=..)."""
    if not one_hot:
        raise NotImplementedError('the iterative-inference scripts always request one_hot=True')
    bs = batch_size[{'train': 0, 'val': 1, 'valid': 1, 'test': 2}[which_set]]
    return SyntheticSegmentationIterator(n_images, bs, height, width, 11, seed, first_image)
