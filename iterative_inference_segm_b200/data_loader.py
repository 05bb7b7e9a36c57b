"""Drop-in for data_loader.load_data (data_loader.py:7-121).

The reference wraps the external, un-vendored `dataset_loaders` package (README.md:22).  What the path needs from it
is the iterator's contract (iterative_inference.py:117-125, 233-234; train_dae.py:130-160):
`.next()` -> (X float32 (B,3,H,W) in [0,1] (or [0,255] with return_0_255), L one-hot float32 (B,C+1,H,W), last
channel = void), `.nbatches`, `.non_void_nclasses`, `.void_labels`, `.data_shape` (channels first), `.cmap`,
`.mask_labels`.  Two iterators provide it:

  * `CamvidDirectoryIterator` -- reads a CamVid tree in the layout `dataset_loaders.images.camvid.CamvidDataset`
    reads (the SegNet-tutorial layout: `<path>/<set>/*.png` images, `<path>/<set>annot/*.png` label maps with
    values 0..10 and 11 = void), with the training-time crop / horizontal flip of `data_augm_kwargs`, optional
    shuffling and a prefetch thread (`use_threads=True` in the reference, data_loader.py:58);
  * `SyntheticSegmentationIterator` -- CamVid-shaped seeded random data (benchmarks, tests, no dataset on the box).

`load_data` keeps the reference's call shape and picks the directory reader when a dataset root is given
(`path=` or $IISEG_DATA/<dataset>), else the synthetic iterator.
"""
import os
import queue
import threading

import numpy as np

CAMVID_LABELS = ['sky', 'building', 'column_pole', 'road', 'sidewalk', 'tree', 'sign', 'fence', 'car', 'pedestrian',
                 'byciclist', 'void']
CAMVID_CMAP = np.array([(128, 128, 128), (128, 0, 0), (192, 192, 128), (128, 64, 128), (0, 0, 192), (128, 128, 0),
                        (192, 128, 128), (64, 64, 128), (64, 0, 128), (64, 64, 0), (0, 128, 192), (0, 0, 0)],
                       dtype=np.float32) / 255.0


class _IteratorBase(object):
    non_void_nclasses = 11
    void_labels = [11]

    def __next__(self):
        return self.next()

    def __iter__(self):
        return self


class SyntheticSegmentationIterator(_IteratorBase):
    def __init__(self, n_images=10, batch_size=10, height=360, width=480, n_classes=11, seed=0,
                 first_image=0, return_0_255=False):
        self.non_void_nclasses = n_classes
        self.void_labels = [n_classes]
        self.data_shape = (3, height, width)
        self.batch_size = batch_size
        self.n_images = n_images
        self.nbatches = (n_images + batch_size - 1) // batch_size
        self.cmap = np.linspace(0, 1, (n_classes + 1) * 3).reshape(n_classes + 1, 3)
        self.mask_labels = ['class%d' % i for i in range(n_classes)] + ['void']
        self._seed, self._first, self._pos = seed, first_image, 0
        self._scale = 255.0 if return_0_255 else 1.0

    def _image(self, idx):
        rng = np.random.RandomState(self._seed * 1000003 + idx)
        C, H, W = self.data_shape
        X = rng.rand(C, H, W).astype(np.float32) * np.float32(self._scale)
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


class CamvidDirectoryIterator(_IteratorBase):
    """CamVid from disk.  `path`/<which_set>/NAME.png (RGB) pairs with `path`/<which_set>annot/NAME.png (label map,
    uint8 values 0..n_classes, n_classes = void); every image of the set is visited once per epoch (`nbatches` =
    ceil(n / batch_size), the last batch may be short), in sorted order or reshuffled at each epoch.
    `data_augm_kwargs`: `crop_size` (h, w) random crop and `horizontal_flip` probability (training only).
    `shard` = (rank, world): this process iterates over its contiguous shard of the BATCHES (sharding.shard_range)."""

    def __init__(self, path, which_set='test', batch_size=10, data_augm_kwargs=None, return_0_255=False,
                 shuffle_at_each_epoch=False, use_threads=True, n_classes=11, seed=0, shard=None, prefetch=2):
        from PIL import Image             # the only image dependency; probed in this image (PIL 12.2)
        self._Image = Image
        self.image_path = os.path.join(path, which_set)
        self.mask_path = os.path.join(path, which_set + 'annot')
        if not os.path.isdir(self.image_path) or not os.path.isdir(self.mask_path):
            raise IOError('CamVid tree not found: %s and %s must exist' % (self.image_path, self.mask_path))
        names = sorted(f for f in os.listdir(self.image_path) if f.lower().endswith(('.png', '.jpg', '.jpeg', '.bmp')))
        if not names:
            raise IOError('no images under %s' % self.image_path)
        self.names = names
        self.non_void_nclasses = n_classes
        self.void_labels = [n_classes]
        self.batch_size = batch_size
        self.cmap = CAMVID_CMAP if n_classes == 11 else np.linspace(0, 1, (n_classes + 1) * 3).reshape(n_classes + 1, 3)
        self.mask_labels = CAMVID_LABELS if n_classes == 11 else ['class%d' % i for i in range(n_classes)] + ['void']
        self._aug = dict(data_augm_kwargs or {})
        self._div = np.float32(1.0 if return_0_255 else 255.0)
        self._shuffle = shuffle_at_each_epoch
        self._rng = np.random.RandomState(seed)
        X0, _ = self._load(0, augment=False)
        crop = self._aug.get('crop_size')
        self.data_shape = (X0.shape[0],) + (tuple(crop) if crop else X0.shape[1:])
        n_all = (len(names) + batch_size - 1) // batch_size
        if shard is not None:
            from .sharding import shard_range
            self._b_lo, self._b_hi = shard_range(n_all, shard[0], shard[1])
        else:
            self._b_lo, self._b_hi = 0, n_all
        self.nbatches = self._b_hi - self._b_lo
        self.nbatches_total = n_all
        self._order, self._pos = None, self._b_lo
        self._q, self._thread = None, None
        if use_threads and self.nbatches > 0:
            self._q = queue.Queue(maxsize=prefetch)
            self._thread = threading.Thread(target=self._producer, daemon=True)
            self._thread.start()

    def _load(self, i, augment=True):
        name = self.names[i]
        img = np.asarray(self._Image.open(os.path.join(self.image_path, name)).convert('RGB'), dtype=np.uint8)
        stem = os.path.splitext(name)[0]
        mpath = os.path.join(self.mask_path, name)
        if not os.path.exists(mpath):
            mpath = os.path.join(self.mask_path, stem + '.png')
        lab = np.asarray(self._Image.open(mpath), dtype=np.uint8)
        if lab.ndim == 3:
            lab = lab[..., 0]
        if lab.shape != img.shape[:2]:
            raise ValueError('%s: image %s and label map %s differ in size' % (name, img.shape[:2], lab.shape))
        if augment and self._aug:
            crop = self._aug.get('crop_size')
            if crop:
                ch, cw = crop
                top = self._rng.randint(0, img.shape[0] - ch + 1)
                left = self._rng.randint(0, img.shape[1] - cw + 1)
                img, lab = img[top:top + ch, left:left + cw], lab[top:top + ch, left:left + cw]
            if self._rng.rand() < float(self._aug.get('horizontal_flip', 0.0)):
                img, lab = img[:, ::-1], lab[:, ::-1]
        X = img.transpose(2, 0, 1).astype(np.float32) / self._div
        lab = np.minimum(lab, self.non_void_nclasses)                 # anything above the last class is void
        L = np.eye(self.non_void_nclasses + 1, dtype=np.float32)[lab].transpose(2, 0, 1)
        return X, np.ascontiguousarray(L)

    def _batch(self):
        if self._order is None or self._pos >= self._b_hi:
            self._order = self._rng.permutation(len(self.names)) if self._shuffle else np.arange(len(self.names))
            self._pos = self._b_lo
        idx = self._order[self._pos * self.batch_size:(self._pos + 1) * self.batch_size]
        self._pos += 1
        XL = [self._load(int(i)) for i in idx]
        return np.stack([a for a, _ in XL]), np.stack([b for _, b in XL])

    def _producer(self):
        while True:
            try:
                self._q.put(self._batch())
            except Exception as e:          # surfaces in next() instead of killing the thread silently
                self._q.put(e)
                return

    def next(self):
        if self._q is None:
            return self._batch()
        item = self._q.get()
        if isinstance(item, Exception):
            raise item
        return item


def load_data(dataset='camvid', data_augm_kwargs={}, one_hot=True, batch_size=[10, 10, 10], return_0_255=False,
              which_set='test', shuffle_train=True, path=None, shard=None, n_images=10, height=360, width=480, seed=0,
              first_image=0, **_):
    """Same call shape as data_loader.load_data(dataset, {}, one_hot=True, batch_size=[..], which_set=..)
    (data_loader.py:7-9).  `path` (or $IISEG_DATA/<dataset>) selects the on-disk CamVid reader; otherwise the data are
    synthetic (`n_images`, `height`, `width`, `seed`).  which_set='all' returns [train, val, test] like the reference."""
    if not one_hot:
        raise NotImplementedError('the iterative-inference scripts always request one_hot=True')
    if which_set == 'all':
        return [load_data(dataset, data_augm_kwargs, one_hot, batch_size, return_0_255, s, shuffle_train, path, shard,
                          n_images, height, width, seed, first_image) for s in ('train', 'val', 'test')]
    if which_set not in ('train', 'val', 'valid', 'test'):
        raise ValueError('which_set must be all, train, val or test')
    bs = batch_size[{'train': 0, 'val': 1, 'valid': 1, 'test': 2}[which_set]]
    if path is None and os.environ.get('IISEG_DATA'):
        cand = os.path.join(os.environ['IISEG_DATA'], dataset)
        path = cand if os.path.isdir(cand) else None
    train = which_set == 'train'
    if path is not None:
        if dataset != 'camvid':
            raise NotImplementedError('on-disk reader: camvid (the benchmark dataset); %r is not built' % dataset)
        return CamvidDirectoryIterator(path, 'val' if which_set == 'valid' else which_set, bs,
                                       data_augm_kwargs if train else None, return_0_255,
                                       shuffle_at_each_epoch=train and shuffle_train, seed=seed, shard=shard)
    crop = (data_augm_kwargs or {}).get('crop_size') if train else None
    if crop:
        height, width = crop
    return SyntheticSegmentationIterator(n_images, bs, height, width, 11, seed, first_image, return_0_255)
