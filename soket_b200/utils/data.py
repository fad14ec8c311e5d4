"""soket_b200.utils.data -- the input pipeline of soket/utils/data, device resident.

SURVEY.md section 8(f)-2.  The reference's ``DataLoader.__next__``
(soket/utils/data/loader.py:64-81) indexes the dataset once per SAMPLE, wraps every
sample in a Tensor (an H2D copy each on the GPU device) and ``stack``s them
(soket/tensor/util.pyx:35-40): 2 x batch_size small copies per step.  Here the same
classes exist with the same constructor arguments, the same batch order (the
``numpy.array_split`` of ``arange`` / ``numpy.random.permutation`` the reference
computes, loader.py:47-61) and the same batch contents, but a dataset that exposes its
samples as whole arrays (``.data`` / ``.targets``, as ``MNIST`` does, mnist.py:28-41)
is uploaded ONCE and each batch is ONE row-gather kernel per array
(``sk_gather_rows``) driven by an index array that is itself uploaded once per epoch:
no host->device traffic inside the epoch at all.

Names / behaviour kept from the reference: ``Dataset`` (dataset.py:5-24), ``MNIST``
(datasets/mnist.py:8-63; idx-gz layout, magic numbers 2051 / 2049, pixels / 255 in
float32, uint8 labels), ``collate`` and ``DataLoader`` (loader.py:9-81; ``max_iter``,
``ordering``, ``idx`` attributes).  Added: ``ResidentDataset`` and the ``shard=(rank,
world)`` argument (data-parallel ranks take contiguous row blocks of every global batch,
SURVEY.md section 8e).
"""
from __future__ import annotations

import gzip as _gz
import struct as _struct
import warnings
from abc import ABC, abstractmethod
from math import ceil

import numpy as np

from soket_b200 import _core as B
from soket_b200.engine import Tensor, stack
from soket_b200.transforms import ToTensor, Transform  # noqa: F401


# ------------------------------------------------------------------ host-side order logic
def batch_bounds(n: int, max_iter: int):
    """Start offsets (max_iter + 1 of them) of the chunks ``numpy.array_split(x, max_iter)``
    cuts a length-n axis into (loader.py:49,59): the first ``n % max_iter`` chunks hold one
    element more."""
    if max_iter <= 0:
        return [0]
    each, extras = divmod(n, max_iter)
    sizes = [each + 1] * extras + [each] * (max_iter - extras)
    out = [0]
    for s in sizes:
        out.append(out[-1] + s)
    return out


def shard_bounds(batch_len: int, rank: int, world: int):
    """Rows [lo, hi) of one global batch that data-parallel rank `rank` of `world` takes
    (contiguous, sizes differing by at most one row)."""
    if not (0 <= rank < world):
        raise ValueError(f'shard: rank {rank} outside world of {world}')
    b = batch_bounds(batch_len, world)
    return b[rank], b[rank + 1]


def epoch_permutation(n: int, shuffle: bool):
    """The sample order of one epoch: ``arange(n)`` or, shuffling, ONE draw of the legacy
    global ``numpy.random.permutation(n)`` -- the same stream the reference consumes
    (loader.py:58), so seeding ``numpy.random`` reproduces its batches."""
    return np.random.permutation(n) if shuffle else np.arange(n)


# ------------------------------------------------------------------ datasets
class Dataset(ABC):
    """dataset.py:5-24."""

    def __init__(self, transforms: Transform = None):
        self.transforms = transforms

    @abstractmethod
    def __getitem__(self, index):
        return NotImplementedError

    @abstractmethod
    def __len__(self):
        return NotImplementedError


def read_idx_images(filename):
    """mnist.py:21-35: big-endian (magic 2051, count, height, width) header, uint8 pixels,
    flattened per sample, float32, divided by 255 in place."""
    with _gz.open(filename, 'rb') as f:
        magic, num, height, width = _struct.unpack('>iiii', f.read(16))
        if magic != 2051:
            raise AssertionError(f'{filename}: image file magic is {magic}, expected 2051')
        data = np.frombuffer(f.read(), dtype=np.uint8).reshape(num, height * width).astype(np.float32)
        data /= 255.
    return data


def read_idx_labels(filename):
    """mnist.py:37-43: big-endian (magic 2049, count) header, uint8 labels."""
    with _gz.open(filename, 'rb') as f:
        magic, num = _struct.unpack('>ii', f.read(8))
        if magic != 2049:
            raise AssertionError(f'{filename}: label file magic is {magic}, expected 2049')
        return np.frombuffer(f.read(), dtype=np.uint8)


class MNIST(Dataset):
    """datasets/mnist.py:8-63.  ``.data`` (N, 784) float32 in [0, 1], ``.targets`` (N,) uint8,
    both host arrays; ``DataLoader`` uploads them once (see ``ResidentDataset``)."""

    def __init__(self, images_filename, labels_filename, transforms=None, target_transforms=None):
        super().__init__(transforms)
        self.target_transforms = target_transforms
        self.data = read_idx_images(images_filename)
        self.targets = read_idx_labels(labels_filename)

    def __getitem__(self, index):
        img, target = self.data[index], self.targets[index]
        if self.transforms is not None:
            img = self.transforms(img)
        if self.target_transforms is not None:
            target = self.target_transforms(target)
        return img, target

    def __len__(self):
        return len(self.data)


class ArrayDataset(Dataset):
    """Samples held as whole host arrays (synthetic data of the bench configurations)."""

    def __init__(self, data, targets=None, transforms=None, target_transforms=None):
        super().__init__(transforms)
        self.target_transforms = target_transforms
        self.data = np.asarray(data)
        self.targets = None if targets is None else np.asarray(targets)
        if self.targets is not None and len(self.targets) != len(self.data):
            raise ValueError('ArrayDataset: data and targets differ in length')

    def __getitem__(self, index):
        x = self.data[index]
        if self.transforms is not None:
            x = self.transforms(x)
        if self.targets is None:
            return x
        t = self.targets[index]
        if self.target_transforms is not None:
            t = self.target_transforms(t)
        return x, t

    def __len__(self):
        return len(self.data)


class ResidentDataset(Dataset):
    """A dataset whose sample arrays live in HBM.  ``gather(index)`` returns one batch
    (Tensor, or (Tensor, Tensor) with targets) with one ``sk_gather_rows`` launch per array;
    ``index`` is a device or host integer array.  ``ds[i]`` keeps the per-sample protocol
    (zero-copy row views)."""

    def __init__(self, data, targets=None):
        super().__init__(None)
        self.data = B.ascontiguousarray(B.asarray(data))
        self.targets = None if targets is None else B.ascontiguousarray(B.asarray(targets))
        if self.targets is not None and self.targets.shape[0] != self.data.shape[0]:
            raise ValueError('ResidentDataset: data and targets differ in length')

    @staticmethod
    def from_dataset(ds):
        """Upload a dataset that exposes ``.data`` (and optionally ``.targets``) arrays, provided
        its per-sample transforms are absent or ``ToTensor`` (anything else changes sample
        values and has to run per sample)."""
        if isinstance(ds, ResidentDataset):
            return ds
        if not is_array_backed(ds):
            raise TypeError('ResidentDataset.from_dataset: dataset has no whole-array .data with identity/ToTensor transforms')
        return ResidentDataset(ds.data, getattr(ds, 'targets', None))

    def gather(self, index):
        x = Tensor._const(self.data[index])
        if self.targets is None:
            return x
        return x, Tensor._const(self.targets[index])

    def __getitem__(self, index):
        x = Tensor._const(self.data[index])
        if self.targets is None:
            return x
        return x, Tensor._const(self.targets[index])

    def __len__(self):
        return int(self.data.shape[0])


def _identity_or_totensor(t):
    return t is None or type(t) is ToTensor


def is_array_backed(ds):
    """True when batches of `ds` can be gathered from whole arrays without changing what the
    per-sample path would produce."""
    if isinstance(ds, ResidentDataset):
        return True
    data = getattr(ds, 'data', None)
    if not isinstance(data, (np.ndarray, B.ndarray)):
        return False
    targets = getattr(ds, 'targets', None)
    if targets is not None and not isinstance(targets, (np.ndarray, B.ndarray)):
        return False
    return _identity_or_totensor(getattr(ds, 'transforms', None)) and \
        _identity_or_totensor(getattr(ds, 'target_transforms', None))


# ------------------------------------------------------------------ loader
def collate(sequence):
    """loader.py:9-18."""
    if isinstance(sequence[0], Tensor):
        return stack(sequence)
    elif isinstance(sequence[0], (int, float, bool)):
        return Tensor(sequence)
    warnings.warn('Received sequence form dataset in dataloader is not Tensor. '
                  'Try using soket.transforms.ToTensor() to transform samples to Tensor otherwise '
                  'use the sequence at your own risk!')
    return sequence


class DataLoader:
    """loader.py:21-81 with a device-resident fast path.

    resident: None (default) = use the gather path whenever ``is_array_backed(dataset)``;
              True = require it; False = always the reference's per-sample path.
    shard:    (rank, world) -- this process takes its contiguous row block of every batch.
    """

    def __init__(self, dataset, batch_size=1, shuffle=False, resident=None, shard=None):
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.max_iter = ceil(len(dataset) / batch_size)
        self.shard = shard
        if shard is not None:
            shard_bounds(1, *shard)  # validates rank / world
        if resident is None:
            resident = is_array_backed(dataset)
        elif resident and not is_array_backed(dataset):
            raise TypeError('DataLoader(resident=True): dataset is not array backed')
        self.resident = bool(resident)
        self._resident_ds = None
        self._order_dev = None
        self._bounds = batch_bounds(len(dataset), self.max_iter)
        if not shuffle:
            self._set_order(epoch_permutation(len(dataset), False))

    def _set_order(self, perm):
        self._perm = perm
        # same object layout as the reference: a list of max_iter index arrays
        self.ordering = [perm[self._bounds[i]:self._bounds[i + 1]] for i in range(self.max_iter)]
        self._order_dev = None

    def __len__(self):
        return self.max_iter

    def __iter__(self):
        self.idx = 0
        if self.shuffle:
            self._set_order(epoch_permutation(len(self.dataset), True))
        return self

    def _batch_range(self, i):
        lo, hi = self._bounds[i], self._bounds[i + 1]
        if self.shard is not None:
            s_lo, s_hi = shard_bounds(hi - lo, *self.shard)
            lo, hi = lo + s_lo, lo + s_hi
        return lo, hi

    def __next__(self):
        if self.idx >= self.max_iter:
            raise StopIteration()
        lo, hi = self._batch_range(self.idx)
        self.idx += 1
        if self.resident:
            if self._resident_ds is None:
                self._resident_ds = ResidentDataset.from_dataset(self.dataset)   # one upload, ever
            if self._order_dev is None:
                self._order_dev = B.array(np.ascontiguousarray(self._perm, dtype=np.int64))  # one per epoch
            return self._resident_ds.gather(self._order_dev[lo:hi])
        samples = [self.dataset[i] for i in self._perm[lo:hi]]
        if isinstance(samples[0], (list, tuple)):
            data, targets = [x[0] for x in samples], [x[1] for x in samples]
            return collate(data), collate(targets)
        return collate(samples)
