"""Drop-in for ``dataloaders/dataset.py`` of the reference (modelled on the working ``dataset_semi.py:17-171``).

The reference decodes one h5 slice per sample in DataLoader workers and augments it with numpy/scipy
(``RandomGenerator``: rot90+flip | rotate(order 0), then ``zoom(order 0)`` to the patch size).  At ~10 k images/s per GPU
eight such workers cannot feed the step, so here the 1 902 training slices (about 0.5 GB) live on the GPU in a ragged store
and a whole batch is augmented by ONE kernel (``wsl_augment_batch``).  All stages are nearest-neighbour, i.e. exact index
maps, and the random parameters are drawn on the host from the same ``random`` / ``numpy.random`` streams in the same
order as the reference's transform, so a batch is bit-identical to what the reference pipeline produces for the same
draws (tests/test_gpu_data.py).
"""
import os
import random

import numpy as np
import torch

from .._lib import LIB, call

_SAMPLE = np.dtype([("off", "<i8"), ("h", "<i4"), ("w", "<i4"), ("mode", "<i4"), ("k", "<i4"), ("axis", "<i4"),
                    ("lab_cval", "<i4"), ("m00", "<f8"), ("m01", "<f8"), ("m10", "<f8"), ("m11", "<f8"), ("o0", "<f8"),
                    ("o1", "<f8")], align=False)


def _rotation(angle, shape):
    """Matrix and offset of ``scipy.ndimage.rotate(..., reshape=False)`` (same float64 expressions as SciPy)."""
    try:
        from scipy import special
        c, s = special.cosdg(angle), special.sindg(angle)
    except ImportError:                                   # pragma: no cover - SciPy is a dependency of the reference too
        c, s = np.cos(np.deg2rad(angle)), np.sin(np.deg2rad(angle))
    rot = np.array([[c, s], [-s, c]])
    plane = np.asarray(shape)
    out_center = rot @ ((plane - 1) / 2)
    in_center = (plane - 1) / 2
    return rot, in_center - out_center


class RandomGenerator:
    """``RandomGenerator(output_size)`` of dataset_semi.py:146-171.  ``draw`` consumes ``random`` / ``np.random`` exactly like
    the reference's ``__call__`` (:155-163) and returns the augmentation parameters instead of applying them."""

    def __init__(self, output_size):
        self.output_size = tuple(int(v) for v in output_size)

    @staticmethod
    def draw(label_has_4):
        if random.random() > 0.5:
            k = int(np.random.randint(0, 4))              # random_rot_flip, :126-134
            axis = int(np.random.randint(0, 2))
            return (1, k, axis, 0, 0)
        elif random.random() > 0.5:
            angle = int(np.random.randint(-20, 20))       # random_rotate, :137-143
            return (2, 0, 0, angle, 4 if label_has_4 else 0)
        return (0, 0, 0, 0, 0)

    def __call__(self, sample):
        """Single-sample form with the reference's signature (numpy in, CPU tensors out); runs the same GPU kernel."""
        image = np.ascontiguousarray(sample["image"], dtype=np.float32)
        label = np.ascontiguousarray(sample["label"]).astype(np.uint8)
        store = SliceStore.from_arrays([image], [label])
        params = [self.draw(bool(store.has4[0]))]
        img, lab = store.augment([0], params, self.output_size)
        return {"image": img[0].cpu(), "label": lab[0].cpu()}


class SliceStore:
    """All training slices resident on the GPU: flat fp32 images, flat uint8 labels, per-slice offset / shape."""

    def __init__(self, images, labels, offs, shapes, has4, names):
        self.images, self.labels = images, labels
        self.offs, self.shapes, self.has4, self.names = offs, shapes, has4, names
        assert LIB.load().wsl_augment_sample_bytes() == _SAMPLE.itemsize, "AugSample layout mismatch"

    def __len__(self):
        return len(self.offs)

    @classmethod
    def from_arrays(cls, images, labels, names=None, device=None):
        device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        assert len(images) == len(labels) and len(images) > 0
        shapes = np.array([im.shape for im in images], dtype=np.int32)
        for im, lb in zip(images, labels):
            assert im.ndim == 2 and im.shape == lb.shape, "slices are 2-D and image / label shapes agree"
        sizes = shapes[:, 0].astype(np.int64) * shapes[:, 1]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        flat_i = np.concatenate([np.asarray(im, dtype=np.float32).ravel() for im in images])
        flat_l = np.concatenate([np.asarray(lb).astype(np.uint8).ravel() for lb in labels])
        has4 = np.array([bool((np.asarray(lb) == 4).any()) for lb in labels])
        names = list(names) if names is not None else [str(i) for i in range(len(images))]
        return cls(torch.from_numpy(flat_i).to(device), torch.from_numpy(flat_l).to(device), offs, shapes, has4, names)

    @classmethod
    def from_h5_dir(cls, base_dir, cases, sup_type="scribble", device=None):
        """``ACDC_training_slices/<case>`` files with datasets 'image' and `sup_type` (dataset_semi.py:105-121)."""
        try:
            import h5py
        except ImportError as e:                          # no silent substitute: the caller must provide arrays
            raise ImportError("h5py is required to read the ACDC slice files; use SliceStore.from_arrays otherwise") from e
        ims, lbs = [], []
        for case in cases:
            with h5py.File(os.path.join(base_dir, "ACDC_training_slices", case), "r") as f:
                ims.append(f["image"][:])
                lbs.append(f[sup_type][:])
        return cls.from_arrays(ims, lbs, [c.split("_")[0] for c in cases], device)

    def augment(self, indices, params, output_size):
        """indices: slice ids; params: (mode, k, axis, angle, lab_cval) per sample -> (image [B,1,H,W] fp32, label [B,H,W] u8)."""
        B = len(indices)
        tab = np.zeros(B, dtype=_SAMPLE)
        for b, (i, (mode, k, axis, angle, cval)) in enumerate(zip(indices, params)):
            h, w = int(self.shapes[i, 0]), int(self.shapes[i, 1])
            row = tab[b]
            row["off"], row["h"], row["w"] = self.offs[i], h, w
            row["mode"], row["k"], row["axis"], row["lab_cval"] = mode, k, axis, cval
            if mode == 2:
                rot, off = _rotation(angle, (h, w))
                row["m00"], row["m01"], row["m10"], row["m11"] = rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1]
                row["o0"], row["o1"] = off[0], off[1]
        dev = self.images.device
        tab_d = torch.from_numpy(tab.view(np.uint8)).to(dev, non_blocking=True)
        OH, OW = output_size
        img = torch.empty(B, 1, OH, OW, dtype=torch.float32, device=dev)
        lab = torch.empty(B, OH, OW, dtype=torch.uint8, device=dev)
        call("wsl_augment_batch", self.images, self.labels, tab_d, B, OH, OW, img, lab)
        return img, lab


class BaseDataSets(SliceStore):
    """Name kept for the scripts' ``from dataloaders.dataset import BaseDataSets, RandomGenerator``; construct with
    ``BaseDataSets.from_h5_dir`` / ``from_arrays`` and iterate with ``GpuLoader``."""


class GpuLoader:
    """Replaces ``DataLoader(db_train, batch_size, shuffle=True)`` + ``RandomGenerator`` for the training split: yields
    ``{'image': [B,1,H,W] fp32 cuda, 'label': [B,H,W] uint8 cuda, 'idx': [...]}`` (script keys, dataset_semi.py:120-124)."""

    def __init__(self, store, batch_size, output_size=(256, 256), shuffle=True, drop_last=False, rank=0, world_size=1):
        """`batch_size` is per rank.  With world_size > 1 every rank walks the SAME permutation (all ranks seed torch alike, as
        the scripts do with args.seed) and takes the strided share rank::world_size of it, like DistributedSampler."""
        self.store, self.batch_size, self.shuffle, self.drop_last = store, int(batch_size), shuffle, drop_last
        self.rank, self.world_size = int(rank), int(world_size)
        assert 0 <= self.rank < self.world_size
        self.transform = RandomGenerator(output_size)

    @staticmethod
    def shard(order, rank, world_size):
        """DistributedSampler's partition: pad the permutation by wrapping to a multiple of world_size, then stride."""
        if world_size == 1:
            return list(order)
        total = (len(order) + world_size - 1) // world_size * world_size
        padded = list(order) + list(order)[: total - len(order)]
        return padded[rank:total:world_size]

    def __len__(self):
        n = (len(self.store) + self.world_size - 1) // self.world_size
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.store)
        if self.shuffle:     # torch.utils.data.RandomSampler: fresh generator seeded from the default stream, then randperm
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        order = self.shard(order, self.rank, self.world_size)
        n = len(order)
        for s in range(0, n, self.batch_size):
            idx = order[s:s + self.batch_size]
            if len(idx) < self.batch_size and self.drop_last:
                break
            params = [self.transform.draw(bool(self.store.has4[i])) for i in idx]
            img, lab = self.store.augment(idx, params, self.transform.output_size)
            yield {"image": img, "label": lab, "idx": [self.store.names[i] for i in idx]}
