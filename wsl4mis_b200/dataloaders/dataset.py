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
        """Single-sample form with the reference's signature (numpy in, CPU tensors out).  Runs on the HOST: the scripts call
        it inside DataLoader worker processes (train_weakly_supervised_pCE_2D.py:72-73), where the GPU is off limits.  The
        batched GPU kernel (`SliceStore.augment`) produces bit-identical samples for the same draws (tests/test_gpu_data.py)."""
        image = np.asarray(sample["image"], dtype=np.float32)
        label = np.asarray(sample["label"]).astype(np.uint8)
        img, lab = apply_host(image, label, self.draw(bool((label == 4).any())), self.output_size)
        return {"image": torch.from_numpy(img).unsqueeze(0), "label": torch.from_numpy(lab)}


def apply_host(image, label, params, output_size):
    """One (mode, k, axis, angle, lab_cval) decision on host arrays with the reference's library calls: rot90 + flip
    (dataset_semi.py:126-134) | scipy rotate, order 0, label padded with lab_cval (:137-143); then zoom(order=0) to the
    patch size (:164-167)."""
    from scipy import ndimage
    mode, k, axis, angle, cval = params
    if mode == 1:
        image, label = np.flip(np.rot90(image, k), axis=axis), np.flip(np.rot90(label, k), axis=axis)
    elif mode == 2:
        image = ndimage.rotate(image, angle, order=0, reshape=False)
        label = ndimage.rotate(label, angle, order=0, reshape=False, mode="constant", cval=cval)
    fy, fx = output_size[0] / image.shape[0], output_size[1] / image.shape[1]
    image = ndimage.zoom(np.ascontiguousarray(image), (fy, fx), order=0)
    label = ndimage.zoom(np.ascontiguousarray(label), (fy, fx), order=0)
    return image.astype(np.float32), label.astype(np.uint8)


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


def fold_ids(fold):
    """ACDC five-fold split of the reference (dataset_semi.py:62-101): fold k tests on patients 20(k-1)+1 .. 20k."""
    k = {"fold1": 1, "fold2": 2, "fold3": 3, "fold4": 4, "fold5": 5}.get(fold)
    if k is None:
        raise ValueError(f"unknown fold {fold!r} (fold1..fold5)")
    test = ["patient{:0>3}".format(i) for i in range(20 * (k - 1) + 1, 20 * k + 1)]
    train = ["patient{:0>3}".format(i) for i in range(1, 101) if "patient{:0>3}".format(i) not in test]
    return train, test


def synthetic_acdc(n_patients=100, seed=2022, slices=(6, 11)):
    """ACDC-shaped synthetic cases for runs without the dataset (no h5py / no files on the box): per patient a short-axis stack
    of `slices` ragged 2-D slices, image in [0, 1] (the reference min-max normalises, acdc_data_processing.py:44-45), dense
    labels 0..3 as nested structures and a scribble map with 4 = unlabelled (~3 % labelled pixels).  Deterministic in `seed`."""
    rs = np.random.RandomState(seed)
    shapes = [(256, 216), (216, 256), (224, 154), (232, 256), (174, 208)]
    cases = {}
    for pid in range(1, n_patients + 1):
        h, w = shapes[rs.randint(len(shapes))]
        d = int(rs.randint(slices[0], slices[1]))
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        cy, cx = h * (0.4 + 0.2 * rs.rand()), w * (0.4 + 0.2 * rs.rand())
        vol_i, vol_l, vol_s = [], [], []
        for z in range(d):
            r0 = min(h, w) * (0.10 + 0.04 * np.sin(z / max(d - 1, 1) * np.pi) + 0.02 * rs.rand())
            rr = np.sqrt((yy - cy) ** 2 + ((xx - cx) * (0.9 + 0.2 * rs.rand())) ** 2)
            lab = np.zeros((h, w), np.uint8)
            lab[rr < 2.2 * r0] = 1
            lab[rr < 1.6 * r0] = 2
            lab[rr < 1.0 * r0] = 3
            img = 0.15 + 0.2 * lab.astype(np.float32) + 0.08 * rs.randn(h, w).astype(np.float32)
            img = (img - img.min()) / (img.max() - img.min())
            scr = np.full((h, w), 4, np.uint8)
            keep = rs.rand(h, w) < 0.03
            scr[keep] = lab[keep]
            vol_i.append(img.astype(np.float32))
            vol_l.append(lab)
            vol_s.append(scr)
        cases["patient{:0>3}".format(pid)] = (np.stack(vol_i), np.stack(vol_l), np.stack(vol_s))
    return cases


class BaseDataSets(torch.utils.data.Dataset):
    """``BaseDataSets(base_dir, split, transform, fold, sup_type)`` as every WSS script constructs it
    (train_weakly_supervised_pCE_GatedCRFLoss_2D.py:75-80; the class the scripts need is the one in dataset_semi.py:17-125 --
    the shipped dataloaders/dataset.py lacks fold= / sup_type=).

    split='train': one sample per slice of the fold's training patients, ``{'image': f32 [1,H,W], 'label': u8 [H,W], 'idx'}``
    after `transform`; split='val': one sample per volume of the fold's test patients, ``{'image': [D,h,w], 'label': [D,h,w],
    'idx'}``.  Samples are produced on the HOST (numpy): the scripts iterate this object inside ``DataLoader(num_workers=8)``
    worker processes, which must not touch the GPU.  ``to_slice_store()`` hands the same slices to the resident GPU pipeline
    (`SliceStore` / `GpuLoader`), which is what keeps up with the B200 step.

    base_dir: the ACDC directory (``ACDC_training_slices/*.h5`` with datasets image / label / scribble and
    ``ACDC_training_volumes/*.h5``; needs h5py), or ``'synthetic'`` / ``'synthetic:<patients>'`` for the deterministic
    ACDC-shaped generator above (what the tests and benchmarks use: the GPU boxes have neither h5py nor the data)."""

    def __init__(self, base_dir=None, split='train', transform=None, fold="fold1", sup_type="label", num=None):
        self._base_dir, self.split, self.transform, self.sup_type = base_dir, split, transform, sup_type
        train_ids, test_ids = fold_ids(fold)
        ids = set(train_ids if split == "train" else test_ids)
        if split not in ("train", "val"):
            raise ValueError(f"split must be 'train' or 'val' (got {split!r})")
        self._synthetic = None
        if base_dir is None or str(base_dir).startswith("synthetic"):
            n = int(str(base_dir).split(":")[1]) if base_dir is not None and ":" in str(base_dir) else 100
            cases = synthetic_acdc(n)
            self._synthetic = cases
            if split == "train":
                self.sample_list = [f"{pid}_frame01_slice_{z}.h5" for pid in sorted(cases) if pid in ids for z in range(cases[pid][0].shape[0])]
            else:
                self.sample_list = [f"{pid}_frame01.h5" for pid in sorted(cases) if pid in ids]
        else:
            sub = "ACDC_training_slices" if split == "train" else "ACDC_training_volumes"
            names = sorted(os.listdir(os.path.join(base_dir, sub)))
            self.sample_list = [f for f in names if f.split("_")[0] in ids]
        if num is not None and split == "train":
            self.sample_list = self.sample_list[:num]

    def __len__(self):
        return len(self.sample_list)

    def _read(self, case):
        """-> (image, label) numpy arrays of one train slice / one validation volume"""
        key = self.sup_type if self.split == "train" else "label"
        if self._synthetic is not None:
            img, lab, scr = self._synthetic[case.split("_")[0]]
            if self.split == "train":
                z = int(case.rsplit("_", 1)[1].split(".")[0])
                return img[z], (scr if key == "scribble" else lab)[z]
            return img, lab
        try:
            import h5py
        except ImportError as e:
            raise ImportError("h5py is required to read the ACDC files; pass base_dir='synthetic' for the generated stand-in") from e
        sub = "ACDC_training_slices" if self.split == "train" else "ACDC_training_volumes"
        with h5py.File(os.path.join(self._base_dir, sub, case), "r") as f:
            return f["image"][:], f[key][:]

    def __getitem__(self, idx):
        case = self.sample_list[idx]
        image, label = self._read(case)
        sample = {"image": image, "label": label}
        if self.split == "train" and self.transform is not None:
            sample = self.transform(sample)
        sample["idx"] = case.split("_")[0]
        return sample

    def to_slice_store(self, device=None):
        """the training slices as a resident GPU `SliceStore` (feed it to `GpuLoader`)"""
        assert self.split == "train"
        pairs = [self._read(c) for c in self.sample_list]
        return SliceStore.from_arrays([p[0] for p in pairs], [p[1] for p in pairs], [c.split("_")[0] for c in self.sample_list], device)


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
