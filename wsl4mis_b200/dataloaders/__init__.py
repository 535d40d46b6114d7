"""GPU input pipeline: resident slice store + batch augmentation kernel."""
from .dataset import BaseDataSets, GpuLoader, RandomGenerator, SliceStore  # noqa: F401
