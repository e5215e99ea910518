"""dl_attack_on_imagenet_b200 -- B200-native (sm_100a) ADiL attack-learning hot path.

`ADIL` mirrors the reference class (attacks/attacks_classes/adil.py); `ops` exposes the CUDA kernels behind the
C ABI of include/adil_b200.h.  Importing the package never touches the GPU; the shared library is loaded on
first use and its absence is an error (no CPU fallback).
"""
from . import ops  # noqa: F401
from .adil import ADIL, AdilState, Attack_dict_model, split_normalize  # noqa: F401
from .build import build_library  # noqa: F401
from .data import HostBatchPrefetcher, IndexedTensorDataset, Normalize, build_classifier, synthetic_images  # noqa: F401

__all__ = ["ADIL", "AdilState", "Attack_dict_model", "split_normalize", "ops", "build_library",
           "HostBatchPrefetcher", "IndexedTensorDataset", "Normalize", "build_classifier", "synthetic_images"]
