"""Per-frame cache of encoder outputs (API of the reference's deva/inference/image_feature_store.py:7-48)."""
import warnings
from typing import Dict, Iterable, Tuple

import torch


class ImageFeatureStore:
    """Maps a frame index to (ms_features, key_feat, key, shrinkage, selection); the caller deletes entries."""
    def __init__(self, network, no_warning: bool = False):
        self.network = network
        self.no_warning = no_warning
        self._store: Dict[int, Tuple] = {}

    def _entry(self, index: int, image: torch.Tensor) -> Tuple:
        hit = self._store.get(index)
        if hit is None:
            ms_features, feat = self.network.encode_image(image)
            hit = (ms_features, feat, *self.network.transform_key(feat))
            self._store[index] = hit
        return hit

    def get_ms_features(self, index: int, image: torch.Tensor) -> Iterable[torch.Tensor]:
        return self._entry(index, image)[0]

    def get_key(self, index: int, image: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return self._entry(index, image)[2:]

    def delete(self, index: int) -> None:
        self._store.pop(index, None)

    def __len__(self):
        return len(self._store)

    def __del__(self):
        if len(self._store) > 0 and not self.no_warning:
            warnings.warn(f'Leaking {self._store.keys()} in the image feature store')
