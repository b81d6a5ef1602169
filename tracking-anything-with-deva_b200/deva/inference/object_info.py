"""Per-object metadata record (API of the reference's deva/inference/object_info.py:7-62)."""
from typing import Optional

import numpy as np


class ObjectInfo:
    """Identity is the integer ``id`` (hash / equality), everything else is voting metadata."""
    __slots__ = ('id', 'category_ids', 'scores', 'isthing', 'poke_count')

    def __init__(self, id: int, category_id: Optional[int] = None, isthing: Optional[bool] = None,
                 score: Optional[float] = None):
        self.id = id
        self.category_ids = [category_id]
        self.scores = [score]
        self.isthing = isthing
        self.poke_count = 0  # detections in a row that missed this object

    def poke(self) -> None:
        self.poke_count += 1

    def unpoke(self) -> None:
        self.poke_count = 0

    def merge(self, other: 'ObjectInfo') -> None:
        self.category_ids += other.category_ids
        self.scores += other.scores

    def copy_meta_info(self, other: 'ObjectInfo') -> None:
        self.category_ids, self.scores, self.isthing = other.category_ids, other.scores, other.isthing

    def vote_category_id(self) -> Optional[int]:
        votes = [c for c in self.category_ids if c is not None]
        if not votes:
            return None
        # most frequent, smallest on ties (scipy.stats.mode semantics used by the reference, :40)
        vals, counts = np.unique(np.asarray(votes), return_counts=True)
        return int(vals[np.argmax(counts)])

    def vote_score(self) -> Optional[float]:
        votes = [s for s in self.scores if s is not None]
        return float(np.mean(votes)) if votes else None

    def get_rgb(self) -> np.ndarray:
        # panoptic-style id -> colour (reference deva/utils/pano_utils.py id_to_rgb)
        i = int(self.id)
        return np.array([i % 256, (i // 256) % 256, (i // 65536) % 256], dtype=np.uint8)

    def __hash__(self):
        return hash(self.id)

    def __eq__(self, other):
        return self.id == other.id

    def __repr__(self):
        return f'(ID: {self.id}, cat: {self.category_ids}, isthing: {self.isthing}, score: {self.scores})'
