"""Reference-shaped view of the memory banks (API of deva/inference/kv_memory_store.py:241-277).

The reference keeps two independent ``KeyValueMemoryStore`` objects (working and long-term), each a
set of dicts of ever-growing tensors.  Here both are *views* over the same ``BucketBank`` objects
(deva/inference/memory_bank.py): ``kind='work'`` exposes the working region of every bucket,
``kind='long'`` the long-term region.  The views exist so code written against the reference
(``store.key[bucket]``, ``store.value[obj]``, ``store.size(b)``, ``store.engaged()``...) keeps working;
mutation goes through ``MemoryManager``.
"""
from typing import Dict, List, Optional

import torch

from deva.inference.memory_bank import BucketBank


class _LazyDict:
    """Read-only mapping whose values are built on access (tensor views into the banks)."""
    def __init__(self, keys_fn, get_fn):
        self._keys_fn, self._get_fn = keys_fn, get_fn

    def __getitem__(self, k):
        if k not in self._keys_fn():
            raise KeyError(k)
        return self._get_fn(k)

    def __contains__(self, k):
        return k in self._keys_fn()

    def __iter__(self):
        return iter(self._keys_fn())

    def __len__(self):
        return len(self._keys_fn())

    def keys(self):
        return list(self._keys_fn())

    def items(self):
        return [(k, self._get_fn(k)) for k in self._keys_fn()]

    def values(self):
        return [self._get_fn(k) for k in self._keys_fn()]


class KeyValueMemoryStore:
    def __init__(self, banks: Dict[int, BucketBank], kind: str, save_selection: bool = False,
                 save_usage: bool = False):
        assert kind in ('work', 'long')
        self._banks = banks
        self.kind = kind
        self.save_selection = save_selection
        self.save_usage = save_usage

    # ---- which buckets exist in this store -------------------------------------------------
    def _live(self) -> List[int]:
        if self.kind == 'work':
            return list(self._banks.keys())
        return [b for b, bank in self._banks.items() if bank.long_size > 0]

    def _bank_of(self, obj: int) -> BucketBank:
        for b in self._live():
            if obj in self._banks[b].objects:
                return self._banks[b]
        raise KeyError(obj)

    def _objects(self) -> List[int]:
        return [o for b in self._live() for o in self._banks[b].objects]

    @property
    def buckets(self) -> Dict[int, List[int]]:
        return {b: list(self._banks[b].objects) for b in self._live()}

    def size(self, bucket_id: int) -> int:
        bank = self._banks.get(bucket_id)
        if bank is None:
            return 0
        return bank.work_size if self.kind == 'work' else bank.long_size

    def engaged(self, bucket_id: Optional[int] = None) -> bool:
        live = self._live()
        return len(live) > 0 if bucket_id is None else bucket_id in live

    def get_v_size(self, obj_id: int) -> int:
        bank = self._bank_of(obj_id)
        return bank.work_size if self.kind == 'work' else bank.long_size

    @property
    def num_objects(self) -> int:
        return len(self._objects())

    def __contains__(self, obj) -> bool:
        return obj in self._objects()

    # ---- tensor views (reference layouts: key [CK,N], shrinkage [1,N], value [CV,N]) -------
    @property
    def key(self) -> Dict[int, torch.Tensor]:
        return _LazyDict(self._live, lambda b: self._banks[b].key_view(self.kind))

    @property
    def shrinkage(self) -> Dict[int, torch.Tensor]:
        return _LazyDict(self._live, lambda b: self._banks[b].shrinkage_view(self.kind))

    @property
    def selection(self) -> Dict[int, torch.Tensor]:
        if not self.save_selection:
            raise AttributeError('this store does not keep the selection term')
        return _LazyDict(self._live, lambda b: self._banks[b].selection_view(self.kind))

    @property
    def value(self) -> Dict[int, torch.Tensor]:
        """fp16 views [CV, N] (the bank stores values in the readout GEMM's operand precision)."""
        return _LazyDict(self._objects, lambda o: self._bank_of(o).value_view(o, self.kind))

    def get_usage(self, bucket_id: int) -> torch.Tensor:
        if not self.save_usage:
            raise RuntimeError('I did not count usage!')
        bank = self._banks[bucket_id]
        a, b = (bank.lo, bank.base) if self.kind == 'long' else (bank.base, bank.hi)
        return bank.use_cnt[a:b] / bank.life_cnt[a:b]
