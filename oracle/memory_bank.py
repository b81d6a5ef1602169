"""Oracle: working / long-term memory bank bookkeeping (test infrastructure).

Restates deva/inference/kv_memory_store.py:35-239 and
deva/inference/memory_manager.py:91-276 of the reference with one flat ``Bucket`` record per
object group instead of five parallel dicts.  Tensors stay channel-major fp32 and grow by
concatenation exactly like the reference, so sizes, ordering and usage counters can be
compared slot by slot.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import memory_math as mm


@dataclass
class Bucket:
    objects: List[int]
    key: Optional[torch.Tensor] = None  # [CK, N]
    shrinkage: Optional[torch.Tensor] = None  # [1, N]
    selection: Optional[torch.Tensor] = None  # [CK, N] (working memory only)
    use_cnt: Optional[torch.Tensor] = None  # [N]
    life_cnt: Optional[torch.Tensor] = None  # [N]

    @property
    def size(self) -> int:
        return 0 if self.key is None else self.key.shape[-1]


def _cat(a, b):
    return b if a is None else torch.cat([a, b], -1)


class Store:
    """One key/value store (working or long-term).  kv_memory_store.py:4-33."""
    def __init__(self, keep_selection: bool, keep_usage: bool):
        self.keep_selection = keep_selection
        self.keep_usage = keep_usage
        self.next_bucket = 0
        self.buckets: Dict[int, Bucket] = {}
        self.values: Dict[int, torch.Tensor] = {}  # object id -> [CV, N]

    # -- kv_memory_store.py:35-116 ---------------------------------------------------------
    def append(self, key, values: Dict[int, torch.Tensor], shrinkage, selection,
               forced_bucket: int = -1) -> None:
        touched = []
        if forced_bucket >= 0:
            exists = forced_bucket in self.buckets
            for obj, v in values.items():
                if exists:
                    assert obj in self.values and obj in self.buckets[forced_bucket].objects
                    self.values[obj] = torch.cat([self.values[obj], v], -1)
                else:
                    assert obj not in self.values
                    self.values[obj] = v
            if exists:
                self.buckets[forced_bucket].objects = list(values.keys())
            else:
                self.buckets[forced_bucket] = Bucket(list(values.keys()))
            touched = [forced_bucket]
        else:
            fresh = None
            for obj, v in values.items():
                if obj in self.values:
                    self.values[obj] = torch.cat([self.values[obj], v], -1)
                    owner = [b for b, rec in self.buckets.items() if obj in rec.objects]
                    assert len(owner) == 1
                    if owner[0] not in touched:
                        touched.append(owner[0])
                else:
                    self.values[obj] = v
                    if fresh is None:
                        fresh = self.next_bucket
                        self.next_bucket += 1
                        self.buckets[fresh] = Bucket([])
                    self.buckets[fresh].objects.append(obj)
                    if fresh not in touched:
                        touched.append(fresh)
        n_new = key.shape[1]
        for b, rec in self.buckets.items():
            if b not in touched:
                continue
            rec.key = _cat(rec.key, key)
            rec.shrinkage = _cat(rec.shrinkage, shrinkage)
            if self.keep_selection:
                rec.selection = _cat(rec.selection, selection)
            if self.keep_usage:
                rec.use_cnt = _cat(rec.use_cnt, torch.zeros(n_new))
                rec.life_cnt = _cat(rec.life_cnt, torch.zeros(n_new) + 1e-7)  # :94

    # -- kv_memory_store.py:118-125 --------------------------------------------------------
    def add_usage(self, bucket: int, usage: torch.Tensor) -> None:
        if not self.keep_usage:
            return
        rec = self.buckets[bucket]
        rec.use_cnt = rec.use_cnt + usage.reshape(-1)
        rec.life_cnt = rec.life_cnt + 1

    def usage(self, bucket: int) -> torch.Tensor:  # :187-193
        rec = self.buckets[bucket]
        return rec.use_cnt / rec.life_cnt

    # -- kv_memory_store.py:127-159 --------------------------------------------------------
    def drop_range(self, bucket: int, start: int, end: int, min_size: int) -> None:
        rec = self.buckets[bucket]
        n = rec.size
        if n <= min_size:
            return
        if end == 0:
            end = n
        assert end < 0
        keep = torch.cat([torch.arange(0, start), torch.arange(n + end, n)])
        self._take(bucket, keep)

    # -- kv_memory_store.py:164-185 --------------------------------------------------------
    def evict_least_used(self, bucket: int, max_size: int) -> None:
        u = self.usage(bucket)
        smallest, _ = torch.topk(u, k=self.buckets[bucket].size - max_size, largest=False,
                                 sorted=True)
        self._take(bucket, torch.nonzero(u > smallest[-1]).reshape(-1))  # strict '>' (Q5)

    def _take(self, bucket: int, keep: torch.Tensor) -> None:
        rec = self.buckets[bucket]
        rec.key = rec.key[:, keep]
        rec.shrinkage = rec.shrinkage[:, keep]
        if self.keep_selection and rec.selection is not None:
            rec.selection = rec.selection[:, keep]
        if self.keep_usage:
            rec.use_cnt = rec.use_cnt[keep]
            rec.life_cnt = rec.life_cnt[keep]
        for obj in rec.objects:
            self.values[obj] = self.values[obj][:, keep]

    # -- kv_memory_store.py:213-239 --------------------------------------------------------
    def keep_only(self, keep_ids) -> None:
        keep_ids = set(keep_ids)
        for b in list(self.buckets.keys()):
            rec = self.buckets[b]
            rec.objects = [o for o in rec.objects if o in keep_ids]
            if not rec.objects:
                del self.buckets[b]
        self.values = {o: v for o, v in self.values.items() if o in keep_ids}

    def size(self, bucket: int) -> int:
        return self.buckets[bucket].size if bucket in self.buckets else 0

    def engaged(self, bucket: Optional[int] = None) -> bool:
        return len(self.buckets) > 0 if bucket is None else bucket in self.buckets


class MemoryOracle:
    """memory_manager.py:14-292 restated on top of ``Store``."""
    def __init__(self, config: Dict):
        self.top_k = config['top_k']
        self.long_term = config['enable_long_term']
        self.count_long_usage = config['enable_long_term_count_usage']
        self.value_dim = config['value_dim']
        if self.long_term:
            self.max_frames = config['max_mid_term_frames']
            self.min_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long = config['max_long_term_elements']
        self.work = Store(keep_selection=self.long_term, keep_usage=self.long_term)
        self.long = Store(False, self.count_long_usage) if self.long_term else None
        self.sensory: Dict[int, torch.Tensor] = {}
        self.hw = None
        self.engaged = False

    # -- memory_manager.py:91-169 ----------------------------------------------------------
    def read(self, query_key: torch.Tensor, selection: torch.Tensor) -> Dict[int, torch.Tensor]:
        h, w = query_key.shape[-2:]
        qk = query_key[0].flatten(1)
        qe = selection[0].flatten(1)
        out = {}
        for b, rec in self.work.buckets.items():
            key, shr = rec.key, rec.shrinkage
            n_long = 0
            use_long = self.long_term and self.long.engaged(b)
            if use_long:
                lrec = self.long.buckets[b]
                n_long = lrec.size
                key = torch.cat([lrec.key, key], -1)
                shr = torch.cat([lrec.shrinkage, shr], -1)
            sim = mm.similarity(key, shr.reshape(-1), qk, qe)
            aff = mm.dense_affinity(sim, self.top_k)
            if self.long_term:
                u = mm.usage_of(aff)
                self.work.add_usage(b, u[n_long:])
                if use_long and self.count_long_usage:
                    self.long.add_usage(b, u[:n_long])
            for obj in rec.objects:
                v = self.work.values[obj]
                if use_long and obj in self.long.values:
                    v = torch.cat([self.long.values[obj], v], -1)
                out[obj] = mm.readout(aff, v).view(-1, h, w)
        return out

    # -- memory_manager.py:171-218 ---------------------------------------------------------
    def add(self, key, shrinkage, value, objects: List[int], selection=None) -> None:
        self.engaged = True
        self.hw = value.shape[-2] * value.shape[-1]
        key = key[0].flatten(1)
        shrinkage = shrinkage[0].flatten(1)
        value = value[0].flatten(2)
        if selection is not None:
            selection = selection[0].flatten(1)
        self.work.append(key, {o: value[i] for i, o in enumerate(objects)}, shrinkage, selection)
        if not self.long_term:
            return
        for b in list(self.work.buckets.keys()):
            if self.work.size(b) >= self.max_frames * self.hw:
                if self.long.size(b) >= self.max_long - self.num_prototypes:
                    self.long.evict_least_used(b, self.max_long - self.num_prototypes)
                self._consolidate(b)

    # -- memory_manager.py:231-276 ---------------------------------------------------------
    def _consolidate(self, b: int) -> None:
        hw = self.hw
        rec = self.work.buckets[b]
        lo, hi = hw, -self.min_frames * hw + hw
        sl = slice(lo, None) if hi == 0 else slice(lo, hi)
        cand_key, cand_shr, cand_sel = rec.key[:, sl], rec.shrinkage[:, sl], rec.selection[:, sl]
        cand_val = {o: self.work.values[o][:, sl] for o in rec.objects}
        usage = self.work.usage(b)[sl]
        _, top = torch.topk(usage, k=self.num_prototypes, dim=-1, sorted=True)
        proto_key, proto_sel = cand_key[:, top], cand_sel[:, top]
        sim = mm.similarity(cand_key, cand_shr.reshape(-1), proto_key, proto_sel)
        aff = mm.dense_affinity(sim, None)
        proto_val = {o: mm.readout(aff, v) for o, v in cand_val.items()}
        proto_shr = mm.readout(aff, cand_shr)
        self.work.drop_range(b, lo, hi, min_size=self.min_frames * hw + hw)
        self.long.append(proto_key, proto_val, proto_shr, None, forced_bucket=b)

    # -- memory_manager.py:220-229 ---------------------------------------------------------
    def keep_only(self, keep_ids) -> None:
        self.work.keep_only(keep_ids)
        if self.long_term and self.long.engaged():
            self.long.keep_only(keep_ids)
        self.sensory = {k: v for k, v in self.sensory.items() if k in keep_ids}
        if not self.work.engaged():
            self.engaged = False

    # -- memory_manager.py:278-292 ---------------------------------------------------------
    def sensory_for(self, ids: List[int], like: torch.Tensor) -> torch.Tensor:
        h, w = like.shape[-2:]
        for o in ids:
            if o not in self.sensory:
                self.sensory[o] = torch.zeros(self.value_dim, h, w)
        return torch.stack([self.sensory[o] for o in ids], 0).unsqueeze(0)

    def set_sensory(self, sensory: torch.Tensor, ids: List[int]) -> None:
        for i, o in enumerate(ids):
            self.sensory[o] = sensory[0, i]

    def sizes(self) -> Dict[int, tuple]:
        return {b: (self.work.size(b), self.long.size(b) if self.long_term else 0)
                for b in self.work.buckets}
