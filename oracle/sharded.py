"""Oracle: the bank-sharded read and consolidation softmax as a collective protocol (test infrastructure, see
oracle/__init__.py).

The reference has no multi-GPU read; what is restated here is the PROTOCOL of the product's bank-sharded mode
(tracking-anything-with-deva_b200/deva/inference/sharded_core.py:137-218 ``ShardedMemoryManager.match_memory`` and :246-298
``consolidation``) on top of the reference's own math (oracle/memory_math.py = deva/model/memory_utils.py:6-76,
deva/inference/memory_manager.py:64-75, :251-276), with torch.distributed collectives on CPU tensors, so that a gloo run at
any world size can be compared with the unsharded oracle:

  1. every rank: similarity of the queries to ITS slots, local top-k (value, global slot id);
  2. all-gather of the candidate lists;
  3. every rank: global top-k of the R*k candidates (ties -> lower slot id, like the merge kernel) + softmax;
  4. every rank: partial read-out of ALL objects over the selected slots it owns, usage of its own slots;
  5. reduce-scatter by object: rank r ends with the complete read-out of objects [r*per, (r+1)*per).
"""
from typing import Tuple

import torch
import torch.distributed as dist

from oracle import memory_math as mm


def token_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """sharded_core.py:38-42 (sizes differ by <= 1)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_read(mem_key_loc, mem_shr_loc, values_loc, slot_offset: int, qry_key, qry_sel, top_k: int, num_objects: int,
                 group=None):
    """This rank's slice (keys [CK, n_loc], shrinkage [n_loc], values [K*CV, n_loc], global id of its first slot) ->
    (read-out [per*CV, Q] of the objects this rank owns, usage [n_loc] of its own slots, global idx [k, Q], weights [k, Q])."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    q = qry_key.shape[1]
    n_loc = mem_key_loc.shape[1]
    k_loc = min(top_k, n_loc)
    # 1. local candidates, padded to top_k with (-inf, -1)
    vals = torch.full((top_k, q), float('-inf'), dtype=qry_key.dtype)
    ids = torch.full((top_k, q), -1, dtype=torch.int64)
    if k_loc > 0:
        sim = mm.similarity(mem_key_loc, mem_shr_loc, qry_key, qry_sel)  # [n_loc, Q]
        v, i = torch.topk(sim, k=k_loc, dim=0)
        vals[:k_loc], ids[:k_loc] = v, i + slot_offset
    # 2. all-gather
    all_v = [torch.empty_like(vals) for _ in range(world)]
    all_i = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(all_v, vals, group=group)
    dist.all_gather(all_i, ids, group=group)
    cand_v, cand_i = torch.cat(all_v, 0), torch.cat(all_i, 0)  # [R*k, Q]
    # 3. global top-k: value descending, ties -> lower slot id; empty entries (-1) last
    key_i = torch.where(cand_i >= 0, cand_i, torch.full_like(cand_i, 2**62))
    order = torch.argsort(key_i, dim=0, stable=True)
    cand_v, cand_i = torch.gather(cand_v, 0, order), torch.gather(cand_i, 0, order)
    order = torch.argsort(cand_v, dim=0, descending=True, stable=True)[:top_k]
    sel_v, sel_i = torch.gather(cand_v, 0, order), torch.gather(cand_i, 0, order)
    e = torch.exp(sel_v - sel_v[0:1])
    w = e / e.sum(0, keepdim=True)
    # 4. partial read-out over MY slots, usage of my slots
    mine = (sel_i >= slot_offset) & (sel_i < slot_offset + n_loc)
    aff = torch.zeros(max(n_loc, 1), q, dtype=qry_key.dtype)
    aff.scatter_add_(0, torch.where(mine, sel_i - slot_offset, torch.zeros_like(sel_i)), torch.where(mine, w, torch.zeros_like(w)))
    aff = aff[:n_loc]
    partial = values_loc @ aff  # [K*CV, Q]
    usage = aff.sum(1)
    # 5. reduce-scatter by object (blocks of ceil(K / R) objects; padded so every rank contributes equal chunks)
    cv = values_loc.shape[0] // num_objects
    per = -(-num_objects // world)
    padded = torch.zeros(world * per * cv, q, dtype=partial.dtype)
    padded[:partial.shape[0]] = partial
    dist.all_reduce(padded, group=group)  # gloo has no reduce_scatter; the slice below is what the owner keeps
    lo, hi = min(num_objects, rank * per), min(num_objects, (rank + 1) * per)
    return padded[lo * cv:hi * cv], usage, sel_i, w


def sharded_row_softmax(sim_loc: torch.Tensor, group=None) -> torch.Tensor:
    """Softmax over ALL candidates of every prototype row when each rank holds a column slice [P, n_loc] of the similarity:
    row max and row sum all-reduced over the ranks (sharded_core.py:280-284; memory_utils.py:66-71 evaluated distributedly)."""
    m = sim_loc.max(dim=1, keepdim=True)[0] if sim_loc.shape[1] else torch.full((sim_loc.shape[0], 1), float('-inf'))
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    e = torch.exp(sim_loc - m)
    z = e.sum(dim=1, keepdim=True)
    dist.all_reduce(z, group=group)
    return e / z
