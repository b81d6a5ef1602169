"""Oracle: XMem-style memory reading math (test infrastructure, see oracle/__init__.py).

Restates deva/model/memory_utils.py:6-94 and deva/inference/memory_manager.py:64-75
of the reference as un-batched 2-D functions.

Layouts (all channel-major, exactly as the reference keeps them):
    mem_key   [CK, N]   memory keys
    mem_shr   [N]       memory shrinkage (>= 1)
    qry_key   [CK, Q]   query keys
    qry_sel   [CK, Q]   query selection in (0, 1)
    values    [R, N]    value rows (R = num_objects * CV)
"""
import math
from typing import Optional, Tuple

import torch


def similarity(mem_key: torch.Tensor, mem_shr: Optional[torch.Tensor], qry_key: torch.Tensor,
               qry_sel: Optional[torch.Tensor]) -> torch.Tensor:
    """Anisotropic L2 similarity [N, Q].  Reference: memory_utils.py:6-45.

    sim[n, q] = -(shr[n] / sqrt(CK)) * sum_c sel[c, q] * (key[c, n] - qk[c, q])^2,
    evaluated in the reference's expanded order (a^2 term, 2ab term, b^2 term).
    """
    ck = mem_key.shape[0]
    mk_t = mem_key.t()  # [N, CK]
    if qry_sel is not None:
        a_sq = (mk_t * mk_t) @ qry_sel  # memory_utils.py:30
        two_ab = 2 * (mk_t @ (qry_key * qry_sel))  # :31
        b_sq = (qry_sel * qry_key * qry_key).sum(0, keepdim=True)  # :32
        sim = -a_sq + two_ab - b_sq  # :33
    else:
        a_sq = (mem_key * mem_key).sum(0).unsqueeze(1)  # :36
        sim = -a_sq + 2 * (mk_t @ qry_key)  # :37-38
    if mem_shr is not None:
        sim = sim * mem_shr.reshape(-1, 1) / math.sqrt(ck)  # :41
    else:
        sim = sim / math.sqrt(ck)  # :43
    return sim


def topk_softmax(sim: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-query (column) top-k then softmax over the k kept entries.

    Returns (idx [k, Q] int64, weight [k, Q]) sorted by descending similarity.
    Reference: memory_utils.py:56-60.  The reference exponentiates without subtracting the
    maximum; we subtract it, which is identical wherever the reference is finite
    (SURVEY.md section 8 quirk Q3).
    """
    vals, idx = torch.topk(sim, k=top_k, dim=0)
    e = torch.exp(vals - vals[0:1])
    return idx, e / e.sum(0, keepdim=True)


def dense_affinity(sim: torch.Tensor, top_k: Optional[int]) -> torch.Tensor:
    """Dense affinity [N, Q] as the reference materialises it.  memory_utils.py:48-71."""
    if top_k is None:
        e = torch.exp(sim - sim.max(0, keepdim=True)[0])  # :67-70
        return e / e.sum(0, keepdim=True)
    idx, w = topk_softmax(sim, top_k)
    return torch.zeros_like(sim).scatter_(0, idx, w)  # :62-65


def usage_of(affinity: torch.Tensor) -> torch.Tensor:
    """Per-slot usage = row sums of the dense affinity.  memory_utils.py:73-74."""
    return affinity.sum(1)


def readout(affinity: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    """values [R, N] @ affinity [N, Q] -> [R, Q].  memory_manager.py:64-75."""
    return values @ affinity


def read(mem_key, mem_shr, qry_key, qry_sel, values, top_k):
    """Whole read: returns (readout [R, Q], usage [N], idx [k, Q], weight [k, Q])."""
    sim = similarity(mem_key, mem_shr, qry_key, qry_sel)
    idx, w = topk_softmax(sim, top_k)
    aff = torch.zeros_like(sim).scatter_(0, idx, w)
    return readout(aff, values), usage_of(aff), idx, w
