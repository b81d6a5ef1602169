"""CPU oracle for the DEVA temporal-propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product package (``tracking-anything-with-deva_b200/deva``)
never imports this package and fails loudly when its CUDA library is missing.

The oracle is a plain PyTorch-CPU fp32 restatement of the reference algorithm
(hkchengrex/Tracking-Anything-with-DEVA @ 404a112), written as flat functions over a
checkpoint ``state_dict`` instead of the reference's ``nn.Module`` tree.  Every function
cites the reference file:line it restates.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (reference imported read-only from /root/reference with
three import shims) and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py``
replays every fixture through this package.
"""
