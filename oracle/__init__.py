"""CPU oracle for the simple_dqn hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import anything from this package.  The product (``simple_dqn_b200``) never
imports it and fails loudly when its CUDA library is missing.

Pinning status:
  * replay half (``mt19937``, ``replay_oracle``): PINNED — checked against the
    unmodified reference file ``src/replay_memory.py`` / ``src/state_buffer.py``
    imported in the build container, and against CPython's own ``random`` module;
    golden vectors under ``tests/golden/`` were produced by that reference code.
  * network half (``dqn_oracle``): PARITY UNPINNED against Neon — the arithmetic
    lives in NervanaSystems/neon (unpinned dependency; snapshots record
    1.3.0+344372b), which is absent from /root/reference and cannot be installed.
    The restatement follows Neon's published algorithm at the call sites in
    ``src/deepqnetwork.py`` and is cross-checked against an independent
    torch-CPU autograd implementation and the shipped snapshot weights.
"""
