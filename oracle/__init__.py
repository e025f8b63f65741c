"""CPU oracle for the opponent-critique path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path
(``adversarial-spec_b200/``) never does and fails loudly without its CUDA
library.  See DESIGN.md "Oracle".
"""
