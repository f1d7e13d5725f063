"""CPU oracle for the video->spike hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product path (``video-spike_b200/``) may import, call, link or
execute anything in this package.  The only permitted users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.

The reference (PPWangyc/video-spike) is pure Python on top of PyTorch; there is
no C/C++ source to compile, so there is no ``oracle/_ref`` binary.  The oracle
is a CPU restatement (torch CPU + numpy) of the reference algorithms, each
function citing the reference file:line it follows.

Pinning: the reference ships no tests or golden vectors.
The oracle is pinned instead against outputs of the reference's own modules,
imported unmodified from ``/root/reference/src`` by ``oracle/make_golden.py``
and committed as small fixtures under ``tests/golden/``
(``tests/test_oracle_golden.py`` replays them on every CPU test run).
"""
