"""CPU oracle for the PVD volume-rendering hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker.  The product (``aaai2023-pvd_b200/``) never imports it
and has no CPU fallback.

Contents
--------
``pvd_oracle.c``   plain-C restatement of the reference CUDA kernels (cited per function)
``cpu.py``         ctypes/numpy front-end to it (``build()`` compiles it with gcc)
``field.py``       torch-CPU restatement of the field networks (ATen ops are their own oracle)
``sh_reference.py``independent double-precision real-SH construction (associated Legendre recurrences)
``ref_glue.py``    restatement of the reference's Python glue around ``oracle/_ref`` (GPU oracle)
``build_ref.py``   recipe that compiles the unmodified reference extensions into ``oracle/_ref/``

Parity pinning: the reference has no tests, golden vectors or fixtures of its own.  The oracle is
pinned against outputs of the reference's kernels themselves, run on a B200 from ``oracle/_ref`` and committed
under ``tests/golden/`` with the generating script.
"""
