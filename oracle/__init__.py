"""TEST INFRASTRUCTURE ONLY — CPU oracles for the chunk mesher.

``oracle.binding`` wraps two shared libraries:
  * ``libgcoracle.so``      — plain-C restatement of the reference mesher (``gc_mesh_cpu.c``);
  * ``_ref/libgcref.so``    — the reference's own sources compiled verbatim (``ref/ref_harness.cpp``),
                               available only where ``/root/reference`` existed at build time.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package.
"""
