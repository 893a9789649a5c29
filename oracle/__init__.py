"""TEST INFRASTRUCTURE ONLY -- the parity oracle for the omr_a2s_multimodal_transformer hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product package
``omr_a2s_multimodal_transformer_b200`` never imports from here and fails loudly when its CUDA
library is missing.

Contents
--------
``restate.py``   CPU restatement (plain torch CPU ops, fp32 or fp64) of the reference's
                 encoder / 2-D PE / mixers / decoder / CE / greedy loop, each function citing the
                 reference file:line it follows.
``synth.py``     deterministic synthetic weights (keyed by state-dict name, independent of module
                 construction order) and synthetic GrandStaff-shaped batches.
``shim.py``      loader for the REAL reference at /root/reference (stubbing the packages missing
                 from this image).  Only usable in the build container; used to pin ``restate.py``
                 and to generate ``tests/golden``.
``make_golden.py`` script that produced ``tests/golden/*.pt`` from the real reference.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container
by ``make_golden.py`` (committed fixtures) and re-checked live by ``tests/test_oracle_pin.py``
whenever /root/reference is present.
"""
