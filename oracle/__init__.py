"""CPU oracle for the surface-projection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tissue_image_processing_b200/`` may
import this package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker
and as the timed CPU baseline, never as the product path.

Parity pinning: the reference repository ships no tests with assertions and no
golden vectors (reference ``tissue_analyzing_tool/Tests.py`` only plots), so the
oracle is pinned against the *reference itself*, imported unmodified in the build
container by ``oracle/reference_runner.py`` (I/O dependencies stubbed), and the
outputs are frozen under ``tests/golden/`` by ``oracle/make_golden.py``.
``bin_size > 1`` depends on scikit-image, which is neither vendored nor pinned by
the reference and is absent here: that branch is **parity unpinned**.
"""
