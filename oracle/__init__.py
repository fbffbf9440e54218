"""CPU oracle for the two-stage detection glue.

TEST INFRASTRUCTURE ONLY.  Nothing under ``faster_rcnn_pytorch_multimodal_b200/``
may import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only as
the checker (or as the timed CPU arm), never as the shipped path.

Pinning status (see DESIGN.md "Oracle"): the reference ships no tests or golden
vectors.  The restatement in ``glue_oracle.py`` is pinned by
``tests/golden/*.npz``, which ``oracle/gen_golden.py`` produced by importing and
running the *unmodified* reference functions from ``/root/reference/lib`` in the
build container, plus the two known-answer vectors the reference carries
(9 canonical anchors, 3 rotated BEV boxes).
"""
