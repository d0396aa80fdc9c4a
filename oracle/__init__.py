"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy) of the reference's heatmap-codec algorithms.  It is the
checker the CUDA path is compared against; it is never the thing shipped or
measured as the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``mindpose_b200/`` imports this package.

Parity status of each module (see DESIGN.md "Oracle"):

* ``topdown_encode``, ``affine``, ``grouping``, ``bottomup_encode`` restate the
  reference's numpy/cv2/scipy half.  They are PINNED: ``oracle/gen_golden.py``
  runs the unmodified reference functions imported from ``/root/reference``
  (``oracle/ref_loader.py``) on seeded inputs, stores the outputs under
  ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks the restatements
  against those vectors on every run.
* ``warp`` restates OpenCV's fixed-point ``warpAffine`` (third-party, pinned
  ``opencv-python<=4.5.4.60`` by the reference; 4.13.0 in this image).  PINNED
  against ``cv2.warpAffine`` outputs stored in ``tests/golden/``.
* ``lsap`` restates scipy's rectangular LSAP (third-party, ``scipy>=1.5.4``;
  1.18.1 in this image).  PINNED against scipy outputs in ``tests/golden/``.
* ``topdown_decode``, ``bottomup_decode`` restate MindSpore graphs that cannot
  be executed here (no ``mindspore``): PARITY UNPINNED -- the restatement is the
  oracle of record; every MindSpore-semantics assumption is listed in its
  docstring.
"""
