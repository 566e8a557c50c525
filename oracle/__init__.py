"""CPU oracle for the PCGCv2 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing under
``pcgcv2_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to anything here.

What it restates (every function cites the reference file:line it follows):

* ``sparse_ref``   -- the MinkowskiEngine operator semantics the reference's
  ``autoencoder.py`` relies on (coordinate maps, k=3 kernel maps, k3/k2s2/k1
  convolutions, generative transposed convolution, pruning).  MinkowskiEngine
  (">=0.5", un-pinned, ``README.md:19``) is a third-party dependency whose
  source is NOT under /root/reference and which is not installable here.
  **PARITY UNPINNED** for this part: the reference holds no tests or golden
  vectors; the conventions are pinned only behaviourally by the shipped
  checkpoints (SURVEY.md Appendix E.7 -- wrong conventions cost 4-19 dB) and
  by brute-force property tests in ``tests/``.
* ``rangecoder_ref`` + ``rc_ref.c`` -- torchac 0.9.3's float-CDF range coder
  (``entropy_model.py:174,192`` call sites; third-party, source absent).
  **PARITY UNPINNED** against real torchac bytes; pinned by encode->decode
  round trips and the ideal-code-length bound.
* ``entropy_ref``  -- ``entropy_model.py:82-196``.  **PINNED**: the reference
  module itself imports in the build container (with a stub ``torchac``) and
  ``tests/golden/make_golden.py`` froze its outputs on the r3/r7 checkpoints
  into ``tests/golden/entropy_*.npz``.
* ``codec_ref``    -- the encode/decode flow of ``coder.py:80-112`` over the
  topology of ``autoencoder.py:52-57,138-147,239-273``.
* ``metrics_ref``  -- D1 point-to-point PSNR as the bundled ``pc_error_d``
  computes it (``pc_error.py:44-54``), cross-checked against that binary in
  the build container.
"""
