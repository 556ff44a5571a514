"""CPU oracle for the RFF maximiser-sampling + MPES hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product package never imports this module and fails
loudly when its CUDA library is missing.

Contents
--------
``rff_oracle``   float64 NumPy restatement of ``fourier_features.py`` (ff.py) —
                 every function cites the reference file:line it follows.
``mpes_oracle``  float64 NumPy restatement of ``acquisitions/pes.py``,
                 ``acquisitions/rank_pes.py`` and the tail of ``acquisitions/dts.py``.
``ref_loader``   loads the *unmodified* reference sources from ``/root/reference``
                 (when that directory exists, i.e. in the build container only):
                 ``rank_pes.py`` imports as-is (NumPy/SciPy only); ``fourier_features.py``,
                 ``pes.py`` and ``dts.py`` execute under ``tf_shim`` (a NumPy stand-in for the
                 handful of TensorFlow / TFP / GPflow symbols those files touch, because
                 TF 2.1 / TFP 0.9 / GPflow 2.0.0-rc1 cannot be installed offline).
``tf_shim``      the stand-in itself.

Parity pinning: the reference's own tests hold no vector for this path
(``test.py`` covers ``objectives.py`` only), so the restatement is pinned against
outputs of the reference sources executed here (``tests/golden/make_golden.py`` ->
``tests/golden/*.npz``, committed) and re-checked live whenever ``/root/reference``
is present.
"""
