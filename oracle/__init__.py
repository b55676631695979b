"""CPU oracle for the FITS-reduction hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``astrophotography_b200`` imports this package.  The only callers
are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``, and there only as the checker or as
the timed CPU baseline -- never as the product path.

Contents
--------
``calibrate_oracle``  numpy restatement of ``ApCalibrate`` arithmetic
                      (reference ``AstroPhotography/core/ApCalibrate.py:166-190,439-474``).
                      PINNED: checked against the reference source executed
                      verbatim (``ref_exec``) -> ``tests/golden/calibrate_*.npz``.
``badpix_oracle``     numpy restatement of ``ApFixBadPixels.fix_bad_pixels``
                      (``core/ApFixBadPixels.py:292-445``) and of the
                      ``ApFindBadPixels`` mask rules (``core/ApFindBadPixels.py:70-217``).
                      PINNED the same way -> ``tests/golden/badpix_*.npz``.
``combine_oracle``    float64 restatement of ``ccdproc.combine`` as configured at
                      ``scripts/ap_combine_darks.py:394-420``.  The arithmetic lives
                      in third-party ccdproc (>=2.1, effectively 2.4.x) and astropy
                      (>=6.0), neither of which is vendored in the reference or
                      installable here, and the reference holds no test or golden
                      vector for it:  **PARITY UNPINNED** for this stage.  The
                      restatement follows the published algorithm
                      (Combiner.sigma_clipping -> astropy.stats.sigma_clip,
                      average_combine / median_combine) and is anchored on
                      hand-computable known-answer cases in ``tests/golden``.
``ref_exec``          loader that executes the reference's own
                      ``core/ApCalibrate.py`` / ``core/ApFixBadPixels.py`` /
                      ``core/ApFindBadPixels.py`` verbatim from ``/root/reference``
                      behind a fake in-memory ``astropy.io.fits``.  Only usable where
                      ``/root/reference`` exists (the authoring container); used by
                      ``oracle/make_golden.py`` to mint the committed fixtures.
"""
