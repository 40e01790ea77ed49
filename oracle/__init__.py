"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).

Nothing under ``fluidsim_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` use it, and there only as the checker / the CPU baseline.

Contents
--------
``fluidfft_np``   numpy/scipy restatement of the third-party layer that is NOT in
                  /root/reference: fluidfft 0.4.2 (``pdm.lock:920-921``) FFT plugin
                  class + ``OperatorsPseudoSpectral{2,3}D`` base + ``vector_product``
                  and fluiddyn 0.6.6 ``SetOfVariables``.  Call-site contracts:
                  ``fluidsim/solvers/ns3d/solver.py:199-241``,
                  ``fluidsim/operators/operators3d.py:222-231``.
``step_np``       self-contained numpy restatement of the reference's own hot path
                  (RK2/RK4 ``base/time_stepping/pseudo_spect.py:469-517,798-984``;
                  ``tendencies_nonlin`` of ns3d / ns3d.strat / ns2d).  Travels to the
                  GPU box (no /root/reference there).
``refshim``       stub modules that let the *unmodified* reference modules be imported
                  from /root/reference in the build container; used by
                  ``tests/golden/make_golden.py`` and ``tests/test_oracle_vs_reference.py``
                  to pin ``step_np`` against the reference's own code.

Parity status
-------------
The reference tree stores no golden vectors for this path.
``step_np`` is pinned against the reference's own Python executed here through
``refshim`` (committed fixtures in ``tests/golden`` + generating script).  The
fluidfft layer itself is absent from /root/reference and cannot be installed
offline, so its restatement is pinned only by the reference's call-site contracts
and identity tests: **parity of the fluidfft layer is unpinned** (in particular the
comparator of the default "cubic" dealiasing mask; the CUDA path therefore takes
the mask as an input array instead of re-deriving it).
"""
