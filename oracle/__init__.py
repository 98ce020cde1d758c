"""CPU oracle for the jQMC walker hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``jqmc_b200/`` imports it; the
product path fails loudly when the CUDA library is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):
* physics (AO/MO/geminal/Jastrow/Coulomb/ECP/kinetic/local energy): PINNED against the TurboRVB
  known-answer numbers hard-coded in the reference's own tests
  (tests/test_comparison_with_turborvb_ECP.py:105-129, :231-274, ...), see tests/test_oracle_golden.py.
* RNG stream (Threefry-2x32 ``jax.random`` restatement): the Threefry block function is pinned
  against the Random123 known-answer vectors; the derived ``split/randint/uniform/normal`` stream is
  **parity unpinned** -- JAX is not installed in the build image and the reference holds no
  known-answer test for its random stream.
"""
