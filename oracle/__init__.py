"""CPU oracle for the Parasitoids drift-diffusion forward solve.

TEST INFRASTRUCTURE ONLY.  Nothing under ``parasitoids_b200/`` imports this
package.  The only legitimate importers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- always as the checker or as the CPU baseline being
timed, never as the product path.

Contents
--------
pm_oracle   numpy restatement of ParasitoidModel.py:231-613 (flight
            probability, Genz BVN cell masses, per-day kernel ``prob_mass``).
cs_oracle   numpy/scipy restatement of CalcSol.py:11-324 (FFT convolution
            chain, threshold/renormalise, back_solve, population model).
ref_loader  imports the *unmodified* reference modules from /root/reference
            with the 3-line ``mvn.mvnun`` shim (build-container only).
make_golden generates tests/golden/*.npz from the reference itself.

Parity pinning: pinned.  The restatement is checked (tests/test_oracle.py)
against golden vectors produced by running the reference's own code in the
build container (oracle/make_golden.py, vectors committed under
tests/golden/), and against the reference's own known-answer tests
(tests/test_CalcSol.py:75-139 in the reference tree, restated in
tests/test_oracle.py::test_reference_known_answers and tests/test_chain.py).
"""
