"""CPU oracle for the block-bordered (Schur-complement) KKT solve.

TEST INFRASTRUCTURE ONLY.  Nothing under ``parapint_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline -- never as the thing shipped.

Contents
--------
``schur_oracle``      numpy/scipy restatement of the reference's serial and
                      rank-partitioned Schur-complement algorithm with SciPy
                      (SuperLU) leaves.
``kkt_generator``     restatement of the reference's seeded synthetic
                      block-bordered KKT generator (family G).
``reference_loader``  loads the UNMODIFIED reference files from
                      ``/root/reference`` against PyNumero stand-ins; only
                      usable in the build container (the reference does not
                      travel to the GPU box) and only used to pin the
                      restatement and to emit ``tests/golden`` fixtures.
``ipm``               flat restatement of the interior-point loop and of the
                      stochastic / dynamic interface layouts (farmer LP,
                      dynamics QP, separable two-stage QPs).  UNPINNED against
                      ``ip_solve`` itself (pyomo is absent): pinned only by the
                      reference's end values (farmer acreage, the nine control
                      values of the dynamics example) and the layout tests.
``kkt_families``      synthetic IPM-shaped KKT systems at the shapes of BASELINE
                      configs 3 and 4 (family P); nothing to pin.
``parallel_baseline`` the reference algorithm partitioned over the host cores
                      (process per "rank"): the CPU arm of ``bench.py``.

Parity status of ``schur_oracle`` / ``kkt_generator``: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified
reference solver here and committed its outputs; ``tests/test_oracle.py`` checks
this restatement against those fixtures and against the reference's own golden
values (8x8 known-answer system, 3x3 leaf system, max_err 0.3163456780448639).
The arithmetic of the leaves is SciPy's SuperLU / LAPACK (third-party, scipy
1.18.1 here, unpinned in reference ``setup.py:14``).  MA27/MUMPS-specific
behaviour (pivot order, ``cntl(1)`` threshold pivoting) is absent from this
image: parity UNPINNED for that leaf family.
"""
