"""GPU, world_size 2 on ONE device: the real CUDA backend under the multi-rank control flow.

Two processes share cuda:0 and all-reduce their device buffers through gloo (NCCL refuses two ranks on one GPU),
so the path several GPUs take -- no host synchronisation in the local phase, status / overflow flag / inertia in
the tail of the Schur all-reduce, the coupling phase reading that tail, the residual exchange -- is exercised
by the GPU test tier on a single B200.  Mirrors reference test_mpi_explicit_schur_complement.py:19-115.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.kkt_generator import EstimationModel
        from oracle.schur_oracle import dense_inertia, solve_partitioned, sym_full
        from parapint_b200 import B200SchurComplementLinearSolver, Communicator, LinearSolverStatus
        from tests.helpers import block_vector, bordered_from_dense, stochastic_ipm_system

        comm = Communicator()
        if case == "generator":
            full = EstimationModel(6, 40, 3, 8)
            local = EstimationModel(6, 40, 3, 8, local_blocks=[i for i in range(6) if i % world == rank])
            kkt, rhs = local.build_kkt(), local.build_rhs()
            solver = B200SchurComplementLinearSolver(comm=comm)
            assert solver._defer == 2
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            for _ in range(2):
                assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
                x = solver.do_back_solve(rhs)
            _, x_ref, _ = solve_partitioned(full.build_kkt(), full.build_rhs(), world)
            for i in local.local_blocks:
                assert np.allclose(x.get_block(i), x_ref.get_block(i), rtol=1e-9, atol=1e-9)
            assert np.allclose(x.get_block(6), x_ref.get_block(6), rtol=1e-9, atol=1e-9)
            assert solver.get_inertia() == full.expected_inertia()
            assert solver.last_residual is not None and solver.last_residual <= 1e-10
        elif case == "overflow":
            # no delayed-pivot capacity: the sparse path of some rank overflows, every rank learns it from the reduced
            # tail and repeats its local phase (the overflowing one densely)
            kkt, sizes = stochastic_ipm_system(3, 4, 300, 240, 30, 10)
            rhs = block_vector(np.random.default_rng(3).standard_normal(sum(sizes)), sizes)
            solver = B200SchurComplementLinearSolver(comm=comm, options={"sparse_dmax": 0})
            solver.do_symbolic_factorization(kkt)
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
            x = solver.do_back_solve(rhs)
            dense = sym_full(kkt).toarray()
            x_ref = np.linalg.solve(dense, rhs.flatten())
            off = np.concatenate(([0], np.cumsum(sizes)))
            for i in list(solver.local_block_indices) + [len(sizes) - 1]:
                assert np.allclose(np.asarray(x.get_block(i)).ravel(), x_ref[off[i]:off[i + 1]], rtol=1e-7, atol=1e-9)
            assert solver.get_inertia() == dense_inertia(dense, "ldl")
        elif case == "singular":
            dense = np.zeros((5, 5))
            dense[:2, :2] = [[1.0, 2.0], [2.0, 4.0]]  # block 0 (rank 0) is singular
            dense[2:4, 2:4] = np.eye(2)
            dense[4, 4] = 1.0
            dense[4, 0] = dense[0, 4] = 1.0
            kkt = bordered_from_dense(dense, [2, 2, 1])
            solver = B200SchurComplementLinearSolver(comm=comm)
            solver.do_symbolic_factorization(kkt)
            res = solver.do_numeric_factorization(kkt, raise_on_error=False)
            assert res.status == LinearSolverStatus.singular  # every rank agrees (mpi...:19-30)
        elif case == "sparse_coupling":
            # time-decomposed layout, blocks dealt over two ranks: the pattern of S needs the borders of BOTH ranks'
            # blocks (all-gathered in the symbolic phase), the all-reduce moves the pattern values only, every rank
            # factorises S level by level and gets the same coupling solution
            from tests.helpers import dynamic_ipm_system
            N = 21
            full, sizes = dynamic_ipm_system(6, N, 60, 30, 5, 4)
            kkt, _ = dynamic_ipm_system(6, N, 60, 30, 5, 4, local_blocks=[i for i in range(N) if i % world == rank])
            rhs = block_vector(np.random.default_rng(6).standard_normal(sum(sizes)), sizes)
            solver = B200SchurComplementLinearSolver(comm=comm, options={"coupling_min_sparse": 16})
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            cs = solver.backend.coupling_stats()
            assert cs["levels"] >= 1 and solver.backend.schur_size == cs["schur_size"] < cs["m_c"] ** 2 // 2
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
            x = solver.do_back_solve(rhs)
            dense = sym_full(full).toarray()
            x_ref = np.linalg.solve(dense, rhs.flatten())
            off = np.concatenate(([0], np.cumsum(sizes)))
            for i in list(solver.local_block_indices) + [N]:
                assert np.allclose(np.asarray(x.get_block(i)).ravel(), x_ref[off[i]:off[i + 1]], rtol=1e-7, atol=1e-8)
            assert solver.get_inertia() == dense_inertia(dense, "ldl")
            assert solver.last_residual is not None and solver.last_residual <= 1e-10
        elif case == "device_regularization":
            # shifts on the device under the multi-rank control flow: the retry re-uses the values on every rank
            from oracle.ipm import StochasticInterface, device_regularized, random_stochastic_qp
            from oracle.schur_oracle import SchurOracle
            scen, fs = random_stochastic_qp(1, 4, 40, 14, 6, 4, 0.3)
            itf = device_regularized(StochasticInterface)(scen, fs)
            itf.set_barrier_parameter(0.1)
            for s_ in itf.sc:
                s_.nlp.x = np.full(s_.nlp.n, 0.3)
            kkt, rhs = itf.evaluate_primal_dual_kkt_matrix(), itf.evaluate_primal_dual_kkt_rhs()
            solver = B200SchurComplementLinearSolver(comm=comm, regularization_classes=itf.regularization_classes())
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            assert solver.do_numeric_factorization(kkt, raise_on_error=False).status in (LinearSolverStatus.successful, LinearSolverStatus.singular)
            reg = itf.regularize_hessian(itf.regularize_equality_gradient(kkt.copy(), -1e-2, False), 1e-2, False)
            assert solver.do_numeric_factorization(reg).status == LinearSolverStatus.successful
            x = solver.do_back_solve(rhs)
            o = SchurOracle(compute_inertia=True, inertia_method="ldl")
            full = reg.materialize()
            o.symbolic(full)
            assert o.numeric(full) == 0
            x_ref = o.solve(rhs)
            for i in list(solver.local_block_indices) + [4]:
                assert np.allclose(np.asarray(x.get_block(i)).flatten(), np.asarray(x_ref.get_block(i)).flatten(), rtol=1e-7, atol=1e-9)
            assert solver.get_inertia() == o.inertia()
            assert solver.symbolic_calls == 1 and solver.backend.value_uploads() == 1
        elif case == "ipm_vectors":
            # the interior-point vector kernels on distributed iterates: every rank holds half of the entries, the step
            # lengths / maxima / sums are reduced over the ranks (the reference's MPIBlockVector reductions)
            from oracle import ipm as O
            from parapint_b200.ipm_vectors import DeviceIpmVectors
            from tests.test_gpu_ipm_vectors import _random_group, _ref_ftb
            rng = np.random.default_rng(9)
            gx, gs = _random_group(rng, 4001), _random_group(rng, 1203)       # (x, dx, lb, ub, zl, zu) of primals / slacks
            lam_eq, lam_in = rng.standard_normal(777), rng.standard_normal(1203)
            tau, barrier = 0.99, 2e-3
            mine = lambda v: v[rank::world].copy()                            # noqa: E731 - this rank's entries
            dv = DeviceIpmVectors(comm=comm)
            names = (("primals", "delta_primals", "primals_lb", "primals_ub", "duals_primals_lb", "duals_primals_ub"),
                     ("slacks", "delta_slacks", "ineq_lb", "ineq_ub", "duals_slacks_lb", "duals_slacks_ub"))
            for grp, nm in zip((gx, gs), names):
                for v, name in zip(grp, nm):
                    dv.v[name] = dv.k.to_device(mine(v))
            dv.v["duals_eq"], dv.v["duals_ineq"] = dv.k.to_device(mine(lam_eq)), dv.k.to_device(mine(lam_in))
            a_p, a_d = dv.fraction_to_the_boundary(tau, barrier)
            with np.errstate(all="ignore"):
                rx, rs = _ref_ftb(tau, barrier, *gx), _ref_ftb(tau, barrier, *gs)
            assert a_p == min(rx[0], rs[0], 1.0) and a_d == min(rx[1], rs[1], 1.0)          # bit-exact on every rank
            c_inf, dual_scaling, compl_scaling = dv.complementarity(barrier, 100.0)

            def resid(x, lb, ub, zl, zu):
                lb_m, ub_m = np.where(np.isneginf(lb), 0.0, lb), np.where(np.isinf(ub), 0.0, ub)
                r_l, r_u = (x - lb_m) * zl - barrier, (ub_m - x) * zu - barrier
                r_l[np.isneginf(lb)] = 0
                r_u[np.isinf(ub)] = 0
                return max(O._max_abs(r_l), O._max_abs(r_u))
            assert c_inf == max(resid(gx[0], gx[2], gx[3], gx[4], gx[5]), resid(gs[0], gs[2], gs[3], gs[4], gs[5]))
            bsum = sum(np.abs(g[4]).sum() + np.abs(g[5]).sum() for g in (gx, gs))
            nb = sum(np.isfinite(g[2]).sum() + np.isfinite(g[3]).sum() for g in (gx, gs))
            ds = (np.abs(lam_eq).sum() + np.abs(lam_in).sum() + bsum) / (lam_eq.size + lam_in.size + nb)
            assert np.isclose(dual_scaling, max(100.0, ds) / 100.0, rtol=1e-13)
            assert np.isclose(compl_scaling, max(100.0, bsum / nb) / 100.0, rtol=1e-13)
        open(os.path.join(out_dir, f"ok_{case}_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["generator", "overflow", "singular", "sparse_coupling", "device_regularization",
                                  "ipm_vectors"])
def test_two_ranks_on_one_gpu(tmp_path, case):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok_{case}_0", f"ok_{case}_1"]
