"""CPU, world_size 2 (gloo): the multi-rank host path -- partition, all-reduces, status agreement.

The kernels are replaced by the numpy stand-in of the C ABI (tests/fake_backend.py); what is under
test is the solver's own distributed logic, which is identical on NCCL.
Mirrors reference test_mpi_explicit_schur_complement.py:19-115 (run there under mpirun -np 2/3).
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


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.kkt_generator import EstimationModel
        from oracle.schur_oracle import solve_partitioned
        from parapint_b200 import B200SchurComplementLinearSolver, Communicator, LinearSolverStatus
        from tests.fake_backend import FakeBackend
        from tests.helpers import block_vector, bordered_from_dense

        comm = Communicator()
        assert (comm.rank, comm.size) == (rank, world)
        solver = B200SchurComplementLinearSolver(backend=FakeBackend(), comm=comm)
        if case == "generator":
            full = EstimationModel(5, 12, 2, 3)
            local = EstimationModel(5, 12, 2, 3, local_blocks=[i for i in range(5) if i % world == rank])
            kkt, rhs = local.build_kkt(), local.build_rhs()
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            assert solver.local_block_indices == local.local_blocks
            for _ in range(2):  # refactor + re-solve reuse (test_mpi...:113-115)
                assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
                x = solver.do_back_solve(rhs)
            st, x_ref, _ = solve_partitioned(full.build_kkt(), full.build_rhs(), world)
            for i in local.local_blocks:
                assert np.allclose(x.get_block(i), x_ref.get_block(i), rtol=1e-9, atol=1e-9)
            assert np.allclose(x.get_block(5), x_ref.get_block(5), rtol=1e-9, atol=1e-9)  # replicated coupling
            assert solver.get_inertia() == full.expected_inertia()
            assert abs(full.check_result(x_ref) - local.check_result(x)) < 1e-6 or world > 1
        elif case == "singular":
            dense = np.zeros((5, 5))
            dense[:2, :2] = [[1.0, 2.0], [2.0, 4.0]]  # block 0 (rank 0) is singular
            dense[2:4, 2:4] = np.eye(2)
            dense[4, 4] = 1.0
            dense[4, 0] = dense[0, 4] = 1.0
            kkt = bordered_from_dense(dense, [2, 2, 1])
            solver.do_symbolic_factorization(kkt)
            res = solver.do_numeric_factorization(kkt, raise_on_error=False)
            assert res.status == LinearSolverStatus.singular  # every rank agrees (mpi...:19-30)
            try:
                solver.do_numeric_factorization(kkt, raise_on_error=True)
                raise AssertionError("expected RuntimeError on every rank")
            except RuntimeError:
                pass
        elif case == "cliques":
            # the pattern of a sparse Schur complement needs the nonzero border rows of the blocks of EVERY rank
            # (mpi_explicit_schur_complement.py:244-247): every rank hands the same global list to the native symbolic phase
            from tests.helpers import dynamic_ipm_system
            N = 7
            full, sizes = dynamic_ipm_system(2, N, 30, 12, 3, 2)
            kkt, _ = dynamic_ipm_system(2, N, 30, 12, 3, 2, local_blocks=[i for i in range(N) if i % world == rank])
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            ptr, rows = solver.backend.cliques
            assert len(ptr) == N + 1 and ptr[-1] == rows.size
            got = sorted(tuple(rows[ptr[k]:ptr[k + 1]]) for k in range(N))
            want = sorted(tuple(np.unique(full.get_block(N, i).tocoo().row)) for i in range(N))
            assert got == want
        elif case == "rank_local_failure":
            # a run-time failure on ONE rank must become the same status on EVERY rank (no hang, no
            # exception on one rank only), in the symbolic phase and in the numeric phase.
            local = EstimationModel(4, 8, 2, 3, local_blocks=[i for i in range(4) if i % world == rank])
            kkt = local.build_kkt()
            be = solver.backend
            be.fail_symbolic = rank == 1
            res = solver.do_symbolic_factorization(kkt, raise_on_error=False)
            assert res.status == LinearSolverStatus.error
            be.fail_symbolic = False
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            be.fail_numeric = rank == 0
            res = solver.do_numeric_factorization(kkt, raise_on_error=False)
            assert res.status == LinearSolverStatus.error
            be.fail_numeric = False
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
            # warning (4) on one rank must not mask singular (2) on another: severity order, not enum order
            t = solver._finish(4 if rank == 0 else 2, False, "x")
            assert t.status == LinearSolverStatus.singular
        elif case == "pattern_change_one_rank":
            # the COO pattern of ONE rank's block changes between factorisations: every rank repeats the symbolic
            # phase together (it is collective), then the factorisation succeeds
            import scipy.sparse as sp
            full = EstimationModel(4, 8, 2, 3)
            local = EstimationModel(4, 8, 2, 3, local_blocks=[i for i in range(4) if i % world == rank])
            kkt, rhs = local.build_kkt(), local.build_rhs()
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
            before = solver.symbolic_calls
            if rank == 1:
                i = local.local_blocks[0]
                K = kkt.get_block(i, i).tocoo()
                n = K.shape[0]
                K2 = sp.coo_matrix((np.concatenate([K.data, np.zeros(n)]),
                                    (np.concatenate([K.row, np.arange(n)]), np.concatenate([K.col, np.arange(n)]))), shape=K.shape)
                kkt.set_block(i, i, K2)
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
            assert solver.symbolic_calls == before + 1          # on both ranks
            x = solver.do_back_solve(rhs)
            st, x_ref, _ = solve_partitioned(full.build_kkt(), full.build_rhs(), world)
            for i in local.local_blocks:
                assert np.allclose(x.get_block(i), x_ref.get_block(i), rtol=1e-9, atol=1e-9)
            assert solver.get_inertia() == full.expected_inertia()
        elif case == "device_regularization":
            # the inertia-correction retry under the multi-rank control flow: base matrix + diagonal shifts, the values
            # of the evaluation are re-used on every rank (no second gather, no second symbolic phase), and a whole
            # interior-point solve takes the path of the single-process reference algorithm
            from oracle.ipm import StochasticInterface, device_regularized, ip_solve, random_stochastic_qp
            from oracle.schur_oracle import OraclePlugin, SchurOracle
            args = (1, 4, 40, 14, 6, 4, 0.3)
            scen, fs = random_stochastic_qp(*args)
            itf = device_regularized(StochasticInterface)(scen, fs)
            itf.set_barrier_parameter(0.1)
            for s_ in itf.sc:
                s_.nlp.x = np.full(s_.nlp.n, 0.3)
            kkt, rhs = itf.evaluate_primal_dual_kkt_matrix(), itf.evaluate_primal_dual_kkt_rhs()
            solver = B200SchurComplementLinearSolver(backend=FakeBackend(), comm=comm,
                                                     regularization_classes=itf.regularization_classes())
            assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
            first = solver.do_numeric_factorization(kkt, raise_on_error=False).status
            assert first in (LinearSolverStatus.successful, LinearSolverStatus.singular)
            reg = itf.regularize_hessian(itf.regularize_equality_gradient(kkt.copy(), -1e-2, False), 1e-2, False)
            assert solver.do_numeric_factorization(reg).status == LinearSolverStatus.successful
            x = solver.do_back_solve(rhs)
            o = SchurOracle(compute_inertia=True, inertia_method="ldl")
            full = reg.materialize()
            o.symbolic(full)
            assert o.numeric(full) == 0
            x_ref = o.solve(rhs)
            for i in list(solver.local_block_indices) + [4]:
                assert np.allclose(np.asarray(x.get_block(i)).flatten(), np.asarray(x_ref.get_block(i)).flatten(),
                                   rtol=1e-7, atol=1e-9)
            assert solver.get_inertia() == o.inertia()
            assert solver.symbolic_calls == 1 and solver.backend.value_uploads() == 1

            class Gathered:          # the restated loop keeps its iterates replicated: every rank needs every block
                def __init__(self, s): self.s = s
                def __getattr__(self, name): return getattr(self.s, name)
                def do_back_solve(self, rhs):
                    sol = self.s.do_back_solve(rhs)
                    for part in comm.allgather_object({i: sol.get_block(i) for i in self.s.local_block_indices}):
                        for i, blk in part.items():
                            sol.set_block(i, blk)
                    return sol

            scen, fs = random_stochastic_qp(*args)
            itf = device_regularized(StochasticInterface)(scen, fs)
            solver = B200SchurComplementLinearSolver(backend=FakeBackend(), comm=comm,
                                                     regularization_classes=itf.regularization_classes())
            out = ip_solve(itf, Gathered(solver))
            scen, fs = random_stochastic_qp(*args)
            ref = ip_solve(StochasticInterface(scen, fs), OraclePlugin(inertia_method="ldl"))
            assert out["status"] == ref["status"] == "optimal" and out["iterations"] == ref["iterations"]
            assert [r[1:] for r in out["reg"]] == [r[1:] for r in ref["reg"]]          # same retries, same inertia
            assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
            assert len(out["reg"]) > out["iterations"]                                  # there were retries ...
            assert solver.symbolic_calls == 1 and solver.backend.value_uploads() == out["iterations"]   # ... for free
        elif case == "fuzz":
            # tools/fuzz_host.py over several ranks: every rank draws the same random systems (input forms, nested
            # blocks, absent borders, pattern changes on re-factorisation) and checks its own blocks; with fewer blocks
            # than ranks a rank owns nothing and still takes part in every collective
            import importlib.util
            spec = importlib.util.spec_from_file_location("fuzz_host", os.path.join(ROOT, "tools", "fuzz_host.py"))
            fuzz = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(fuzz)
            rng = np.random.default_rng(77)
            for c in range(30):
                fuzz.one(rng, c, comm=comm)
        open(os.path.join(out_dir, f"ok_{case}_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["generator", "singular", "rank_local_failure", "pattern_change_one_rank", "cliques",
                                  "device_regularization"])
def test_world_size_2(tmp_path, case):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok_{case}_0", f"ok_{case}_1"]


def test_world_size_3_uneven(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(3, port, "generator", str(tmp_path)), nprocs=3, join=True)
    assert len(os.listdir(tmp_path)) == 3


@pytest.mark.parametrize("world", [2, 3])
def test_fuzzed_inputs_over_ranks(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, "fuzz", str(tmp_path)), nprocs=world, join=True)
    assert len(os.listdir(tmp_path)) == world
