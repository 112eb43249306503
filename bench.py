"""Benchmark of the hot path: KKT numeric factorisation + inertia + back-solve per IPM iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the reference's own synthetic block-bordered KKT generator,
``Model(64 * N, n_q=150, n_y_multiplier=6, n_theta=50)`` -> 64 blocks of 2000 rows per GPU and 50
coupling variables (weak scaling: the block count grows with N, blocks are dealt round-robin as in
reference ``mpi_sc_ip_interface.py:14-19``; the Schur complement and the coupling right-hand side are
all-reduced over NCCL).  A step is one ``do_numeric_factorization`` + ``get_inertia`` +
``do_back_solve`` (symbolic phase excluded and reported separately, as SURVEY.md 8(d) defines).

One JSON line on stdout (rank 0).  ``value``: device-resident inputs/outputs; ``e2e``: the plugin API
with host ``BlockMatrix`` / ``BlockVector`` objects, host<->device copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_Q, N_Y_MULT, N_THETA, BLOCKS_PER_GPU = 150, 6, 50, 64
METRIC = "kkt_factor_solve_ms_per_ipm_iter"


def workload_name(n_gpus):
    return (f"parapint synthetic block-bordered KKT Model({BLOCKS_PER_GPU * n_gpus},{N_Q},{N_Y_MULT},{N_THETA}): "
            f"{BLOCKS_PER_GPU * n_gpus} blocks x 2000 rows, 50 coupling vars; numeric factorization + inertia + back_solve")


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.power, self.reason_bits = [], [], 0
        self.sm_max = None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.02)  # NVML queries contend with kernel launches for the driver lock: keep them sparse

    def finish(self):
        self._halt.set()
        self.join(timeout=10)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
        reasons = [n for n, attr in names if self.reason_bits & int(getattr(nv, attr, 0))]
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(self.power)}


def update_flops(n, m, nb):
    """Algorithmic flops of the DMMA trailing updates of one front: sum over panels of r^2 * kw, with
    r the trailing order after the panel (lower triangle: r^2/2 entries, 2 flops per entry per column)."""
    nf, k, tot, launches = n + m, 0, 0.0, 0
    while k < n:
        kw = min(nb - 1, n - k) if n - k > nb else n - k
        k += kw
        r = nf - k
        if r > 0:
            tot += float(r) * r * kw
            launches += 1
    return tot, launches


def fp64_peak():
    path = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return d["cublas_dgemm_8192_tflops"], "cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json holds no FP64 figure"
    except Exception:  # noqa: BLE001
        return 37.0, "nominal B200 FP64 (no measured file)"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_class):
    """DRAM bytes per launch of the kernels of a profile class at the HEADLINE workload, from the committed ncu
    captures (profiles/ncu_traffic_r02.json, generated by tools/ncu_summary.py --traffic from profiles/ncu_full_r02_*.json:
    `ncu --set full` of tools/profile_case.py 64 2 = config 2)."""
    parts = {"subtree": ("subtree_factor", "subtree_leaf_kernel"), "forward": ("subtree_forward",),
             "backward": ("subtree_backward",), "panel": ("front_small",)}.get(kernel_class)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as fh:
            table = json.load(fh)
        return float(sum(table[p]["bytes"] for p in parts)) if parts else None
    except Exception:  # noqa: BLE001
        return None


def _residual(kkt, rhs, x, st, comm, dev):
    """||K x - b|| / ||b|| of the whole block-bordered system, each rank contributing the rows of its blocks (the
    matrices are used as given: both triangles of K_i are stored, the lower border A_i stands for both borders)."""
    import torch
    N = st.n_blocks
    x_c = np.asarray(x.get_block(N)).ravel()
    acc = np.zeros(2)
    coupling = np.zeros(st.m_c)
    for i in st.local_blocks:
        K = kkt.get_block(i, i).tocsr()
        A = kkt.get_block(N, i)
        xi, ri = np.asarray(x.get_block(i)).ravel(), np.asarray(rhs.get_block(i)).ravel()
        r = K @ xi - ri
        if A is not None:
            A = A.tocsr()
            r += A.T @ x_c
            coupling += A @ xi
        acc += (np.sum(r * r), np.sum(ri * ri))
    red = torch.tensor(np.concatenate([acc, coupling]), device=dev)
    comm.allreduce_sum_(red)
    red = red.cpu().numpy()
    Q = kkt.get_block(N, N)
    bc = np.asarray(rhs.get_block(N)).ravel()
    rc = red[2:] - bc
    if Q is not None and st.m_c:
        Qc = Q.tocsr()
        rc = rc + sp_tril_sym(Qc) @ x_c
    return float(np.sqrt(red[0] + np.sum(rc * rc)) / np.sqrt(red[1] + np.sum(bc * bc)))


def sp_tril_sym(Q):
    """Q as the solver reads it: the lower triangle, mirrored."""
    import scipy.sparse as sp
    L = sp.tril(Q)
    return (L + sp.tril(Q, -1).T).tocsr()


def _worst_block_vs_superlu(kkt, rhs, x, st):
    """Per-block residual of the solution on this rank, and -- for the worst block -- the residual the reference's
    SciPy leaf (SuperLU, scipy_interface.py:25-62) reaches on the same block system K_i x_i = b_i - A_i^T x_c."""
    import scipy.sparse.linalg as spla
    N = st.n_blocks
    x_c = np.asarray(x.get_block(N)).ravel()
    worst, worst_i, worst_sys = -1.0, None, None
    for i in st.local_blocks:
        K = kkt.get_block(i, i).tocsc()
        A = kkt.get_block(N, i)
        bi = np.asarray(rhs.get_block(i)).ravel()
        ri = bi - (A.tocsr().T @ x_c if A is not None else 0.0)
        res = float(np.linalg.norm(K @ np.asarray(x.get_block(i)).ravel() - ri) / np.linalg.norm(bi))
        if res > worst:
            worst, worst_i, worst_sys = res, i, (K, ri, bi)
    if worst_i is None:
        return None
    K, ri, bi = worst_sys
    t0 = time.perf_counter()
    x_ref = spla.splu(K).solve(ri)
    leaf = float(np.linalg.norm(K @ x_ref - ri) / np.linalg.norm(bi))
    return {"block": int(worst_i), "residual_b200": worst, "residual_superlu_leaf": leaf, "superlu_seconds": time.perf_counter() - t0}


def run_leg(name, kkt, rhs, comm, dev, flush, reps=3, warmup=2, options=None, expected_inertia=None, model=None,
            leaf_check=False):
    """One more workload of BASELINE.json through the same measurement as the headline: correctness gate (inertia,
    residual), `e2e` through the plugin API with host buffers, `value` with device-resident inputs, per-kernel-class
    device times and the roofline of the dominant class."""
    import torch
    from parapint_b200 import B200SchurComplementLinearSolver, LinearSolverStatus
    opts = {"profile": 1}
    opts.update(options or {})
    solver = B200SchurComplementLinearSolver(comm=comm, options=opts)
    be = solver.backend

    def sync_all():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    torch.cuda.synchronize()
    sym_s = time.perf_counter() - t0
    st = solver._st
    assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    inertia = solver.get_inertia()
    x = solver.do_back_solve(rhs)
    rel_res = _residual(kkt, rhs, x, st, comm, dev)
    check = {"inertia": list(inertia), "rel_residual": rel_res, "refine_steps": solver.refine_steps}
    if expected_inertia is not None:
        check["inertia_expected"] = list(expected_inertia)
        assert tuple(inertia) == tuple(expected_inertia), (name, inertia, expected_inertia)
    else:
        assert sum(inertia) == st.m_c + comm_sum_int(comm, dev, int(st.local_dim)) and inertia[2] == 0, (name, inertia)
    if model is not None:
        check["max_err"] = float(model.check_result(x))
    if leaf_check and comm.rank == 0:
        # ill-conditioned inputs: the bar is the reference leaf's residual on the same block
        wb = _worst_block_vs_superlu(kkt, rhs, x, st)
        check["worst_block_on_rank0"] = wb
        assert wb is None or wb["residual_b200"] <= max(1e-10, 10.0 * wb["residual_superlu_leaf"]), wb

    def e2e_step():
        flush.fill_(1.0)
        solver.do_numeric_factorization(kkt)
        solver.get_inertia()
        return solver.do_back_solve(rhs)

    be.set_option("profile", 0)     # per-kernel events only in the `value` region below
    for _ in range(warmup):
        e2e_step()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        e2e_step()
    ev1.record()
    sync_all()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - tw) * 1e3) / reps

    values_dev, rhs_dev, rhsc_dev = be.values_pin.to(dev), be.rhs_pin.to(dev), be.rhsc_pin.to(dev)
    x_dev, xc_dev = torch.empty_like(rhs_dev), torch.empty_like(rhsc_dev)

    def dev_step():
        flush.fill_(1.0)
        _, s_local = be.numeric_local_device(values_dev)
        s_local = comm.allreduce_sum_(s_local)
        code = be.numeric_coupling(s_local)
        be.inertia_coupling()
        be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_)
        return code

    be.set_option("profile", 1)
    for _ in range(warmup):
        dev_step()
    be.profile(reset=True)
    l0 = be.kernel_launches()
    sync_all()
    ev0.record()
    for _ in range(reps):
        code = dev_step()
    ev1.record()
    sync_all()
    assert code == 0
    dev_ms = ev0.elapsed_time(ev1) / reps
    launches = (be.kernel_launches() - l0) // reps
    prof = be.profile(reset=True)
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    comm.allreduce_max_(t)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    per = {k: v["ms"] / reps for k, v in prof.items()}
    stats = be.plan_stats(0) if st.n_local else {}
    cstats = be.coupling_stats()
    nnz_in = int((st.dest_front >= 0).sum())
    nnz_l = sum(be.plan_stats(b)["nnz_l_subtree"] for b in range(st.n_local))
    out = {"workload": name, "value": dev_ms, "unit": "ms", "e2e": {"value": e2e_ms, "unit": "ms",
           "h2d_bytes_per_step": int((st.nvals + st.local_dim + st.m_c) * 8), "d2h_bytes_per_step": int((st.local_dim + st.m_c) * 8 + 56)},
           "steps": reps, "warmup": warmup, "symbolic_s": sym_s, "gpu_launches": int(launches), "kernels_ms_per_step": per,
           "blocks_this_rank": st.n_local, "coupling": cstats, "factor_bytes": be.factor_bytes(), "check": check,
           "symbolic": {k: stats.get(k) for k in ("supernodes", "root_cols", "delay_slots", "nnz_l_subtree", "n_plans",
                                                  "fell_back_dense", "delayed_to_root")}}
    out["roofline"] = leg_roofline(per, prof, reps, st, stats, nnz_in, nnz_l)
    del solver, be
    torch.cuda.empty_cache()
    return out


def comm_sum_int(comm, dev, v):
    import torch
    t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    comm.allreduce_sum_(t)
    return int(round(float(t[0])))


def leg_roofline(per, prof, reps, st, stats, nnz_in, nnz_l):
    """Roofline of the dominant kernel class (device time per step, CUDA events around every launch)."""
    dominant = max(per, key=per.get)
    n_launch = max(prof[dominant]["launches"] // reps, 1)
    avg_s = per[dominant] / n_launch * 1e-3
    hbm, hbm_src = hbm_peak()
    peak64, peak64_src = fp64_peak()
    m_blk = int(st.border_ptr[1] - st.border_ptr[0]) if st.n_local else 0
    root_n = (stats.get("root_cols", 0) or 0) + (stats.get("delayed_to_root", 0) or 0)
    if dominant == "update":
        flops, launches = update_flops(root_n, m_blk, 64)
        ach = flops * st.n_local / (per[dominant] * 1e-3) / 1e12
        return {"kernel": "front_update_kernel", "bound": "tensor", "achieved": ach, "peak": peak64, "unit": "TFLOP/s",
                "frac": ach / peak64, "traffic": None, "peak_source": peak64_src,
                "algorithmic_flops_per_step": flops * st.n_local, "launches_per_step": n_launch, "share_of_step": per[dominant] / max(sum(per.values()), 1e-30)}
    if dominant == "subtree":
        alg, kname = 12.0 * nnz_in + 8.0 * nnz_l, "subtree_factor_kernel (+ subtree_leaf_kernel)"
    elif dominant in ("forward", "backward"):
        nf = root_n + m_blk
        alg, kname = 8.0 * nnz_l + 24.0 * st.local_dim + 8.0 * nf * nf / 2 * st.n_local, f"subtree_{dominant}_kernel + front_{dominant}_kernel"
    elif dominant == "panel":
        # every pivot column re-reads the panel computed so far: sum_k (nf - k) * (k mod 64) * 8 B  ~  nf * n * 32 * 8 B
        nf = root_n + m_blk
        alg, kname = 8.0 * 32.0 * root_n * (nf - root_n / 2.0) * st.n_local, "front_panel_(cluster_)kernel"
    else:
        nf = root_n + m_blk
        alg, kname = 8.0 * nf * nf / 2 * st.n_local, dominant
    ach = alg / n_launch / avg_s / 1e9
    return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
            "peak_source": hbm_src, "algorithmic_bytes_per_launch": alg / n_launch, "launches_per_step": n_launch,
            "share_of_step": per[dominant] / max(sum(per.values()), 1e-30)}


def config5_leg(comm, dev, flush, world, rank):
    """BASELINE config 5 (wide-coupling stress: 128 blocks x 20 000 rows, 2 000 coupling variables), STRONG scaling:
    the 128 blocks are dealt over the ranks.  Dense root fronts of ~4 082 rows: panel kernel + FP64 DMMA update."""
    from oracle.kkt_generator import EstimationModel
    nb, nq, ym, nth = 128, 2000, 4, 2000
    local = [i for i in range(nb) if i % world == rank]
    m = EstimationModel(nb, nq, ym, nth, local_blocks=local)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    full = EstimationModel.__new__(EstimationModel)
    full.n_blocks, full.n_q, full.n_y, full.n_theta = nb, nq, nq * ym, nth
    out = run_leg(f"config 5: Model({nb},{nq},{ym},{nth}) = {nb} blocks x {m.block_dim} rows, {nth} coupling vars, over {world} GPU(s)",
                  kkt, rhs, comm, dev, flush, reps=2, warmup=1, expected_inertia=full.expected_inertia(), model=m)
    out["scaling"] = "strong"
    return out


def config4_leg(comm, dev, flush, world, rank, per_gpu=128):
    """BASELINE config 4's shape (two-stage stochastic NLP, scenarios of 20 200 rows, 200 first-stage variables):
    128 scenarios per GPU -- the share one GPU holds when the 1 024 scenarios run on 8 (weak scaling towards it;
    generating 1 024 scenarios on one host takes longer than the driver's time limit for this run)."""
    from oracle.kkt_families import stochastic_ipm_system
    from tests.helpers import block_vector
    nb = per_gpu * world
    local = [i for i in range(nb) if i % world == rank]
    kkt, sizes = stochastic_ipm_system(7, nb, 10000, 8000, 1000, 200, same_pattern=True, local_blocks=local)
    rhs = local_block_vector(np.random.default_rng(11), sizes, local)
    out = run_leg(f"config 4 share: {nb} scenarios x {sizes[0]} rows (n_x 10000, n_eq 8000, n_in 1000), 200 first-stage vars, "
                  f"{per_gpu} per GPU", kkt, rhs, comm, dev, flush, reps=2, warmup=1, leaf_check=True)
    out["scaling"] = "weak"
    out["check"]["note"] = ("family-P blocks at this size are ill conditioned (barrier terms over 8 decades): the residual bar is "
                            "'no worse than the SuperLU leaf on the same block' -- worst_block_on_rank0 measures it here, "
                            "tests/test_gpu_parity.py::test_config4_shape_* pins it")
    return out


def config3_leg(comm, dev, flush, world, rank):
    """BASELINE config 3's shape (time decomposition, 256 time blocks x ~10 050 rows, 50 interface states => 25 500
    coupling rows, block-tridiagonal S held sparse and factorised level by level), STRONG scaling."""
    from oracle.kkt_families import dynamic_ipm_system
    nb, n_s = 256, 50
    local = [i for i in range(nb) if i % world == rank]
    kkt, sizes = dynamic_ipm_system(9, nb, 5000, 4800, 100, n_s, same_pattern=True, local_blocks=local)
    rhs = local_block_vector(np.random.default_rng(13), sizes, local)
    out = run_leg(f"config 3: {nb} time blocks x {sizes[1]} rows (n_x 5000, n_eq 4800, n_in 100), {n_s} interface states, "
                  f"m_c = {sizes[-1]}, over {world} GPU(s)", kkt, rhs, comm, dev, flush, reps=2, warmup=1)
    out["scaling"] = "strong"
    return out


def local_block_vector(rng, sizes, local):
    from parapint_b200 import BlockVector
    v = BlockVector(len(sizes))
    for i, n in enumerate(sizes[:-1]):
        r = rng.standard_normal(n)      # drawn on every rank so that the streams stay aligned
        if i in local:
            v.set_block(i, r)
    v.set_block(len(sizes) - 1, rng.standard_normal(sizes[-1]))
    return v


class _GatheredSolve:
    """The restated IPM loop keeps its iterates replicated on every rank (flat vectors); the solver is distributed.
    After the back-solve every rank needs every block of the step: what MPIBlockVector operations do in the reference."""

    def __init__(self, solver, comm):
        self.s, self.comm = solver, comm
        self.seconds = 0.0
        self.calls = 0

    def _timed(self, fn, *a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        return out

    def do_symbolic_factorization(self, **k): return self._timed(self.s.do_symbolic_factorization, **k)
    def do_numeric_factorization(self, **k): return self._timed(self.s.do_numeric_factorization, **k)
    def get_inertia(self): return self._timed(self.s.get_inertia)
    def increase_memory_allocation(self, f): return self.s.increase_memory_allocation(f)

    def do_back_solve(self, rhs):
        sol = self._timed(self.s.do_back_solve, rhs)
        if self.comm.size > 1:
            mine = {i: sol.get_block(i) for i in self.s.local_block_indices}
            for part in self.comm.allgather_object(mine):
                for i, blk in part.items():
                    sol.set_block(i, blk)
        return sol


def ip_solve_leg(comm, dev, world, rank):
    """`ip_solve` wall time (second half of BASELINE.json's metric): the restated interior-point loop
    (oracle/ipm.py, following algorithms/interior_point.py:405-631) on a separable two-stage QP with negative curvature
    (the inertia-correction loop refactorises), driven (a) by this solver behind the plain interface, (b) by this
    solver with device-side regularisation, (c) on rank 0 by the reference algorithm on the host (SciPy SuperLU
    leaves, LAPACK inertia).  Iteration counts and objective must agree."""
    from oracle.ipm import StochasticInterface, device_regularized, ip_solve, random_stochastic_qp
    from oracle.schur_oracle import OraclePlugin
    from parapint_b200 import B200SchurComplementLinearSolver
    args = (3, 32, 200, 100, 12, 6, 0.15)
    out = {"workload": f"ip_solve on random_stochastic_qp{args}: 32 scenarios x 330 rows, 6 first-stage vars, over {world} GPU(s)"}

    def run(make_solver, device_reg):
        scen, fs = random_stochastic_qp(*args)
        cls = device_regularized(StochasticInterface) if device_reg else StochasticInterface
        itf = cls(scen, fs)
        solver = make_solver(itf)
        wrapped = _GatheredSolve(solver, comm)
        comm.barrier()
        t0 = time.perf_counter()
        res = ip_solve(itf, wrapped)
        wall = time.perf_counter() - t0
        return res, wall, wrapped, solver

    res_a, wall_a, wa, sa = run(lambda itf: B200SchurComplementLinearSolver(comm=comm), False)
    res_b, wall_b, wb, sb = run(lambda itf: B200SchurComplementLinearSolver(comm=comm, regularization_classes=itf.regularization_classes()), True)
    out["b200"] = {"wall_s": wall_a, "linear_solver_s": wa.seconds, "iterations": res_a["iterations"], "factorizations": len(res_a["reg"]),
                   "objective": res_a["objective"], "value_uploads": sa.backend.value_uploads(), "symbolic_calls": sa.symbolic_calls}
    out["b200_device_regularization"] = {"wall_s": wall_b, "linear_solver_s": wb.seconds, "iterations": res_b["iterations"],
                                         "factorizations": len(res_b["reg"]), "objective": res_b["objective"],
                                         "value_uploads": sb.backend.value_uploads(), "symbolic_calls": sb.symbolic_calls}
    assert res_a["status"] == res_b["status"] == "optimal" and res_a["iterations"] == res_b["iterations"]
    if rank == 0:
        scen, fs = random_stochastic_qp(*args)
        itf = StochasticInterface(scen, fs)
        cpu = _GatheredSolve(OraclePlugin(inertia_method="ldl"), type("C", (), {"size": 1})())
        t0 = time.perf_counter()
        res_c = ip_solve(itf, cpu)
        wall_c = time.perf_counter() - t0
        out["cpu_reference_algorithm"] = {"wall_s": wall_c, "linear_solver_s": cpu.seconds, "iterations": res_c["iterations"],
                                          "factorizations": len(res_c["reg"]), "objective": res_c["objective"], "cores": 1,
                                          "kind": "port (serial SchurComplementLinearSolver + SuperLU leaves, dsytrf inertia)"}
        assert res_c["iterations"] == res_a["iterations"], (res_c["iterations"], res_a["iterations"])
        assert abs(res_c["objective"] - res_a["objective"]) <= 1e-8 * max(1.0, abs(res_c["objective"]))
        out["iterations_equal"] = True
        out["note"] = ("the interior-point loop itself (KKT evaluation, convergence checks, step rules) is host Python in both arms "
                       "(oracle/ipm.py); linear_solver_s is the time inside the plugin calls")
    comm.barrier()
    return out


def ipm_vectors_leg(comm, dev, world, rank):
    """Interior-point vector kernels (csrc/ipmvec.cuh): one pass each over device-resident iterates
    of config 4's size -- 1 024 scenarios x 10 000 primals, this rank's share -- against the HBM roofline.  Vectors are
    larger than L2 together (6 x 8 B x n); CUDA events, max over ranks."""
    import torch
    from parapint_b200.ipm_vectors import IpmKernels
    k = IpmKernels(dev.index)
    n = 1024 * 10000 // world
    g = torch.Generator(device=dev)
    g.manual_seed(1 + rank)
    lb = -torch.rand(n, dtype=torch.float64, device=dev, generator=g) - 0.1
    ub = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    dx = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    zl = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    zu = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    out = torch.ones(8, dtype=torch.float64, device=dev)
    alpha = torch.tensor([1e-3, 1e-3, 1.0], dtype=torch.float64, device=dev)
    flush = torch.empty(20 * 1024 * 1024, dtype=torch.float64, device=dev)
    hbm, hbm_src = hbm_peak()
    res = {}

    def timed(name, fn, nbytes, reps=10):
        for _ in range(3):
            fn()
        tot = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(reps):
            flush.fill_(1.0)          # 160 MB written between launches: the vectors never come from L2
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            tot += ev0.elapsed_time(ev1)
        t = torch.tensor([tot / reps], device=dev, dtype=torch.float64)
        comm.allreduce_max_(t)
        ms = float(t[0])
        res[name] = {"ms": ms, "achieved": nbytes / ms * 1e-6, "unit": "GB/s", "algorithmic_bytes": nbytes,
                     "frac": nbytes / ms * 1e-6 / hbm}

    timed("fraction_to_boundary", lambda: k.fraction_to_boundary(out[:2], 0.99, 1e-2, x, dx, lb, ub, zl, zu), 48 * n)
    timed("complementarity", lambda: k.complementarity(out[:6], 1e-2, x, lb, ub, zl, zu), 40 * n)
    timed("max_abs", lambda: k.max_abs(out[:2], x, dx), 16 * n)
    timed("step", lambda: k.step(alpha, 1e-2, x, dx, lb, ub, zl, zu), 72 * n)
    timed("axpy", lambda: k.axpy(alpha, 1, x, dx), 24 * n)
    return {"workload": f"interior-point vector kernels on {n} primals per GPU (config 4: 1024 x 10 000 over {world} GPU(s))",
            "bound": "hbm", "peak": hbm, "peak_source": hbm_src, "kernels": res,
            "note": "bytes = 8 B per vector element read or written once; parity (bit-exact minima / maxima) in tests/test_gpu_ipm_vectors.py"}



def build_model(n_gpus, rank):
    from oracle.kkt_generator import EstimationModel  # generator = input synthesis only
    n_blocks = BLOCKS_PER_GPU * n_gpus
    local = [i for i in range(n_blocks) if i % n_gpus == rank]
    return EstimationModel(n_blocks, N_Q, N_Y_MULT, N_THETA, local_blocks=local)


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: pyomo/mpi4py are absent so the
    reference package itself cannot be imported; see oracle/__init__.py) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.parallel_baseline import PartitionedCpuSolver
    model = build_model(1, 0) if args.gpus == 1 else None
    if model is None:
        from oracle.kkt_generator import EstimationModel
        model = EstimationModel(BLOCKS_PER_GPU * args.gpus, N_Q, N_Y_MULT, N_THETA)
    kkt, rhs = model.build_kkt(), model.build_rhs()
    cores = os.cpu_count() or 1
    with PartitionedCpuSolver(kkt, rhs, procs=cores) as solver:
        for _ in range(args.warmup):
            solver.factor_and_solve()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            status, blocks, x_c = solver.factor_and_solve()
        dt = time.perf_counter() - t0
        used = solver.procs
    assert status == 0
    ms = dt / args.steps * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (reference generator, built-in seeds)",
        "config": {"workload": workload_name(args.gpus)},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": used, "kind": "port",
                         "sample": f"full workload, {args.steps} steps; process-per-rank emulation of "
                                   f"MPISchurComplementLinearSolver with SciPy SuperLU leaves on {used} processes "
                                   "(mpirun/mpi4py/MUMPS absent from the image)"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE workloads (configs 3, 4, 5, ip_solve)")
    ap.add_argument("--legs", default="config5,config4,config3,ip_solve,ipm_vectors",
                    help="comma-separated subset of the other workloads to run after the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and os.environ.get("PARAPINT_B200_PIN", "0") == "1":   # measured on this pool (one NUMA node, 32 vCPUs): no gain, off by default
        # every rank keeps to its own share of the cores this job may use (before any worker thread exists: the host
        # copy pool and the NCCL helpers inherit the mask), so that eight interpreters do not migrate over each other
        try:
            cores = sorted(os.sched_getaffinity(0))
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            share = cores[local_rank * len(cores) // local_world:(local_rank + 1) * len(cores) // local_world]
            if share:
                os.sched_setaffinity(0, share)
                os.environ.setdefault("PARAPINT_B200_HOST_THREADS", str(max(1, min(8, len(share) - 1))))
        except (AttributeError, OSError):
            pass
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None) if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN") else None
        # (NCCL prints a version banner on stdout at VERSION/WARN level; stdout must stay the one JSON line)
        dist.init_process_group("nccl", device_id=dev)

    from parapint_b200 import B200SchurComplementLinearSolver, Communicator, LinearSolverStatus
    from parapint_b200 import structure

    comm = Communicator()
    model = build_model(world, rank)
    kkt, rhs = model.build_kkt(), model.build_rhs()
    solver = B200SchurComplementLinearSolver(comm=comm, options={"profile": 1})
    be = solver.backend

    def sync_all():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    torch.cuda.synchronize()
    symbolic_ms = (time.perf_counter() - t0) * 1e3
    st = solver._st

    # ---- correctness gate (every timed run): inertia, residual --------------------------------
    assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    inertia = solver.get_inertia()
    x = solver.do_back_solve(rhs)
    assert inertia == model.expected_inertia(), (inertia, model.expected_inertia())
    x_c = np.asarray(x.get_block(model.n_blocks))
    border = model.border().tocsr()
    res2 = np.zeros(2)
    coupling_res = np.zeros(N_THETA)
    for i in st.local_blocks:
        K = kkt.get_block(i, i).tocsr()
        xi, ri = np.asarray(x.get_block(i)), np.asarray(rhs.get_block(i))
        res2[0] += np.sum((K @ xi + border.T @ x_c - ri) ** 2)
        res2[1] += np.sum(ri ** 2)
        coupling_res += border @ xi
    red = torch.tensor(np.concatenate([res2, coupling_res]), device=dev)
    comm.allreduce_sum_(red)
    red = red.cpu().numpy()
    rel_res = float(np.sqrt(red[0] + np.sum(red[2:] ** 2)) / np.sqrt(red[1]))
    assert rel_res <= 1e-10, rel_res
    max_err = model.check_result(x)

    # ---- end-to-end through the plugin API (host buffers) ----------------------------------------
    # (no instrumentation here: the per-kernel CUDA events of `pp_profile` -- two cudaEventRecord per launch -- are
    #  what a user of the plugin never switches on; they are recorded in the `value` region, where the roofline needs them)
    be.set_option("profile", 0)
    flush = torch.empty(20 * 1024 * 1024, dtype=torch.float64, device=dev)  # 160 MB > 126 MB of L2

    def e2e_step():
        flush.fill_(1.0)  # evict L2 between steps (inside the timed region: ~0.1 ms, counted against us)
        res = solver.do_numeric_factorization(kkt)
        ine = solver.get_inertia()
        sol = solver.do_back_solve(rhs)
        return res, ine, sol

    for _ in range(max(args.warmup, 10)):   # (untimed; the host side -- page faults, thread pools -- warms up slowly)
        e2e_step()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    sync_all()
    e2e_wall_ms = (time.perf_counter() - tw) * 1e3 / args.steps
    e2e_ms = max(ev0.elapsed_time(ev1) / args.steps, e2e_wall_ms)
    h2d = (st.nvals + st.local_dim + st.m_c) * 8
    d2h = (st.local_dim + st.m_c) * 8 + 4 * 2 + 6 * 8

    # ---- the same step on FRESH matrix objects every time (what an interface that rebuilds its KKT hands over,
    #      interfaces/interface.py:432-491: the identity fast path of the gather never applies; the objects are built
    #      outside the timed region -- building them is the interface's work, validating and gathering them is ours)
    fresh = [(model.build_kkt(), model.build_rhs()) for _ in range(6)]
    n_fresh = max(12, min(args.steps, 60))

    def fresh_step(k):
        kk, rr = fresh[k % len(fresh)]
        flush.fill_(1.0)
        solver.do_numeric_factorization(kk)
        solver.get_inertia()
        return solver.do_back_solve(rr)

    for k in range(args.warmup):
        fresh_step(k)
    sync_all()
    tw = time.perf_counter()
    ev0.record()
    for k in range(n_fresh):
        xf = fresh_step(k)
    ev1.record()
    sync_all()
    fresh_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - tw) * 1e3) / n_fresh
    assert np.allclose(xf.get_block(model.n_blocks), x.get_block(model.n_blocks), rtol=1e-9, atol=1e-12)
    del fresh

    # ---- device-resident hot path (`value`) --------------------------------------------------------
    values_dev = be.values_pin.to(dev)
    rhs_dev = be.rhs_pin.to(dev)
    rhsc_dev = be.rhsc_pin.to(dev)
    x_dev = torch.empty_like(rhs_dev)
    xc_dev = torch.empty_like(rhsc_dev)

    def dev_step():
        flush.fill_(1.0)
        code, s_local = be.numeric_local_device(values_dev)
        s_local = comm.allreduce_sum_(s_local)
        code2 = be.numeric_coupling(s_local)
        loc = be.inertia_local()
        cpl = be.inertia_coupling()
        be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_)
        return code, code2, loc, cpl

    sampler = ClockSampler(local_rank) if rank == 0 else None  # NVML init takes milliseconds: before the barrier
    be.set_option("profile", 1)
    for _ in range(args.warmup):
        dev_step()
    be.profile(reset=True)
    launches0 = be.kernel_launches()
    if sampler:
        sampler.start()
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        out = dev_step()
    ev1.record()
    sync_all()
    clocks = sampler.finish() if sampler else None
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (be.kernel_launches() - launches0) // args.steps
    prof = be.profile(reset=True)
    assert out[0] == 0 and out[1] == 0
    x_chk = x_dev.cpu().numpy()
    assert np.allclose(x_chk, be.x_pin.numpy()[: x_chk.size], rtol=1e-12, atol=1e-12)

    t = torch.tensor([dev_ms, e2e_ms, fresh_ms], device=dev, dtype=torch.float64)
    comm.allreduce_max_(t)
    dev_ms, e2e_ms, fresh_ms = float(t[0]), float(t[1]), float(t[2])

    stats = be.plan_stats(0)
    nnz_l = sum(be.plan_stats(b)["nnz_l_subtree"] for b in range(st.n_local))

    # ---- the other workloads BASELINE.json names (every rank takes part; rank 0 reports) ----------------
    others = {}
    if not args.no_secondary:
        del solver, be
        torch.cuda.empty_cache()
        legs = [l.strip() for l in args.legs.split(",") if l.strip()]
        for leg in legs:
            t_leg = time.perf_counter()
            try:
                if leg == "config5":
                    others["config5_strong"] = config5_leg(comm, dev, flush, world, rank)
                elif leg == "config4":
                    others["config4_share"] = config4_leg(comm, dev, flush, world, rank)
                elif leg == "config3":
                    others["config3_strong"] = config3_leg(comm, dev, flush, world, rank)
                elif leg == "ip_solve":
                    others["ip_solve"] = ip_solve_leg(comm, dev, world, rank)
                elif leg == "ipm_vectors":
                    others["ipm_vectors"] = ipm_vectors_leg(comm, dev, world, rank)
                else:
                    raise SystemExit(f"unknown leg {leg}")
                key = list(others)[-1]
                others[key]["leg_wall_s"] = time.perf_counter() - t_leg
            except torch.cuda.OutOfMemoryError as e:   # a leg that does not fit is reported, not hidden
                others[leg] = {"skipped": f"out of device memory: {e}"}
                torch.cuda.empty_cache()
            except (RuntimeError, AssertionError) as e:  # a leg that fails is reported as failed; the headline stands
                others[leg] = {"failed": f"{type(e).__name__}: {str(e)[:400]}"}
                print(f"[bench] leg {leg} failed: {e}", file=sys.stderr, flush=True)
                if world > 1:
                    raise                                  # ranks would desynchronise: better no line than a hang

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------
    per_step = {k: v["ms"] / args.steps for k, v in prof.items()}
    dominant = max(per_step, key=per_step.get)
    nnz_in = int((st.dest_front >= 0).sum())                       # lower-triangle input entries (all local blocks + Q)
    n_launch = max(prof[dominant]["launches"] // args.steps, 1)
    avg_s = per_step[dominant] / n_launch * 1e-3
    hbm, hbm_src = hbm_peak()
    peak64, peak64_src = fp64_peak()
    if dominant == "subtree":
        # SURVEY.md 8(d) "K1 sparse": 12 B per input entry read + 8 B per entry of L written
        alg = 12.0 * nnz_in + 8.0 * nnz_l
        kname = "subtree_factor_kernel (+ subtree_leaf_kernel)"
    elif dominant in ("forward", "backward"):
        # SURVEY.md 8(d) "K7": L read once per sweep (8 B values; indices are per front, not per entry) + 3 vectors
        alg = 8.0 * nnz_l + 24.0 * st.local_dim
        kname = f"subtree_{dominant}_kernel (+ leaf, + front_{dominant}_kernel on the roots)"
    elif dominant == "update":
        alg = None
        kname = "front_update_kernel"
    else:
        root_nf = stats["root_cols"] + stats["delay_slots"] + N_THETA
        alg = 8.0 * root_nf * root_nf / 2 * st.n_local
        kname = {"panel": "front_panel_kernel", "assemble": "assemble_kernel", "schur": "schur_gather_kernel",
                 "swaps": "front_swaps_left_kernel"}.get(dominant, dominant)
    if alg is None:
        root_n = stats["root_cols"] + stats["delay_slots"]
        flops_front, upd_launches = update_flops(root_n, N_THETA, 64)
        achieved = flops_front * st.n_local / max(upd_launches, 1) / avg_s / 1e12
        roofline = {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak64, "unit": "TFLOP/s",
                    "frac": achieved / peak64, "traffic": ncu_traffic(kname), "peak_source": peak64_src}
    else:
        achieved = alg / n_launch / avg_s / 1e9
        roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm, "traffic": ncu_traffic(dominant), "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": alg / n_launch,
                    "note": "multifrontal elimination of ~900 tiny fronts per block: bounded by the dependency "
                            "chain of the assembly tree (latency), not by bandwidth; see DESIGN.md section 4"}

    line = {
        "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference generator create_model.Model, built-in seeds)",
        "config": {"workload": workload_name(world), "blocks_per_gpu": BLOCKS_PER_GPU, "block_rows": model.block_dim,
                   "coupling": N_THETA, "l2": "L2 flushed by a 160 MB device write (> 126 MB of L2) before every step (inside the timed region)",
                   "value_excludes": "the residual check / iterative refinement that do_back_solve runs (e2e includes it)",
                   "instrumentation": "per-kernel CUDA events (pp_profile, for the roofline) are recorded inside the value region only; e2e runs without them",
                   "parallelism": f"blocks round-robin over {world} GPU(s); Schur complement + coupling rhs all-reduced (NCCL)"},
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "wall_ms": e2e_wall_ms, "api": "B200SchurComplementLinearSolver.do_numeric_factorization + get_inertia + do_back_solve",
                "fresh_objects_ms": fresh_ms,
                "fresh_objects_note": "same step on newly built BlockMatrix / BlockVector objects every time (no identity fast path)"},
        "gpu_launches": int(launches),
        "exchange": {"small_payloads": "one-shot peer all-reduce over symmetric memory (csrc/peer.cuh)" if comm._peer else
                     ("none (one rank)" if world == 1 else "NCCL all-reduce"), "fallback_reason": comm.peer_error},
        "roofline": roofline,
        "kernels_ms_per_step": per_step,
        "symbolic": {k: stats[k] for k in ("supernodes", "root_cols", "delay_slots", "nnz_l_subtree", "max_front",
                                           "n_plans", "fell_back_dense", "delayed_to_root")},
        "throughput": {"value": BLOCKS_PER_GPU * world / (dev_ms * 1e-3), "unit": "kkt_blocks/s"},
        "symbolic_ms": symbolic_ms,
        "check": {"inertia": list(inertia), "rel_residual": rel_res, "max_err": float(max_err)},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.parallel_baseline import time_serial
        full_kkt, full_rhs = kkt, rhs
        sec, reps, x_ref = time_serial(full_kkt, full_rhs, min_seconds=8.0)
        rel = float(np.linalg.norm(x.flatten() - x_ref.flatten()) / np.linalg.norm(x_ref.flatten()))
        assert rel <= 1e-8, rel
        line["cpu_baseline"] = {"value": sec * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                                "sample": f"full workload, median of {reps} runs of the serial reference algorithm "
                                          "(oracle port of SchurComplementLinearSolver + ScipyInterface/SuperLU leaves)"}
        line["check"]["rel_diff_vs_cpu_reference"] = rel
    if others:
        line["other_workloads"] = others
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
