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


def ncu_traffic(kernel):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(kernel)
    except Exception:  # noqa: BLE001
        return None


def wide_border_leg(dev):
    """Config-5 slice (BASELINE.json configs[4] is 128 x 20,000 x 2,000; 8 of its blocks here): the regime where the
    work is the dense root fronts (2,082 pivots + 2,000 border rows each) and the FP64 tensor-core update kernel is
    the dominant one -- reported next to the headline so that both rooflines of the path are on the line."""
    import torch
    from oracle.kkt_generator import EstimationModel
    from parapint_b200 import B200SchurComplementLinearSolver
    nb, nq, ym, nth = 8, 2000, 4, 2000
    m = EstimationModel(nb, nq, ym, nth)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s = B200SchurComplementLinearSolver(options={"profile": 1})
    t0 = time.perf_counter()
    s.do_symbolic_factorization(kkt)
    sym_s = time.perf_counter() - t0
    reps, times = 4, []
    for r in range(reps):
        torch.cuda.synchronize(dev)
        if r == 1:
            s.backend.profile()
        t0 = time.perf_counter()
        s.do_numeric_factorization(kkt)
        inertia = s.get_inertia()
        x = s.do_back_solve(rhs)
        torch.cuda.synchronize(dev)
        times.append((time.perf_counter() - t0) * 1e3)
    prof = s.backend.profile()
    stats = s.backend.plan_stats(0)
    per = {k: v["ms"] / (reps - 1) for k, v in prof.items()}
    root_n = stats["root_cols"] + stats["delayed_to_root"]
    flops_front, launches = update_flops(root_n, nth, 64)
    peak64, src = fp64_peak()
    ach = flops_front * nb / (per["update"] * 1e-3) / 1e12
    ok = tuple(inertia) == tuple(m.expected_inertia())
    del s
    torch.cuda.empty_cache()
    return {"workload": f"Model({nb},{nq},{ym},{nth}): {nb} blocks x {m.block_dim} rows, {nth} coupling vars (config-5 slice)",
            "e2e_ms": float(np.median(times[1:])), "symbolic_s": sym_s, "kernels_ms_per_step": per,
            "roofline": {"kernel": "front_update_kernel", "bound": "tensor", "achieved": ach, "peak": peak64,
                         "unit": "TFLOP/s", "frac": ach / peak64, "peak_source": src,
                         "algorithmic_flops_per_step": flops_front * nb, "launches_per_step": launches},
            "check": {"inertia_matches_closed_form": bool(ok), "max_err": float(m.check_result(x))}}


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the wide-border (config-5 slice) leg")
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
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)  # 256 MB > 126 MB of L2

    def e2e_step():
        flush.fill_(1.0)  # evict L2 between steps (inside the timed region: ~0.1 ms, counted against us)
        res = solver.do_numeric_factorization(kkt)
        ine = solver.get_inertia()
        sol = solver.do_back_solve(rhs)
        return res, ine, sol

    for _ in range(args.warmup):
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

    # ---- device-resident hot path (`value`) --------------------------------------------------------
    values_dev = be.values_pin.to(dev)
    rhs_dev = be.rhs_pin.to(dev)
    rhsc_dev = be.rhsc_pin.to(dev)
    x_dev = torch.empty_like(rhs_dev)
    xc_dev = torch.empty_like(rhsc_dev)

    def dev_step():
        flush.fill_(1.0)
        code, s_local = be.numeric_local_device(values_dev)
        comm.allreduce_sum_(s_local)
        code2 = be.numeric_coupling(s_local)
        loc = be.inertia_local()
        cpl = be.inertia_coupling()
        be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_)
        return code, code2, loc, cpl

    sampler = ClockSampler(local_rank) if rank == 0 else None  # NVML init takes milliseconds: before the barrier
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

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    comm.allreduce_max_(t)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------
    per_step = {k: v["ms"] / args.steps for k, v in prof.items()}
    dominant = max(per_step, key=per_step.get)
    stats = be.plan_stats(0)
    nnz_in = int((st.dest_front >= 0).sum())                       # lower-triangle input entries (all local blocks + Q)
    nnz_l = sum(be.plan_stats(b)["nnz_l_subtree"] for b in range(st.n_local))
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
                   "coupling": N_THETA, "l2": "L2 flushed by a 256 MB device write before every step (inside the timed region)",
                   "parallelism": f"blocks round-robin over {world} GPU(s); Schur complement + coupling rhs all-reduced (NCCL)"},
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "wall_ms": e2e_wall_ms, "api": "B200SchurComplementLinearSolver.do_numeric_factorization + get_inertia + do_back_solve"},
        "gpu_launches": int(launches),
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
    if not args.no_secondary and world == 1:
        del solver, be
        torch.cuda.empty_cache()
        line["secondary"] = wide_border_leg(dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
