#!/usr/bin/env python
"""
bench.py — headline benchmark of the B200-native RV log-likelihood path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): RV lnL evals/sec (K-planet, batched), plus the fraction of the FP64
roofline.  Workload at every N: BASELINE.json configs[1] -- 2-planet eccentric + linear drift,
2 instruments, 1000 epochs, one vectorised batch of ndraw = 4096 parameter vectors per step and
per GPU (weak scaling: rows sharded over ranks, lnL all-gathered over NCCL inside the step).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with theta resident in HBM;
`e2e` = the same metric through the public host-buffer API (pinned host theta -> H2D -> kernel ->
D2H lnL) ; `roofline` = algorithmic FP64 flops of the likelihood kernel / its CUDA-event time /
the FP64 peak measured in the same run by a register-resident DFMA loop (MEASURED_PEAKS.json has
no FP64 row); `cpu_baseline` = the CPU oracle (numpy restatement + the reference's own C Kepler
solver from oracle/_ref) on the box's host cores.

`--impl reference` times that CPU implementation alone, on the same config/metric.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RV lnL evals/sec (K-planet, batched)"
UNIT = "lnL/s"
CONFIG_ID = 2
BATCH = 4096


def algorithmic_flops(n_epochs, n_planets, drift_order, mean_iters):
    """F = N [K (51 + 46 I) + 48 + 2 d] FP64 flops per lnL."""
    return n_epochs * (n_planets * (51.0 + 46.0 * mean_iters) + 48.0 + 2.0 * drift_order)


def workload_config(case, batch, world):
    return {"workload": f"config{CONFIG_ID}: {case.n_planets}-planet Keplerian + "
                        f"{'linear drift' if case.drift else 'no drift'}, {case.n_inst} instruments, "
                        f"{case.n_epochs} epochs, batch {batch} theta per step per GPU "
                        f"(UltraNest vectorized ndraw={batch})",
            "n_epochs": case.n_epochs, "n_planets": case.n_planets, "n_inst": case.n_inst,
            "ndim": case.ndim, "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": f"rows sharded x{world}, lnL all-gather",
            "l2": "flushed (256 MiB write) between timed steps"}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure) timed on the host cores
# ------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(fixed, tables, parnames):
    from oracle.rv_oracle import OracleRVModel
    _W["m"] = OracleRVModel(fixed, tables, parnames)


def _cpu_eval(block):
    return _W["m"].log_likelihood_batch(block)


def cpu_rate(case, theta, cores, budget_s=12.0):
    """lnL/s of the CPU oracle on `cores` processes over a bounded sample of `theta`."""
    import multiprocessing as mp
    from oracle import rv_oracle
    rv_oracle.build()
    kind = rv_oracle.solver_kind()
    tables = case.datadict()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init,
                  initargs=(case.fixedpardict, tables, case.parnames)) as pool:
        probe = theta[: max(cores * 8, 64)]
        t0 = time.perf_counter()
        pool.map(_cpu_eval, np.array_split(probe, cores))
        rate0 = len(probe) / (time.perf_counter() - t0)
        probe2 = np.tile(theta, (-(-int(rate0) // len(theta)), 1))[: max(len(probe), int(rate0))]
        t0 = time.perf_counter()  # longer probe: the first one is dominated by start-up costs
        pool.map(_cpu_eval, np.array_split(probe2, cores * 4))
        rate0 = len(probe2) / (time.perf_counter() - t0)
        n = int(max(len(probe), rate0 * budget_s))
        reps = -(-n // len(theta))
        sample = np.tile(theta, (reps, 1))[:n]  # the step's rows, repeated to fill the budget
        t0 = time.perf_counter()
        pool.map(_cpu_eval, np.array_split(sample, cores * 4))
        dt = time.perf_counter() - t0
    return n / dt, n, dt, kind


def run_reference(args, rank, world, emit):
    """The CPU arm: the oracle on all host cores; every step is a bounded sample of the workload,
    sized so that warmup + steps together take about two minutes."""
    if rank != 0:
        return
    import multiprocessing as mp
    from evidence_b200 import synth
    from oracle import rv_oracle
    rv_oracle.build()
    kind = rv_oracle.solver_kind()
    case = synth.make_case(CONFIG_ID)
    cores = os.cpu_count() or 1
    theta = case.draw_theta(BATCH, seed=1000)
    n_steps = args.warmup + args.steps
    per_step_s = min(12.0, max(0.25, 120.0 / max(1, n_steps)))
    rates, n = [], 0
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init,
                  initargs=(case.fixedpardict, case.datadict(), case.parnames)) as pool:
        probe = theta[: max(cores * 8, 64)]
        t0 = time.perf_counter()
        pool.map(_cpu_eval, np.array_split(probe, cores))
        rate0 = len(probe) / (time.perf_counter() - t0)
        probe2 = np.tile(theta, (-(-int(rate0 * 2) // len(theta)), 1))[: max(cores * 8, int(rate0 * 2))]
        t0 = time.perf_counter()  # second, longer probe: the first is dominated by start-up costs
        pool.map(_cpu_eval, np.array_split(probe2, cores * 4))
        rate0 = len(probe2) / (time.perf_counter() - t0)
        n = int(max(cores * 4, rate0 * per_step_s))
        sample = np.tile(theta, (-(-n // len(theta)), 1))[:n]
        chunks = np.array_split(sample, cores * 4)
        for k in range(n_steps):
            t0 = time.perf_counter()
            pool.map(_cpu_eval, chunks)
            dt = time.perf_counter() - t0
            if k >= args.warmup:
                rates.append(n / dt)
    value = float(np.mean(rates)) if rates else 0.0
    sample_txt = (f"{n} lnL evaluations per step (rows of one {BATCH}-theta batch, repeated) on {cores} "
                  f"processes; numpy restatement of RVModel.log_likelihood; Kepler solver = "
                  f"{'the reference trueanomaly.c compiled to oracle/_ref' if kind == 'reference' else 'C restatement'}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / value if value else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(case, BATCH, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample_txt},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "power_w_max": float(np.max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def sweep(model_factory, case, batch, steps, torch, peak_tflops):
    """A larger parameter sweep of another BASELINE shape (extra information, N=1 only)."""
    model = model_factory(case)
    model.set_option("timing", 1)
    theta = torch.from_numpy(case.draw_theta(batch, seed=77)).cuda()
    out = torch.empty(batch, dtype=torch.float64, device="cuda")
    for _ in range(2):
        model.log_likelihood_device(theta, out=out)
    torch.cuda.synchronize()
    model.reset_counters()
    kms = []
    for _ in range(steps):
        model.log_likelihood_device(theta, out=out)
        kms.append(model.last_kernel_ms())
    c = model.counters()
    mean_it = c["n_newton_iters"] / max(1, c["n_solves"])
    F = algorithmic_flops(case.n_epochs, case.n_planets, case.drift, mean_it)
    ms = float(np.mean(kms))
    rate = batch / (ms * 1e-3)
    ach = rate * F / 1e12
    res = {"workload": f"N={case.n_epochs} K={case.n_planets} inst={case.n_inst} batch={batch}",
           "lnl_per_s": rate, "kepler_solves_per_s": rate * case.n_epochs * case.n_planets,
           "kernel_ms": ms, "mean_newton_iters": mean_it, "achieved_tflops": ach,
           "frac_of_fp64_peak": ach / peak_tflops if peak_tflops else None}
    model.close()
    return res


def stress(model_factory, torch, peak_tflops, batch=10_000_000):
    """BASELINE.json configs[4]: raw kernel stress, 1e7 parameter vectors x 1e4 epochs x 3 planets,
    Kepler solves/s against the FP64 peak.  Everything stays on the device: U ~ torch.rand, the
    prior transform fused in front of the likelihood (rvl_transform_loglike_dev)."""
    from evidence_b200 import synth
    case = synth.make_case(5)
    model = model_factory(case)
    model.set_priors(case.priordict)
    model.set_option("timing", 1)
    gen = torch.Generator(device="cuda").manual_seed(5)
    U = torch.rand((100_000, case.ndim), dtype=torch.float64, device="cuda", generator=gen)
    model.transform_loglike_device(U)  # warm launch
    torch.cuda.synchronize()
    U = torch.rand((batch, case.ndim), dtype=torch.float64, device="cuda", generator=gen)
    theta = torch.empty_like(U)
    lnl = torch.empty(batch, dtype=torch.float64, device="cuda")
    model.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.transform_loglike_device(U, theta=theta, lnl=lnl)
    e1.record()
    torch.cuda.synchronize()
    total_ms, k_ms = e0.elapsed_time(e1), model.last_kernel_ms()
    c = model.counters()
    mean_it = c["n_newton_iters"] / max(1, c["n_solves"])
    F = algorithmic_flops(case.n_epochs, case.n_planets, case.drift, mean_it)
    ach = batch * F / (k_ms * 1e-3) / 1e12
    finite = bool(torch.isfinite(lnl).all().item())
    res = {"workload": f"configs[4] stress: N={case.n_epochs} K={case.n_planets} batch={batch}, "
                       "u -> theta -> lnL on the device",
           "kepler_solves": int(c["n_solves"]), "kernel_ms": k_ms, "total_ms_with_prior_transform": total_ms,
           "kepler_solves_per_s": c["n_solves"] / (k_ms * 1e-3), "lnl_per_s": batch / (k_ms * 1e-3),
           "mean_newton_iters": mean_it, "achieved_tflops": ach,
           "frac_of_fp64_peak": ach / peak_tflops if peak_tflops else None,
           "newton_cap_hits": int(c["n_cap_hits"]), "all_finite": finite}
    model.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu baseline and sweeps")
    ap.add_argument("--gather", default="fused", choices=["fused", "fused-barrier", "nccl"],
                    help="multi-GPU: the all-gather fused into the likelihood launch over NVLink "
                         "symmetric memory (peer stores + completion flags; default), the same with "
                         "a symmetric-memory barrier, or NCCL all_gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner,
    # torchrun notices) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from evidence_b200 import build, synth
    from evidence_b200.multigpu import ShardedLikelihood
    from evidence_b200.rvmodel import RVModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU arm")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def make_model(case):
        return RVModel(case.fixedpardict, case.datadict(), case.parnames, device=local_rank)

    case = synth.make_case(CONFIG_ID)
    model = make_model(case)
    B, K, W = args.batch, args.steps, args.warmup
    theta_host = case.draw_theta(B, seed=1000 + rank)
    theta = torch.from_numpy(theta_host).cuda()
    lnl = torch.empty(B, dtype=torch.float64, device="cuda")
    sharded = ShardedLikelihood(lambda blk: model.log_likelihood_device(blk, out=lnl), case.ndim)
    gather = "none" if world == 1 else "nccl all_gather"
    if world > 1 and args.gather != "nccl":
        try:  # all-gather fused into the producing kernel over NVLink symmetric memory
            from evidence_b200.multigpu import FusedGatherLikelihood
            sig = "flags" if args.gather == "fused" else "barrier"
            sharded = FusedGatherLikelihood(model, B, signal=sig)
            gather = ("fused into the likelihood launch (peer stores over NVLink symmetric memory + "
                      + ("completion flags)" if sig == "flags" else "symmetric-memory barrier)"))
        except Exception as exc:  # symmetric memory unavailable: NCCL all-gather
            print(f"fused gather unavailable ({exc!r}); using NCCL all_gather", file=sys.stderr)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        return sharded.evaluate_local(theta)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # At N = 1 a step IS one launch of the likelihood kernel (the work list, the per-point
    # constants and the slice sums all live inside it), so the step's CUDA events time the kernel;
    # with an all-gather in the step the library records its own events around the kernel.
    inner_events = world > 1
    model.set_option("timing", 1 if inner_events else 0)
    for _ in range(W):
        step()
    barrier()
    peak = max(model.fp64_peak_tflops(), model.fp64_peak_tflops())  # after warm-up: clocks are up
    model.reset_counters()
    launches0 = model.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(K)]
    kms = []
    clocks = ClockSampler(local_rank)
    barrier()
    t_wall = time.perf_counter()
    for k in range(K):
        flush.zero_()  # L2 flush, outside the per-step events
        ev[k][0].record()
        step()
        ev[k][1].record()
        if inner_events:
            kms.append(model.last_kernel_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    launches = model.launch_count() - launches0
    if not inner_events:
        if launches != K:
            raise SystemExit(f"expected one launch per step, counted {launches} in {K} steps")
        kms = [a.elapsed_time(b) for a, b in ev]
    cnt = model.counters()

    model.set_option("timing", 0)
    # ---- end to end: pinned host theta -> H2D -> kernel(s) -> D2H lnL, public host API ----
    th_pin = torch.from_numpy(theta_host).pin_memory()
    out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    th_np, out_np = th_pin.numpy(), out_pin.numpy()
    out_all_pin = torch.empty(B * world, dtype=torch.float64).pin_memory()

    def e2e_step():
        if world == 1:
            model.log_likelihood_batch(th_np, out=out_np)  # rvl_loglike: H2D, kernel, D2H, sync
        else:  # H2D, kernel, NCCL all-gather, D2H of the gathered vector (what a sampler sees)
            theta.copy_(th_pin, non_blocking=True)
            gathered = sharded.evaluate_local(theta)
            out_all_pin.copy_(gathered, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(W):
        e2e_step()
    barrier()
    # a full collection of the interpreter's heap (torch, numpy, ... imported) takes ~45 ms and
    # would land inside this wall-clock region: collect now, keep the collector off while timing
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    marks = [t0]
    for k in range(K):
        e2e_step()
        marks.append(time.perf_counter())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    e2e_us = np.diff(np.array(marks)) * 1e6  # per call (every call ends with a stream sync)
    clk = clocks.stop()  # the sampler covers both timed regions

    if world > 1:
        red = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(red[0]), float(red[1])

    mean_it = cnt["n_newton_iters"] / max(1, cnt["n_solves"])
    F = algorithmic_flops(case.n_epochs, case.n_planets, case.drift, mean_it)
    k_ms = float(np.mean(kms))
    traffic = None  # DRAM bytes per launch from the committed `ncu --set full` capture (profiles/)
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"config{CONFIG_ID}_B{B}"
        if key in prof:
            traffic = prof[key]["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    achieved = B * F / (k_ms * 1e-3) / 1e12
    value = world * B * K / (dev_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(case, B, world), gather=gather),
        "clocks": clk,
        "e2e": {"value": world * B * K / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(B * case.ndim * 8), "d2h_bytes_per_step": int(B * 8),
                "us_per_call_percentiles_1_50_99": [float(x) for x in np.percentile(e2e_us, [1, 50, 99])],
                "us_per_call_max": float(e2e_us.max())},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": "rv_lnl_kernel", "kernel_ms": k_ms,
                     "kernel_ms_source": ("library CUDA events around the kernel" if inner_events else
                                          "the step's CUDA events (one launch per step)"),
                     "flops_per_lnl": F, "mean_newton_iters": mean_it,
                     "peak_source": "DFMA loop measured in this run (rvl_fp64_peak); "
                                    "MEASURED_PEAKS.json has no FP64 row",
                     "hbm_bytes_per_lnl": 8 * (case.ndim + 1)},
        "wall_s_timed_region": t_wall,
        "kepler_solves_per_s": value * case.n_epochs * case.n_planets,
        "newton_cap_hits": cnt["n_cap_hits"],
    }

    if rank == 0 and world == 1 and not args.no_extras:
        try:
            sweeps = []
            sweeps.append(sweep(make_model, synth.make_case(3), 65536, 5, torch, peak))
            sweeps.append(sweep(make_model, synth.make_case(5), 32768, 3, torch, peak))
            line["sweeps"] = sweeps
            line["stress"] = stress(make_model, torch, peak)
        except Exception as exc:  # extras must never cost the headline line
            line["sweeps_error"] = repr(exc)
        try:
            cores = os.cpu_count() or 1
            rate, n, dt, kind = cpu_rate(case, theta_host, cores, budget_s=12.0)
            rate1, n1, dt1, _ = cpu_rate(case, theta_host, 1, budget_s=5.0)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n} lnL evaluations over the {B} theta rows of one step, {dt:.1f} s on {cores} processes; "
                          f"numpy restatement of RVModel.log_likelihood with "
                          f"{'the reference trueanomaly.c (oracle/_ref)' if kind == 'reference' else 'the C restatement of trueanomaly'}"
                          f"; single core: {rate1:.0f} lnL/s; the step's rows repeated to fill ~12 s",
                "single_core_value": rate1}
        except Exception as exc:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "failed: " + repr(exc)}
    if rank == 0:
        emit(line)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
