#!/usr/bin/env python
"""bench.py -- GP-predict + acquisition candidates/second on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic candidates: k* build, L^-1 k* contraction,
posterior mean / variance, q acquisition criteria and their arg-max.  Default workload C3 (the config the
metric is quoted on): N=4096, D=16, Matern-5/2 ARD, MGFI q=32, M = 1e7 candidates per step -- ALL of them on one GPU
at --gpus 1, sharded M / N per rank at N GPUs (strong scaling, what BASELINE.json's north_star states).  A weak-scaling
sub-record (1.25e6 candidates per GPU per step at every N), a float64-path sub-record and the fit timings ride in the
same JSON line.  The fit (assembly + Cholesky + L^-1 + solves) runs once before the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5] [--scaling strong|weak] [--impl reference]
Under torchrun (N>1) every rank scores its own shard; the one exchange is a single all-reduce of the q (value, index)
pairs, written on the device and overlapped with the next step's kernel (sharded.ArgmaxExchange).
"""
from __future__ import annotations

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

METRIC = "gp_predict_acq_candidates_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum PER CANDIDATE of the fused kernel, from ncu --set full captures of one launch
# of that workload (profiles/): keyed (workload, products per MAC, kernel generation).  Generation 5 keeps a 148 MB r
# scratch at C3 (does not fit L2: profiles/r01/gen5_final_ncu_summary.txt, 75776 candidates); generation 6 halves it
# (profiles/r02/gen6_final_ncu_summary.txt, 151552 candidates).  ncu-derived, not measured in-run.
NCU_TRAFFIC_PER_CAND = {
    ("C3", 1, 5): (2.047070e9 + 649.5e6) / 75776,
    ("C3", 1, 6): (1.667662e9 + 1.230796e9) / 151552,
}
UNIT = "candidates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m-per-gpu", type=int, default=0, help="override candidates per GPU per step")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the config's M_total per step over all GPUs (default); weak: M_total / config GPUs per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the weak / fp64 / fit sub-records")
    ap.add_argument("--precision", default="fast", choices=["fast", "fp64"],
                    help="fast: tcgen05 split-fp16 pass + exact fp64 re-score of the arg-max band; fp64: DMMA parity path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy/scipy path (the reference itself is pure Python and
# cannot travel to the GPU box; oracle/gp_oracle.py restates it statement by statement)
# --------------------------------------------------------------------------------------------------
def cpu_chunk(N, D):
    return max(64, min(8192, (512 << 20) // (N * D * 8)))  # chunk * N * D * 8 B <= 512 MB


def cpu_fit(w):
    from oracle import gp_oracle as go

    X, y, theta = go.canonical_problem(w.N, w.D)
    corr = go.CORR_NAMES[w.corr]
    return go.fit_fixed(X, y, corr, theta, go.MODE_NOISY, sigma2=1.0, noise_var=w.nugget), go


def cpu_step(ora, go, w, params, Xc):
    """predict(eval_MSE=True) in chunks + q vectorised criteria + arg-max, like one GPU step"""
    from bayesian_optimization_b200 import workloads as wl

    yh, ms = go.predict_chunked(ora, Xc, cpu_chunk(w.N, w.D))
    pl = go.plugin_value(ora.y, True)
    acq = wl.ACQ_IDS[w.acq]
    best = [int(np.argmax(go.acquisition(acq, yh, ms, ora.sigma2, pl, p, True))) for p in params]
    return best


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host core"""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def threads_used():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, w, params):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bayesian_optimization_b200 import workloads as wl

    use_all_host_threads()
    sample = args.cpu_sample or 4 * cpu_chunk(w.N, w.D)
    ora, go = cpu_fit(w)
    Xc = wl.canonical_candidates(sample, w.D)
    for _ in range(args.warmup):
        cpu_step(ora, go, w, params, Xc[: max(64, sample // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(ora, go, w, params, Xc)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    cores = threads_used()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.scaling == "weak" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.name, "describe": w.describe, "N": w.N, "D": w.D, "corr": w.corr, "acq": w.acq,
                   "q": w.q, "candidates_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} candidates/step (bounded sample of the {w.M_total}-candidate step; the path is linear in M), "
                                   f"numpy/scipy oracle port of the reference, {cores} BLAS threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=REAL_STDOUT, flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def executed_macs(w, nprod):
    """tensor-core MACs per candidate the fused kernel EXECUTES (vs the algorithmic N^2 / 2): nprod fp16 products per
    MAC over the accumulator super-tiles (384 columns = a 256- and a 128-column block; blocks above the diagonal skipped
    per 64-wide chunk) plus the Gram MMAs (64 x 16 ceil(D/16) per chunk, always three products).
    absolute_exponential runs generation 1 (512-column super-tiles of two 256-column blocks, no Gram MMA)."""
    ld = -(-w.N // 128) * 128
    gen2 = w.corr != "absolute_exponential"
    gen5 = gen2 and ld >= 512 and int(os.environ.get("B200BO_FAST_KERNEL", "6")) >= 5
    mac = 0
    if gen5:
        for s_ in range(-(-ld // 384)):
            n0, kext = 384 * s_, min(ld, 384 * (s_ + 1))
            for k0 in range(0, kext, 64):
                mac += 64 * (256 if k0 < n0 + 128 else 128 if k0 < n0 + 256 else 0)
                if n0 + 256 < ld:
                    mac += 64 * 128
        gram = (ld // 64) * 64 * 16 * (-(-w.D // 16))
        return nprod * mac + 3 * gram, gen5, gen2
    if gen2:
        gram = 0
        for s_ in range(-(-ld // 384)):
            n0, kext = 384 * s_, min(ld, 384 * (s_ + 1))
            mac += 256 * min(kext, n0 + 256)
            if n0 + 256 < ld:
                mac += 128 * kext
            gram += (kext // 64) * 64 * 16 * (-(-w.D // 16))
        return nprod * mac + 3 * gram, gen5, gen2
    for s_ in range(-(-ld // 512)):
        kext = min(ld, 512 * (s_ + 1))
        for j_ in range(2):
            n0 = 512 * s_ + 256 * j_
            if n0 < ld:
                mac += 256 * min(kext, n0 + 256)
    return 3 * mac, gen5, gen2


def fit_records(b2, wl, w, local):
    """what a fit costs on the device: one likelihood + gradient evaluation at the workload's shape, and the full
    L-BFGS-B fit() the survey timed upstream (N=1024, D=8, RBF, nugget 1e-6, one start, theta0 = 1,
    eval_budget = 60 -> 25.0 s and 43 evaluations on 8 host cores)."""
    out = {}
    X, y, theta = wl.canonical_problem(w.N, w.D)
    gp = b2.GaussianProcess(mean=b2.constant_trend(w.D), corr=w.corr, thetaL=[1e-5] * w.D, thetaU=[1e2] * w.D,
                            nugget=w.nugget, device=local)
    gp.fit_fixed(X, y, theta, 1.0)
    par = np.r_[theta, 1.0]
    try:
        gp.log_likelihood_concentrated(par, eval_grad=True)
        t0 = time.perf_counter()
        for _ in range(3):
            gp.log_likelihood_concentrated(par, eval_grad=True)
        out["fit_llf_grad_ms"] = 1e3 * (time.perf_counter() - t0) / 3
    except Exception as e:  # Matern-5/2 has no theta-gradient upstream (gpr.py:758-759): value only
        t0 = time.perf_counter()
        for _ in range(3):
            gp.log_likelihood_concentrated(par)
        out["fit_llf_ms"] = 1e3 * (time.perf_counter() - t0) / 3
        out["fit_llf_grad_ms"] = None
        out["fit_llf_grad_note"] = str(e)[:120]
    del gp
    X, y, _ = wl.canonical_problem(1024, 8)
    g2 = b2.GaussianProcess(mean=b2.constant_trend(8), corr="squared_exponential", theta0=[1.0] * 8, thetaL=[1e-5] * 8,
                            thetaU=[1e2] * 8, nugget=1e-6, random_start=1, eval_budget=60, device=local)
    np.random.seed(42)
    g2.fit(X, y)          # first call: builds the engine, warms every kernel
    np.random.seed(42)
    g3 = b2.GaussianProcess(mean=b2.constant_trend(8), corr="squared_exponential", theta0=[1.0] * 8, thetaL=[1e-5] * 8,
                            thetaU=[1e2] * 8, nugget=1e-6, random_start=1, eval_budget=60, device=local)
    g3._engine = g2._engine  # same handle: buffers and graphs are warm, as in a BO loop that refits every iteration
    t0 = time.perf_counter()
    g3.fit(X, y)
    out["fit_full_s"] = time.perf_counter() - t0
    out["fit_full"] = {"N": 1024, "D": 8, "corr": "squared_exponential", "evals": int(g3.eval_count), "llf": float(g3.log_likelihood_),
                       "theta_mean": float(np.mean(g3.theta_)), "reference_s": 25.0,
                       "reference_note": "SURVEY.md section 6 probe: upstream fit(), 43 evaluations, 8 host cores"}
    return out


def run_b200(args, w, params):
    import torch
    import torch.distributed as dist

    import bayesian_optimization_b200 as b2
    from bayesian_optimization_b200 import sharded, workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    M_weak = w.M_total // w.gpus                      # the config's per-GPU share at the GPU count it is quoted on
    if args.m_per_gpu:
        M, lo = args.m_per_gpu, rank * args.m_per_gpu
    elif args.scaling == "strong":
        lo, hi = sharded.shard_bounds(w.M_total, world, rank)
        M = hi - lo
    else:
        M, lo = M_weak, rank * M_weak
    M_all = world * M if (args.m_per_gpu or args.scaling == "weak") else w.M_total
    acq_id = wl.ACQ_IDS[w.acq]

    # ---- fit once (replicated per rank, deterministic) -------------------------------------------
    X, y, theta = wl.canonical_problem(w.N, w.D)
    gp = b2.GaussianProcess(mean=b2.constant_trend(w.D), corr=w.corr, thetaL=[1e-5] * w.D, thetaU=[1e2] * w.D,
                            nugget=w.nugget, device=local)  # one engine per rank, on this rank's GPU
    t0 = time.perf_counter()
    llf = gp.fit_fixed(X, y, theta, 1.0)
    fit_wall_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    gp.fit_fixed(X, y, theta, 1.0)
    fit_wall_ms2 = 1e3 * (time.perf_counter() - t0)
    fit_t = gp.engine.fit_timings()
    eng = gp.engine
    fast = args.precision == "fast"
    eng.set_precision(b2._lib.PREC_FAST if fast else b2._lib.PREC_FP64)
    # ImprovementBased.plugin: min of the standardised y (bayes_opt.py:18-25).  NoisyBO (C4) would take min(predict(X))
    # (:185-194) -- UCB has no plug-in, so the value is unused there
    plugin = float(np.min(gp.y))

    # ---- candidates: pinned host buffer (e2e) and a device-resident copy (value); rank g draws shard g's stream -------
    xh = torch.empty((M, w.D), dtype=torch.float64, pin_memory=True)
    wl.canonical_candidates(M, w.D, shard=rank, out=xh.numpy())
    xd = xh.to(dev)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    exch = sharded.ArgmaxExchange(w.q, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(x, steps, m_rows=None):
        """`steps` steps over x (device tensor or pinned host array); the exchange of step k is consumed after step
        k + 1 has been launched.  -> (ms max over ranks, summed engine timings, last merged result)"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kern = np.zeros(b2._lib.N_TIMINGS)
        pending, out = None, None
        for _ in range(steps):
            eng.acq(x, acq_id, True, plugin, params)     # this rank's shard: fused pass + band, results stay on the device too
            kern += eng.timings()
            ticket = exch.submit(eng, lo)                 # pairs written on the device, ONE all-reduce, async copy back
            if pending is not None:
                out = pending.result()
            pending = ticket
        out = pending.result()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kern, out

    for _ in range(args.warmup):
        eng.acq(xd, acq_id, True, plugin, params)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, kern, best = timed(xd, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    band = eng.band_info() if fast else None
    check = eng.fast_check(stride=100) if fast else None   # float64 check of 1 % of this rank's candidates (untimed)
    for _ in range(max(1, args.warmup // 3)):
        eng.acq(xh.numpy(), acq_id, True, plugin, params)
    ms_e2e, _, best2 = timed(xh.numpy(), args.steps)
    assert list(best[1]) == list(best2[1]), "device-resident and host-buffer paths disagree on the arg-max"

    # ---- sub-records (untimed part of the run) -----------------------------------------------------------
    extras = {}
    if not args.no_extras:
        if args.scaling == "strong" and not args.m_per_gpu and M_weak != M:
            # weak scaling next to the strong default: the config's per-GPU share on every rank
            steps_w = max(3, min(args.steps, 10))
            xw = xd[:M_weak] if M_weak <= M else torch.from_numpy(wl.canonical_candidates(M_weak, w.D, shard=rank)).to(dev)
            eng.acq(xw, acq_id, True, plugin, params)
            ms_w, _, _ = timed(xw, steps_w)
            extras["weak_scaling"] = {"value": world * M_weak * steps_w / (ms_w * 1e-3), "unit": UNIT, "steps": steps_w,
                                      "ms_per_step": ms_w / steps_w, "candidates_per_gpu_per_step": M_weak, "scaling": "weak"}
            del xw
        if world == 1 and fast:
            # the drop-in DEFAULT of the Python classes (precision="fp64", values returned to 1e-9): fp64 DMMA path
            m64 = min(M, 1_250_000)
            eng.set_precision(b2._lib.PREC_FP64)
            eng.acq(xd[:m64], acq_id, True, plugin, params)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                eng.acq(xd[:m64], acq_id, True, plugin, params)
            barrier()
            dt = (time.perf_counter() - t0) / 2
            extras["fp64_path"] = {"value": m64 / dt, "unit": UNIT, "ms_per_step": 1e3 * dt, "candidates_per_step": m64,
                                   "kernel": "kstar_kernel + contract_fp64_kernel (fp64 DMMA) + acq_kernel",
                                   "note": "what GaussianProcess.predict / acquisition __call__ run by default"}
            eng.set_precision(b2._lib.PREC_FAST)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if not args.no_extras:
        try:
            extras["fit"] = fit_records(b2, wl, w, local)
        except Exception as e:  # never lose the headline line to a sub-record
            extras["fit"] = {"error": repr(e)[:200]}
    value = M_all * args.steps / (ms_dev * 1e-3)
    e2e = M_all * args.steps / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (the L^-1 k* contraction) -----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    n_contract = kern[4]
    launch_ms = kern[2] / max(n_contract, 1)
    cand_per_launch = M * args.steps / max(n_contract, 1)
    flops_per_cand = float(w.N) ** 2  # mul+add over the lower triangle of L^-1
    achieved = cand_per_launch * flops_per_cand / (launch_ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)",
        "traffic": None, "flops_per_candidate": flops_per_cand, "avg_launch_ms": launch_ms,
        "candidates_per_launch": cand_per_launch, "share_of_step": kern[2] / max(kern[0], 1e-9),
        "hbm_frac": value / world * (8 * w.D) / 1e9 / peaks.get("hbm_gbs", 6550.0),
        "whole_step_frac": (M * flops_per_cand / (ms_dev / args.steps * 1e-3) / 1e12) / peak,
    }
    if fast:
        nprod = int(round(kern[8] / args.steps))
        mac_exec, gen5, gen2 = executed_macs(w, nprod)
        exe = cand_per_launch * 2.0 * mac_exec / (launch_ms * 1e-3) / 1e12
        # DRAM traffic of the fused kernel per launch: ncu's per-candidate figure (dram__bytes_read.sum + dram__bytes_write.sum
        # of one --set full capture of this workload, profiles/) x the candidates of one launch
        gen = int(round(kern[10] / args.steps)) if len(kern) > 10 else 0
        tpc = NCU_TRAFFIC_PER_CAND.get((w.name, nprod, gen))
        roofline["traffic"] = tpc * cand_per_launch if tpc else None
        roofline["traffic_source"] = "ncu --set full capture of this workload (profiles/), bytes per candidate x candidates per launch" if tpc else None
        roofline.update({
            "generation": gen,
            "kernel": ("predict_fused_shared_kernel (tcgen05.mma cta_group::2 kind::f16, M=256, two CTA pairs share a candidate tile: r computed once per group, replayed by TMA from an L2-sized scratch" if gen == 6
                       else "predict_fused_decoupled_kernel (tcgen05.mma cta_group::2 kind::f16, M=256, r computed once per tile and replayed by TMA" if gen5
                       else "predict_fused_pair_kernel (tcgen05.mma cta_group::2 kind::f16, M=256" if gen2
                       else "predict_fused_tc_kernel (tcgen05.mma kind::f16, M=128") + ", fp32 TMEM accumulators)",
            "executed_tensor_tflops": exe, "executed_frac": exe / peak,
            "products_per_mac": nprod,
            "note": "achieved counts ALGORITHMIC flops (N^2 per candidate); the tensor pipe executes products_per_mac x that "
                    "(+ diagonal blocks and the Gram MMAs); executed_frac is the tensor-pipe utilisation",
        })
    else:
        roofline.update({
            "kernel": "contract_fp64_kernel (fp64 DMMA: tcgen05 has no f64 kind)",
            "fp64_nominal_tflops": 40.0, "frac_of_fp64_nominal": achieved / 40.0,
        })
    scaling = "weak" if (args.scaling == "weak" or args.m_per_gpu) else "strong"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f16/f32-acc + f64 re-score" if fast else "f64", "data": "synthetic",
        "config": {"workload": w.name, "describe": w.describe, "N": w.N, "D": w.D, "corr": w.corr, "acq": w.acq,
                   "q": w.q, "candidates_per_gpu_per_step": M, "candidates_per_step": M_all,
                   "parallelism": f"candidate-shards x{world}",
                   "precision": (f"tcgen05 fp16 first pass ({int(round(kern[8] / args.steps))} product(s) per MAC, fp32 TMEM "
                                 "accumulators) + exact fp64 re-score of the arg-max band" if fast
                                 else "fp64 DMMA parity path"),
                   "rescored_per_step": kern[6] / args.steps, "band_passes_per_step": kern[7] / args.steps,
                   "l2": "inputs_exceed_l2 (candidates + k* workspace > 126 MB per step)",
                   "plugin": "min(y) (bayes_opt.py:18-25); NoisyBO's min(predict(X)) is pinned in tests/golden/canonical_big.npz, UCB ignores it",
                   "exchange": "one int64-sum all-reduce of (world, 2q) pairs written on the device, consumed one step later",
                   "fit_ms_device": fit_t[0], "fit_ms_wall_first": fit_wall_ms, "fit_ms_wall": fit_wall_ms2,
                   "fit_split_ms": {"assemble": fit_t[1], "cholesky": fit_t[2], "trtri": fit_t[3], "solves": fit_t[4]},
                   "llf": llf, "argmax": [int(i) for i in best[1][:4]]},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": M * w.D * 8, "d2h_bytes_per_step": 16 * w.q * world + 64,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(kern[5]),
        "kernel_ms_per_step": {"total_device": kern[0] / args.steps, "kstar": kern[1] / args.steps,
                               "contract_or_fused": kern[2] / args.steps, "acq_argmax_or_band": kern[3] / args.steps},
        "roofline": roofline,
    }
    if fast:
        line["fast_pass_check"] = {
            "sample": f"every 100th candidate of rank 0's shard re-evaluated in float64 on the device ({int(check['checked'])} candidates)",
            "max_err_yhat": check["max_err_yhat"], "max_err_mse": check["max_err_mse"],
            "half_width_yhat": check["dy"], "half_width_mse_model_at_ss1": check["ds_model_at_ss1"],
            "half_width_mse_calibrated": check["ds_cal"], "max_error_over_allowed": check["max_ratio_to_allowed"],
            "deterministic_bound_mse": band["ds_deterministic"], "band_max_error_over_allowed": band["band_ratio"],
        }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        sample = args.cpu_sample or 16 * cpu_chunk(w.N, w.D)
        ora, go = cpu_fit(w)
        Xc = wl.canonical_candidates(sample, w.D)
        cpu_step(ora, go, w, params, Xc[: sample // 8])
        t0 = time.perf_counter()
        cb = cpu_step(ora, go, w, params, Xc)
        dt = time.perf_counter() - t0
        bv, bi, _ = eng.acq(Xc, acq_id, True, plugin, params)
        cores = threads_used()
        line["cpu_baseline"] = {
            "value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} of the {M} candidates of one step, oracle port of the reference's numpy/scipy path, "
                      f"{cores} BLAS threads, {dt:.1f} s",
            "argmax_matches_gpu": [int(i) for i in bi] == cb,
        }
    print(json.dumps(line), file=REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


REAL_STDOUT = sys.stdout


def claim_stdout():
    """stdout carries exactly ONE JSON line: route everything libraries print to fd 1 (NCCL's version banner, ...) to
    stderr and keep a private handle on the real stdout for the result"""
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    args = parse()
    claim_stdout()
    from bayesian_optimization_b200 import workloads as wl

    w = wl.WORKLOADS[args.workload]
    params = wl.acquisition_params(w)
    if args.impl == "reference":
        run_reference(args, w, params)
    else:
        run_b200(args, w, params)


if __name__ == "__main__":
    main()
