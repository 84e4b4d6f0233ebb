#!/usr/bin/env python
"""bench.py — lnpost evaluations/sec of the B200 hot path on the 10k-TOA GP pulsar (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference algorithm on host cores

One "step" = one pass of the hot path over one batch of 8192 synthetic walkers per GPU
(weak scaling: every rank evaluates its own 8192-walker shard of the model replica it holds, then
the per-walker lnpost vector is all-gathered over NCCL).  Workload = BASELINE.json configs[2]
("sim3_gp: PowerlawRedNoiseGP Woodbury likelihood, 10k TOAs, 30 Fourier harmonics, 8192 walkers"),
the configuration the headline metric is quoted on; data are synthetic (vela.jl_b200/synth.py).

Keys beyond the base contract: `roofline` (the kernel with the largest share of the step, against the FP64
peak of its pipe measured in this run; every kernel of the step is listed under `roofline.kernels`), `cpu_baseline` (the CPU oracle — a C/OpenMP restatement of
the reference algorithm, NOT Julia — on a bounded sample of the same workload), `e2e` (host buffers
through the public `Pulsar.lnpost_vectorized` call, H2D/D2H inside the timed region), `clocks`,
`gpu_launches`, `parity` (max relative lnlike difference vs the oracle on the CPU sample), `dense_sigma_path` (the same
workload with the structured Sigma path switched off: the dense FP64 tensor-core contraction and its roofline).
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
from __graft_entry__ import load_package  # noqa: E402

WORKLOAD = "config3_gp"
METRIC = "lnpost_evals_per_sec"
UNIT = "evals/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons, "samples": len(self.rows)}


def host_threads() -> int:
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is not it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def algorithmic_flops(ntoa: int, nrows: int, p: int):
    """per-evaluation FLOPs; returns (total, sigma_contraction_only)."""
    chain = 330 * ntoa
    sigma = nrows * p * (p + 1) + 3 * nrows * p
    rest = 3 * nrows + p**3 / 3 + p * p + 10 * p
    return chain + sigma + rest, sigma


def cpu_reference_arm(args, rank: int, world: int):
    """`--impl reference`: the reference's CPU algorithm (oracle port: the reference is Julia and cannot run
    here), all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    vb = load_package()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    from helpers import oracle_eval_factory
    ds = vb.synth.make_dataset(WORKLOAD, oracle_eval_factory(vb, O))
    md = vb.ModelDescriptor(ds.model, ds.toas)
    cores = host_threads()
    probe = ds.walkers(4 * cores, seed=99)
    t0 = time.perf_counter()
    O.lnlike_batch(md, probe, nthreads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    budget_s = 60.0 / max(args.steps + args.warmup, 1)
    sample = int(max(cores, min(8192, rate * min(budget_s, 15.0)) // cores * cores))
    W = ds.walkers(sample, seed=7)
    lnpr = np.zeros(sample)
    for _ in range(args.warmup):
        O.lnlike_batch(md, W, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = lnpr + O.lnlike_batch(md, W, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: PowerlawRedNoiseGP Woodbury lnpost, 10k TOAs, rank 60, 8192 walkers/GPU",
                   "ntoa": int(len(ds.toas)), "rank": int(ds.info["rank"]), "nfree": int(ds.info["nfree"])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} walkers of the {WORKLOAD} batch per step (C/OpenMP restatement of "
                                   "calc_lnpost_vectorized; the Julia reference cannot run in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "finite": bool(np.all(np.isfinite(out))),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=8192, help="walkers per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="after warm-up run ONE device-resident step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        cpu_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    vb = load_package()
    L = vb.lib.lib()

    # ---- synthetic dataset, generated with the CUDA library as the phase evaluator (same seed on every rank)
    t_gen = time.perf_counter()
    ds = vb.synth.make_dataset(WORKLOAD, lambda m, t: vb.Pulsar(m, t, devices=[local_rank]))
    psr = vb.Pulsar(ds.model, ds.toas, devices=[local_rank])
    log(f"[rank {rank}] dataset {ds.info} generated in {time.perf_counter() - t_gen:.1f}s")
    ntoa, p, nd = len(ds.toas), ds.info["rank"], ds.info["nfree"]
    nrows = ntoa
    B = args.walkers

    # priors for lnpost: flat boxes of +-50 sigma_par around the truth for timing parameters, the walker
    # distributions for noise parameters (host-side numpy, as in the reference where priors live on the host)
    priors = []
    for k in range(nd):
        if k in ds.noise_draw:
            kind, a, b = ds.noise_draw[k]
            priors.append(vb.priors.SimplePrior(f"p{k}", vb.priors.LogUniform(a, b) if kind == "loguniform" else
                                                vb.priors.Uniform(a - 10 * b, a + 10 * b) if kind == "normal" else
                                                vb.priors.LogNormal(np.log(a), max(b, 1e-3))))
        else:
            w = 50 * max(ds.sigma_par[k], 1e-300)
            priors.append(vb.priors.SimplePrior(f"p{k}", vb.priors.Uniform(ds.truth[k] - w, ds.truth[k] + w)))
    psr.set_priors(priors)          # uploads them to the device as well (all families have device twins)

    n_sets = 4   # rotate distinct walker batches so no step sees the previous step's inputs
    host_sets = [ds.walkers(B, seed=1000 * (rank + 1) + i) for i in range(n_sets)]
    host_lnpr = [np.ascontiguousarray(vb.priors.lnprior(priors, w)) for w in host_sets]
    dev_sets = [torch.from_numpy(w).cuda() for w in host_sets]
    dev_lnpr = [torch.from_numpy(lp).cuda() for lp in host_lnpr]
    d_out = torch.empty(B, dtype=torch.float64, device="cuda")
    gathered = torch.empty(B * world, dtype=torch.float64, device="cuda") if world > 1 else None
    # a dedicated non-default stream: the library launches on the stream it is given, and torch's events
    # (recorded on the current stream) must see those launches
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device(i):
        psr.lnpost_device(dev_sets[i % n_sets].data_ptr(), B, dev_lnpr[i % n_sets].data_ptr(), d_out.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`)
    for i in range(args.warmup):
        step_device(i)
    barrier()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStart()
        step_device(0)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        log("profiled one step")
        return
    psr.set_stage_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = psr.launch_count()
    stage_ms = np.zeros(5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = psr.launch_count() - launches0
    # per-stage durations: re-run K steps with a sync after each so the per-launch events can be read
    for i in range(args.steps):
        step_device(i)
        torch.cuda.synchronize()
        stage_ms += np.array(psr.last_stage_ms())
    stage_ms /= args.steps
    psr.set_stage_timing(False)
    finite_frac = float(torch.isfinite(d_out).double().mean().item())

    # ---- end-to-end timing through the public host API (`e2e`)
    def step_e2e(i):
        out = psr.lnpost_vectorized(host_sets[i % n_sets])     # pinned staging (priors on the device when every family has a device twin, else host prior vector) + H2D + kernels + D2H
        if world > 1:
            dist.all_gather_into_tensor(gathered, torch.from_numpy(out).cuda())
        return out

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        out_host = step_e2e(i)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # max over ranks
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = tt.tolist()

    if rank == 0:
        total_evals = B * world * args.steps
        value = total_evals / (dev_ms * 1e-3)
        e2e_value = total_evals / (e2e_ms * 1e-3)
        flops_eval, flops_sigma = algorithmic_flops(ntoa, nrows, p)
        # two occupancies of the same register-resident DMMA loop (64 and 8 warps/SM); the better one is the peak
        dmma_peak = max(float(L.vela_b200_fp64_peak_tflops(local_rank, 1)), float(L.vela_b200_fp64_peak_tflops(local_rank, 108)))
        dfma_peak = float(L.vela_b200_fp64_peak_tflops(local_rank, 0))
        sigma_ms = float(stage_ms[2])
        chain_ms, chol_ms = float(stage_ms[1]), float(stage_ms[3])
        structured = psr.sigma_structured()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        yw_bytes = 2 * B * nrows * 8   # y, w written by the chain kernel and read by the contraction
        # Per-kernel roofline entries.  "algorithmic" flops are SURVEY 8(d)'s per-evaluation counts x walkers per
        # launch; the structured Sigma path executes FEWER flops than the reference algorithm's contraction (it is an
        # algebraic identity on harmonic bases), so its entry is rated on the flops it executes and the
        # reference-equivalent rate is given separately.
        n_cols = None
        if structured:
            import ctypes as C
            info = (C.c_int32 * 3)()
            md0 = vb.ModelDescriptor(ds.model, ds.toas)
            L.vela_b200_debug_structured_sigma(md0.byref(), None, None, info)
            n_cols = int(info[1])
        nrows_pad = (nrows + 63) // 64 * 64
        sigma_exec = 2.0 * n_cols * nrows_pad if structured else float(flops_sigma)
        tf = lambda fl, ms: (fl * B / (ms * 1e-3) / 1e12) if ms > 0 else None   # noqa: E731
        # compute-bound kernels of this path run on the FP64 pipe; on B200 scalar DFMA and the DMMA tensor path share it
        # (measured peaks 36.5 vs 37.1 TFLOP/s), so "tensor" with the matching measured peak is the applicable roofline
        fp64_note = "FP64 pipe, scalar DFMA issue (shares the datapath of the FP64 tensor path on B200); peak = measured DFMA loop"
        kernels = [
            {"kernel": "toa_chain_kernel (per-model NVRTC build)" if psr.chain_specialized() else "toa_chain_kernel (generic)",
             "bound": "tensor", "bound_detail": fp64_note, "ms": chain_ms, "algorithmic_flops_per_launch": 330.0 * ntoa * B,
             "achieved": tf(330.0 * ntoa, chain_ms), "peak": dfma_peak, "unit": "TFLOP/s"},
            {"kernel": "colsum_kernel (structured Sigma, FP64 DMMA m8n8k4)" if structured else "sigma_contract_kernel (FP64 DMMA m8n8k4)",
             "bound": "tensor", "ms": sigma_ms, "algorithmic_flops_per_launch": sigma_exec * B,
             "achieved": tf(sigma_exec, sigma_ms), "peak": dmma_peak, "unit": "TFLOP/s",
             "reference_equivalent_tflops": tf(float(flops_sigma), sigma_ms)},
            {"kernel": "chol_finish_kernel", "bound": "tensor", "bound_detail": fp64_note, "ms": chol_ms,
             "algorithmic_flops_per_launch": (p**3 / 3 + p * p + 10 * p) * B,
             "achieved": tf(p**3 / 3 + p * p + 10 * p, chol_ms), "peak": dfma_peak, "unit": "TFLOP/s"},
        ]
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
        # very step (profiles/r1_ncu_full_config3_v9_final.txt); only valid for the default workload and kernel variants
        ncu_traffic = {"toa_chain_kernel (per-model NVRTC build)": 0.004324e9 + 1.261522e9,
                       "colsum_kernel (structured Sigma, FP64 DMMA m8n8k4)": 1.344508e9 + 0.026701e9,
                       "sigma_contract_kernel (FP64 DMMA m8n8k4)": 1.532375e9 + 0.116714e9,   # v2b_dense capture
                       "chol_finish_kernel": 0.049809e9 + 0.000072e9}
        # busy fraction of the bounding pipe from the same capture (sm__pipe_fp64_cycles_active / sm__pipe_tensor_cycles_active,
        # pct_of_peak_sustained_active): what the hardware counters say next to the algorithmic-FLOP fraction
        ncu_pipe = {"toa_chain_kernel (per-model NVRTC build)": 0.638,
                    "colsum_kernel (structured Sigma, FP64 DMMA m8n8k4)": 0.843,
                    "chol_finish_kernel": 0.202}
        for k in kernels:
            k["traffic"] = ncu_traffic.get(k["kernel"]) if (ntoa, p, B) == (10000, 60, 8192) else None
            k["pipe_active_ncu"] = ncu_pipe.get(k["kernel"]) if (ntoa, p, B) == (10000, 60, 8192) else None
            k["frac"] = (k["achieved"] / k["peak"]) if k["achieved"] and k["peak"] > 0 else None
            k["share_of_step"] = k["ms"] / float(stage_ms[4]) if stage_ms[4] > 0 else None
        dom = max(kernels, key=lambda k: k["ms"])
        roofline = {
            "bound": dom["bound"], "bound_detail": dom.get("bound_detail", "FP64 tensor cores (mma.sync m8n8k4 f64 -> DMMA)"),
            "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
            "unit": "TFLOP/s", "frac": dom["frac"], "traffic": dom["traffic"],
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                              "(profiles/r1_ncu_full_config3_v9_final.txt)",
            "peak_source": "measured in this run: register-resident FP64 loops (vela_b200_fp64_peak_tflops: DFMA for the "
                           "scalar kernels, mma.sync.m8n8k4.f64 for the DMMA kernels); MEASURED_PEAKS.json holds only "
                           "HBM and bf16 peaks",
            "fp64_fma_peak_tflops": dfma_peak, "fp64_dmma_peak_tflops": dmma_peak,
            "algorithmic_flops_per_launch": dom["algorithmic_flops_per_launch"],
            "kernel_ms": dom["ms"], "sigma_path": "structured (column sums)" if structured else "dense (pair contraction)",
            "kernels": kernels,
            "stage_ms": {"prologue": float(stage_ms[0]), "chain": chain_ms, "sigma_contract": sigma_ms,
                         "chol_finish": chol_ms, "all_kernels": float(stage_ms[4])},
            "hbm": {"algorithmic_bytes_per_step": yw_bytes * 2,
                    "achieved_gbs_chain_write": (yw_bytes / (chain_ms * 1e-3) / 1e9) if chain_ms > 0 else None,
                    "peak_gbs": peaks.get("hbm_gbs"), "note": "path is FP64-compute bound, not HBM bound (SURVEY 8d)"},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: PowerlawRedNoiseGP Woodbury lnpost, 10k TOAs, rank 60, 8192 walkers/GPU",
                       "ntoa": ntoa, "rank": p, "nfree": nd, "walkers_per_gpu": B, "parallelism": f"walker-shard x{world}",
                       "l2": "working set (y, w: 1.3 GB per step) exceeds L2; 4 distinct walker batches rotated"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (nd + (0 if psr.device_priors else 1)) * 8 * world),
                    "d2h_bytes_per_step": int(B * 8 * world), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": sampler.summary(),
            "finite_fraction": finite_frac,
            "flops_per_eval": flops_eval,
            "reference_equivalent_tflops_whole_step": flops_eval * B * world / (dev_ms / args.steps * 1e-3) / 1e12,
        }
        if world == 1:
            # the same workload with the structured Sigma path switched off: the dense packed-triangle contraction that
            # an arbitrary (non-harmonic) basis gets, rated on the reference algorithm's own flop count
            os.environ["VELA_B200_STRUCTURED"] = "0"
            try:
                dense = vb.Pulsar(ds.model, ds.toas, devices=[local_rank])
                dense.set_priors(priors)
            finally:
                del os.environ["VELA_B200_STRUCTURED"]
            nd_steps = max(3, min(args.steps, 20))
            for i in range(3):
                dense.lnpost_device(dev_sets[i % n_sets].data_ptr(), B, dev_lnpr[i % n_sets].data_ptr(), d_out.data_ptr(), stream)
            torch.cuda.synchronize()
            e0.record()
            for i in range(nd_steps):
                dense.lnpost_device(dev_sets[i % n_sets].data_ptr(), B, dev_lnpr[i % n_sets].data_ptr(), d_out.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            dense_ms = e0.elapsed_time(e1) / nd_steps
            dense.set_stage_timing(True)
            dstage = np.zeros(5)
            for i in range(3):
                dense.lnpost_device(dev_sets[i % n_sets].data_ptr(), B, dev_lnpr[i % n_sets].data_ptr(), d_out.data_ptr(), stream)
                torch.cuda.synchronize()
                dstage += np.array(dense.last_stage_ms())
            dstage /= 3
            dtf = flops_sigma * B / (dstage[2] * 1e-3) / 1e12 if dstage[2] > 0 else None
            line["dense_sigma_path"] = {
                "value": B / (dense_ms * 1e-3), "unit": UNIT, "ms_per_step": dense_ms, "steps": nd_steps,
                "kernel": "sigma_contract_kernel (FP64 DMMA m8n8k4, packed lower triangle)", "kernel_ms": float(dstage[2]),
                "achieved": dtf, "peak": dmma_peak, "unit_roofline": "TFLOP/s", "frac": (dtf / dmma_peak) if dtf else None,
                "note": "VELA_B200_STRUCTURED=0; what a basis without harmonic structure runs"}
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as O
            md = vb.ModelDescriptor(ds.model, ds.toas)
            cores = host_threads()
            probe = host_sets[0][: 4 * cores]
            t0 = time.perf_counter()
            O.lnlike_batch(md, probe, nthreads=cores)
            rate = len(probe) / (time.perf_counter() - t0)
            sample = int(max(cores, min(B, rate * 15.0) // cores * cores))
            t0 = time.perf_counter()
            ref = O.lnlike_batch(md, host_sets[0][:sample], nthreads=cores) + host_lnpr[0][:sample]
            cpu_dt = time.perf_counter() - t0
            got = psr.lnpost_vectorized(host_sets[0][:sample])
            ok = np.isfinite(ref)
            rel = float(np.max(np.abs(got[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else None
            line["cpu_baseline"] = {"value": sample / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} walkers of one {WORKLOAD} batch, {cpu_dt:.1f} s "
                                              "(C/OpenMP restatement of calc_lnpost_vectorized; not Julia)"}
            line["parity"] = {"max_rel_lnpost_diff_vs_oracle": rel, "n": int(ok.sum()), "tolerance": 1e-9}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
