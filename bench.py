#!/usr/bin/env python
"""bench.py — lnpost evaluations/sec of the B200 hot path on the 10k-TOA GP pulsar (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference algorithm on host cores

One "step" = one pass of the hot path over one batch of 8192 synthetic walkers per GPU
(weak scaling: every rank evaluates its own 8192-walker shard of the model replica it holds, then
the per-walker lnpost vector is all-gathered over NCCL).  Workload = BASELINE.json configs[2]
("sim3_gp: PowerlawRedNoiseGP Woodbury likelihood, 10k TOAs, 30 Fourier harmonics, 8192 walkers"),
the configuration the headline metric is quoted on; data are synthetic (vela.jl_b200/synth.py).

Keys beyond the base contract (same JSON line):
  roofline          the kernel with the largest share of the step against the FP64 peak of its pipe measured in this
                    run; every kernel of the step under `roofline.kernels`; ncu-derived fields are read from
                    profiles/r2_ncu_config3.json (committed capture of this very step)
  cpu_baseline      the CPU oracle (C/OpenMP restatement of the reference algorithm, NOT Julia; built -O3 -march=native
                    on this host) on a bounded sample of the same workload;  parity: GPU vs oracle on that sample
  e2e               host buffers through the public `Pulsar.lnpost_vectorized` call, H2D/D2H inside the timed region
  strong            (N > 1) config 3 with 8192 walkers IN TOTAL split over the N ranks
  inprocess         (N > 1) ONE handle over all N devices inside rank 0's process (what the Julia host uses), end to end
  other_configs     BASELINE.json configs[1], [3] (N = 1) and configs[4] (100k TOAs, rank 300, 8192 walkers per rank = 65 536
                    over 8 GPUs) at full size, each with stage times and an oracle spot check
  small_batch       (N = 1) wall-clock per host call at emcee-sized batches
  dense_sigma_path  (N = 1) the same workload with the structured Sigma path off (dense FP64 tensor-core contraction)
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
WALKERS = 8192


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def headline_config(world: int, ntoa: int, rank: int, nfree: int, walkers: int = WALKERS) -> dict:
    """The `config` object of the JSON line; both arms print exactly this."""
    return {"workload": f"{WORKLOAD}: PowerlawRedNoiseGP Woodbury lnpost, 10k TOAs, rank 60, 8192 walkers/GPU",
            "ntoa": int(ntoa), "rank": int(rank), "nfree": int(nfree), "walkers_per_gpu": int(walkers),
            "parallelism": f"walker-shard x{world}",
            "l2": "working set (y, w: 1.3 GB per step) exceeds L2; 4 distinct walker batches rotated"}


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


def algorithmic_flops(ntoa: int, nrows: int, p: int, chain_per_toa: float = 330.0):
    """per-evaluation FLOPs; returns (total, sigma_contraction_only)."""
    chain = chain_per_toa * ntoa
    if p == 0:
        return chain, 0.0
    sigma = nrows * p * (p + 1) + 3 * nrows * p
    rest = 3 * nrows + p**3 / 3 + p * p + 10 * p
    return chain + sigma + rest, sigma


def load_oracle():
    """The CPU oracle, -O3 -march=native build when this host can compile it; returns (module, flags)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    native = O.use_native()
    return O, ("-O3 -march=native -ffp-contract=off" if native else "-O2 -march=x86-64-v3 -ffp-contract=off")


def cpu_reference_arm(args, rank: int, world: int):
    """`--impl reference`: the reference's CPU algorithm (oracle port: the reference is Julia and cannot run
    here), all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    vb = load_package()
    O, flags = load_oracle()
    from helpers import oracle_eval_factory
    ds = vb.synth.make_dataset(WORKLOAD, oracle_eval_factory(vb, O))
    md = vb.ModelDescriptor(ds.model, ds.toas)
    cores = host_threads()
    probe = ds.walkers(4 * cores, seed=99)
    t0 = time.perf_counter()
    O.lnlike_batch(md, probe, nthreads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    budget_s = 60.0 / max(args.steps + args.warmup, 1)
    sample = int(max(cores, min(WALKERS, rate * min(budget_s, 15.0)) // cores * cores))
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
        "config": headline_config(world, len(ds.toas), ds.info["rank"], ds.info["nfree"]),
        "walkers_evaluated_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} walkers of the {WORKLOAD} batch per step (C/OpenMP restatement of "
                                   f"calc_lnpost_vectorized built {flags}; the Julia reference cannot run in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "finite": bool(np.all(np.isfinite(out))),
    }
    print(json.dumps(line), flush=True)


def make_priors(vb, ds):
    """Priors for lnpost: flat boxes of +-50 sigma_par around the truth for timing parameters, the walker
    distributions for noise parameters (every family has a device twin, so lnprior runs on the GPU)."""
    priors = []
    for k in range(ds.info["nfree"]):
        if k in ds.noise_draw:
            kind, a, b = ds.noise_draw[k]
            priors.append(vb.priors.SimplePrior(f"p{k}", vb.priors.LogUniform(a, b) if kind == "loguniform" else
                                                vb.priors.Uniform(a - 10 * b, a + 10 * b) if kind == "normal" else
                                                vb.priors.LogNormal(np.log(a), max(b, 1e-3))))
        else:
            w = 50 * max(ds.sigma_par[k], 1e-300)
            priors.append(vb.priors.SimplePrior(f"p{k}", vb.priors.Uniform(ds.truth[k] - w, ds.truth[k] + w)))
    return priors


def ncu_profile():
    """Per-kernel numbers of the committed `ncu --set full` capture of the headline step (profiles/r2_ncu_config3.json:
    dram bytes per launch, busy fraction of the bounding pipe).  Valid for the default workload only."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_config3.json")))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=WALKERS, help="walkers per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip strong / inprocess / other_configs / small_batch")
    ap.add_argument("--profile", action="store_true",
                    help="after warm-up run ONE device-resident step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`)")
    ap.add_argument("--profile-config", default=WORKLOAD, help="with --profile: which synthetic config to run")
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
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctl = dist.new_group(backend="gloo")     # control plane: parks ranks without putting a kernel on their GPU

    vb = load_package()
    L = vb.lib.lib()
    # a dedicated non-default stream: the library launches on the stream it is given, and torch's events
    # (recorded on the current stream) must see those launches
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        tt = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.tolist()

    n_sets = 4   # rotate distinct walker batches so no step sees the previous step's inputs

    class Case:
        """One synthetic config on this rank's GPU: dataset, handle, device-resident walker batches."""

        def __init__(self, name, B, seed_base=1000):
            t0 = time.perf_counter()
            self.name, self.B = name, B
            self.ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t, devices=[local_rank]))
            self.psr = vb.Pulsar(self.ds.model, self.ds.toas, devices=[local_rank])
            self.priors = make_priors(vb, self.ds)
            self.psr.set_priors(self.priors)          # uploads them to the device as well
            self.host_sets = [self.ds.walkers(B, seed=seed_base * (rank + 1) + i) for i in range(n_sets)]
            self.host_lnpr = [np.ascontiguousarray(vb.priors.lnprior(self.priors, w)) for w in self.host_sets]
            self.dev_sets = [torch.from_numpy(w).cuda() for w in self.host_sets]
            self.dev_lnpr = [torch.from_numpy(lp).cuda() for lp in self.host_lnpr]
            self.d_out = torch.empty(B, dtype=torch.float64, device="cuda")
            self.gathered = torch.empty(B * world, dtype=torch.float64, device="cuda") if world > 1 else None
            self.ntoa, self.p, self.nd = len(self.ds.toas), self.ds.info["rank"], self.ds.info["nfree"]
            self.nrows = self.psr.n_rows
            log(f"[rank {rank}] {name}: {self.ds.info} ready in {time.perf_counter() - t0:.1f}s")

        def step_device(self, i, n=None, gather=True):
            n = n or self.B
            self.psr.lnpost_device(self.dev_sets[i % n_sets].data_ptr(), n, self.dev_lnpr[i % n_sets].data_ptr(),
                                   self.d_out.data_ptr(), stream)
            if world > 1 and gather:
                dist.all_gather_into_tensor(self.gathered[: n * world], self.d_out[:n])

        def time_device(self, steps, warmup, n=None):
            """K device-resident steps bracketed by barrier + synchronize; returns total ms (this rank)."""
            for i in range(warmup):
                self.step_device(i, n)
            barrier()
            e0.record()
            for i in range(steps):
                self.step_device(i, n)
            e1.record()
            barrier()
            return e0.elapsed_time(e1)

        def stage_ms(self, steps, n=None):
            self.psr.set_stage_timing(True)
            acc = np.zeros(5)
            for i in range(steps):
                self.step_device(i, n, gather=False)
                torch.cuda.synchronize()
                acc += np.array(self.psr.last_stage_ms())
            self.psr.set_stage_timing(False)
            return acc / steps

        def parity(self, O, n, threads):
            """max relative |lnpost_gpu - lnpost_oracle| on the first n walkers of batch 0 (tolerance 1e-9)."""
            md = vb.ModelDescriptor(self.ds.model, self.ds.toas)
            W = self.host_sets[0][:n]
            t0 = time.perf_counter()
            ref = O.lnlike_batch(md, W, nthreads=threads) + self.host_lnpr[0][:n]
            dt = time.perf_counter() - t0
            got = self.psr.lnpost_vectorized(W)
            ok = np.isfinite(ref)
            rel = float(np.max(np.abs(got[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else float("nan")
            return rel, int(ok.sum()), dt

    # =========================================================================================== headline: config 3
    c3 = Case(args.profile_config if args.profile else WORKLOAD, args.walkers)
    psr, B, ntoa, p, nd, nrows = c3.psr, c3.B, c3.ntoa, c3.p, c3.nd, c3.nrows

    if args.profile:
        for i in range(args.warmup):
            c3.step_device(i)
        barrier()
        torch.cuda.cudart().cudaProfilerStart()
        c3.step_device(0)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        log("profiled one step")
        return

    # ---- device-resident timing (`value`)
    for i in range(args.warmup):
        c3.step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = psr.launch_count()
    barrier()
    e0.record()
    for i in range(args.steps):
        c3.step_device(i)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = psr.launch_count() - launches0
    # per-stage durations: re-run K steps with a sync after each so the per-launch events can be read
    stage_ms = c3.stage_ms(args.steps)
    finite_frac = float(torch.isfinite(c3.d_out).double().mean().item())

    # ---- end-to-end timing (`e2e`).  One GPU: the public host call, host numpy in, host numpy out (pinned staging + H2D +
    #      device priors + kernels + D2H inside the call).  Several ranks: every step copies this rank's walkers from PINNED
    #      host memory to the device, runs the public device entry point (priors evaluated on the device), gathers the
    #      per-rank results over NCCL from the device buffer and reads the gathered vector back to pinned host memory -
    #      the result never bounces through the host before the collective.
    pin_in = [torch.from_numpy(w).pin_memory() for w in c3.host_sets] if world > 1 else None
    d_in = torch.empty((B, nd), dtype=torch.float64, device="cuda") if world > 1 else None
    pin_all = torch.empty(B * world, dtype=torch.float64).pin_memory() if world > 1 else None

    def step_e2e(case, i, n=None):
        n = n or case.B
        if world == 1:
            return case.psr.lnpost_vectorized(case.host_sets[i % n_sets][:n])
        d_in[:n].copy_(pin_in[i % n_sets][:n], non_blocking=True)
        case.psr.lnpost_device(d_in.data_ptr(), n, 0, case.d_out.data_ptr(), stream)
        dist.all_gather_into_tensor(case.gathered[: n * world], case.d_out[:n])
        pin_all[: n * world].copy_(case.gathered[: n * world], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pin_all

    def time_e2e(case, steps, warmup, n=None):
        for i in range(warmup):
            step_e2e(case, i, n)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            step_e2e(case, i, n)
        e1.record()
        barrier()
        return max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))

    e2e_ms = time_e2e(c3, args.steps, args.warmup)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    dev_ms, e2e_ms = max_over_ranks([dev_ms, e2e_ms])

    extras = {}
    O = flags = None
    if not args.no_extras or not args.no_cpu_baseline:
        O, flags = load_oracle()
    xsteps = max(3, min(args.steps, 20))

    # ============================================================================ strong scaling: 8192 walkers in total
    if world > 1 and not args.no_extras:
        n_s = (WALKERS + world - 1) // world
        s_dev = c3.time_device(args.steps, args.warmup, n_s)
        s_e2e = time_e2e(c3, args.steps, args.warmup, n_s)
        s_dev, s_e2e = max_over_ranks([s_dev, s_e2e])
        t1 = dev_ms / args.steps              # one rank, 8192 walkers: every rank of the weak-scaling run does exactly that
        extras["strong"] = {
            "workload": f"{WORKLOAD}: {WALKERS} walkers in total, {n_s} per rank", "walkers_total": n_s * world,
            "value": n_s * world * args.steps / (s_dev * 1e-3), "unit": UNIT, "ms_per_step": s_dev / args.steps,
            "e2e_value": n_s * world * args.steps / (s_e2e * 1e-3), "e2e_ms_per_step": s_e2e / args.steps,
            "single_gpu_ms_per_step_same_run": t1,
            "efficiency_vs_n1": (t1 / world) / (s_dev / args.steps) * (n_s * world / WALKERS),
            "note": "device-timed like `value` (all-gather inside); the N = 1 time is the weak-scaling step of this run, "
                    "where every rank evaluates 8192 walkers"}

    # ============================================================ in-process handle over all N devices (rank 0's process)
    if world > 1 and not args.no_extras:
        ndev = min(world, int(L.vela_b200_device_count()))
        res = None
        dist.barrier(group=ctl)
        if rank == 0:
            try:
                multi = vb.Pulsar(c3.ds.model, c3.ds.toas, devices=list(range(ndev)))
                multi.set_priors(c3.priors)
                Wm = [c3.ds.walkers(B * ndev, seed=4242 + i) for i in range(2)]
                for i in range(3):
                    out_m = multi.lnpost_vectorized(Wm[i % 2])
                t0 = time.perf_counter()
                for i in range(xsteps):
                    out_m = multi.lnpost_vectorized(Wm[i % 2])
                dt = time.perf_counter() - t0
                one = psr.lnpost_vectorized(Wm[(xsteps - 1) % 2][:256])
                res = {"devices": ndev, "walkers_per_call": B * ndev, "value": B * ndev * xsteps / dt, "unit": UNIT,
                       "ms_per_call": 1e3 * dt / xsteps, "steps": xsteps,
                       "max_rel_diff_vs_single_device_handle": float(np.max(np.abs(out_m[:256] - one) / np.abs(one))),
                       "finite_fraction": float(np.isfinite(out_m).mean()),
                       "note": "one vela_b200 handle with devices=[0..N-1] in ONE process: host matrix in, host vector out, "
                               "no collective; the other ranks are parked on a CPU (gloo) barrier meanwhile"}
                del multi
            except Exception as ex:   # noqa: BLE001
                res = {"error": str(ex)[:300]}
        dist.barrier(group=ctl)
        if rank == 0:
            extras["inprocess"] = res

    # =================================================================================== other BASELINE.json configs
    def run_other(name, Bc, n_par, steps):
        case = Case(name, Bc, seed_base=2000)
        for i in range(3):
            case.step_device(i, gather=False)
        torch.cuda.synchronize()
        ms = case.time_device(steps, 3)
        st = case.stage_ms(min(steps, 5))
        fin = float(torch.isfinite(case.d_out).double().mean().item())
        rel, n_ok, cpu_dt = case.parity(O, n_par, max(1, host_threads() // world))
        ms_max, rel_max = max_over_ranks([ms, rel if rel == rel else 1e300])
        fin_min = -max_over_ranks([-fin])[0]
        flops_eval, flops_sigma = algorithmic_flops(case.ntoa, case.nrows, case.p)
        dfma = float(L.vela_b200_fp64_peak_tflops(local_rank, 0))
        info = {"workload": name, "ntoa": case.ntoa, "rank": case.p, "nfree": case.nd, "walkers_per_gpu": Bc,
                "walkers_total": Bc * world, "value": Bc * world * steps / (ms_max * 1e-3), "unit": UNIT,
                "ms_per_step": ms_max / steps, "steps": steps, "finite_fraction": fin_min,
                "stage_ms": {"prologue": float(st[0]), "chain": float(st[1]), "sigma_contract": float(st[2]),
                             "chol_finish": float(st[3]), "all_kernels": float(st[4])},
                "structured_sigma": bool(case.psr.sigma_structured()) if case.p else None,
                "chain_specialized": bool(case.psr.chain_specialized()),
                "parity": {"max_rel_lnpost_diff_vs_oracle": rel_max, "walkers_checked_per_rank": n_ok, "tolerance": 1e-9,
                           "oracle_seconds": cpu_dt},
                "flops_per_eval": flops_eval,
                "reference_equivalent_tflops_per_gpu": flops_eval * Bc / (ms_max / steps * 1e-3) / 1e12,
                "fp64_fma_peak_tflops": dfma}
        if case.p and case.psr.sigma_structured():
            si = case.psr.sigma_info(Bc)
            info["sigma_plan"] = dict(si, packed_pairs_of_dense_path=case.p * (case.p + 1) // 2)
            if st[2] > 0:
                info["sigma_executed_tflops"] = si["executed_flops_per_walker"] * Bc / (st[2] * 1e-3) / 1e12
        del case
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        return info

    if not args.no_extras:
        others = {}
        plan = [("config5_array", WALKERS, 8, max(3, min(args.steps, 5)))]
        if world == 1:
            plan = [("config2_ell1", 4096, 64, xsteps), ("config4_dd_wb", WALKERS, 16, max(3, min(args.steps, 10)))] + plan
        for name, Bc, n_par, steps in plan:
            if world > 1:     # every rank takes part in the collectives inside: a failure must stop the job, not one rank
                others[name] = run_other(name, Bc, n_par, steps)
                continue
            try:
                others[name] = run_other(name, Bc, n_par, steps)
            except Exception as ex:   # noqa: BLE001
                others[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
                log(f"[rank {rank}] {name} failed: {ex}")
        extras["other_configs"] = others

    # ================================================================= small batches: wall clock per host call (N = 1)
    if world == 1 and not args.no_extras:
        sb = {}
        for bs in (1, 16, 64, 256, 1024):
            Wb = c3.host_sets[0][:bs]
            for _ in range(20):
                psr.lnpost_vectorized(Wb)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                psr.lnpost_vectorized(Wb)
                ts.append(time.perf_counter() - t0)
            sb[str(bs)] = {"ms_per_call_median": 1e3 * float(np.median(ts)), "ms_per_call_min": 1e3 * float(np.min(ts))}
        extras["small_batch"] = {"workload": WORKLOAD, "calls": 200, "what": "wall clock of Pulsar.lnpost_vectorized "
                                 "(host in, host out, device priors)", "by_batch": sb}

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
        # executed flops of the Sigma stage, from the handle itself (vela_b200_sigma_info): the direct column sums
        # (2 n_cols rows_pad) or, for large batches, the two-stage evaluation (stage-1 Chebyshev columns + stage-2 matrices)
        sig_info = psr.sigma_info(B) if structured else None
        sigma_exec = sig_info["executed_flops_per_walker"] if structured else float(flops_sigma)
        tf = lambda fl, ms: (fl * B / (ms * 1e-3) / 1e12) if ms > 0 else None   # noqa: E731
        # compute-bound kernels of this path run on the FP64 pipe; on B200 scalar DFMA and the DMMA tensor path share it
        # (measured peaks 36.5 vs 37.1 TFLOP/s), so "tensor" with the matching measured peak is the applicable roofline
        fp64_note = "FP64 pipe, scalar DFMA issue (shares the datapath of the FP64 tensor path on B200); peak = measured DFMA loop"
        chain_name = "toa_chain_kernel (per-model NVRTC build)" if psr.chain_specialized() else "toa_chain_kernel (generic)"
        sigma_name = "colsum_kernel (structured Sigma, FP64 DMMA m8n8k4)" if structured else "sigma_contract_kernel (FP64 DMMA m8n8k4)"
        chol_name = ("chol_warp_kernel" if p + 1 <= 64 and os.environ.get("VELA_B200_CHOL_WARP", "0") == "1" else
                     "chol_dmma_kernel" if structured and p <= 64 and os.environ.get("VELA_B200_CHOL_DMMA", "1") != "0" else
                     "chol_finish_kernel")
        kernels = [
            {"kernel": chain_name, "ncu_name": "vela_chain_spec",
             "bound": "tensor", "bound_detail": fp64_note, "ms": chain_ms, "algorithmic_flops_per_launch": 330.0 * ntoa * B,
             "achieved": tf(330.0 * ntoa, chain_ms), "peak": dfma_peak, "unit": "TFLOP/s"},
            {"kernel": sigma_name, "ncu_name": "colsum_kernel" if structured else "sigma_contract_kernel",
             "bound": "tensor", "ms": sigma_ms, "algorithmic_flops_per_launch": sigma_exec * B,
             "achieved": tf(sigma_exec, sigma_ms), "peak": dmma_peak, "unit": "TFLOP/s",
             "reference_equivalent_tflops": tf(float(flops_sigma), sigma_ms), "plan": sig_info},
            {"kernel": chol_name, "ncu_name": chol_name, "bound": "tensor", "bound_detail": fp64_note, "ms": chol_ms,
             "algorithmic_flops_per_launch": (p**3 / 3 + p * p + 10 * p) * B,
             "achieved": tf(p**3 / 3 + p * p + 10 * p, chol_ms), "peak": dfma_peak, "unit": "TFLOP/s"},
        ]
        # dram__bytes_read.sum + dram__bytes_write.sum per launch and the busy fraction of the bounding pipe, from the
        # committed `ncu --set full` capture of this very step; only valid for the default workload and kernel variants
        prof = ncu_profile() if (ntoa, p, B) == (10000, 60, WALKERS) else {}
        for k in kernels:
            pk = prof.get("kernels", {}).get(k["ncu_name"], {})
            k["traffic"] = pk.get("dram_bytes")
            k["pipe_active_ncu"] = pk.get("pipe_active")
            k["frac"] = (k["achieved"] / k["peak"]) if k["achieved"] and k["peak"] > 0 else None
            k["share_of_step"] = k["ms"] / float(stage_ms[4]) if stage_ms[4] > 0 else None
        dom = max(kernels, key=lambda k: k["ms"])
        roofline = {
            "bound": dom["bound"], "bound_detail": dom.get("bound_detail", "FP64 tensor cores (mma.sync m8n8k4 f64 -> DMMA)"),
            "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
            "unit": "TFLOP/s", "frac": dom["frac"], "traffic": dom["traffic"],
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                              f"({prof.get('source', 'no committed capture for this shape')})",
            "peak_source": "measured in this run: register-resident FP64 loops (vela_b200_fp64_peak_tflops: DFMA for the "
                           "scalar kernels, mma.sync.m8n8k4.f64 for the DMMA kernels); MEASURED_PEAKS.json holds only "
                           "HBM and bf16 peaks",
            "fp64_fma_peak_tflops": dfma_peak, "fp64_dmma_peak_tflops": dmma_peak,
            "algorithmic_flops_per_launch": dom["algorithmic_flops_per_launch"],
            "frac_note": "achieved = SURVEY 8(d)'s algorithmic count (330 FLOP-equivalents per TOA and walker for the chain) / "
                         "measured kernel time; the chain kernel executes fewer FP64 operations than that count assumes "
                         "(~102 FP64 instructions per pair: logarithms, square root, divisions and the double-double phase "
                         "are replaced by expansions around the default parameters), so `frac` can approach 1 while the "
                         "pipe itself is `pipe_active_ncu` busy (committed ncu capture) - the kernel is latency-bound",
            "pipe_active_ncu": dom.get("pipe_active_ncu"),
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
            "config": headline_config(world, ntoa, p, nd, B),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (nd + (0 if psr.device_priors else 1)) * 8 * world),
                    # one rank reads its B results; with several ranks every rank reads the gathered B x world vector
                    "d2h_bytes_per_step": int(B * 8 * (world * world if world > 1 else 1)), "ms_per_step": e2e_ms / args.steps,
                    "path": "Pulsar.lnpost_vectorized (host numpy in / out)" if world == 1 else
                            "pinned H2D -> Pulsar.lnpost_device (device priors) -> NCCL all_gather -> pinned D2H of the gathered vector"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": sampler.summary(),
            "finite_fraction": finite_frac,
            "flops_per_eval": flops_eval,
            "reference_equivalent_tflops_whole_step": flops_eval * B * world / (dev_ms / args.steps * 1e-3) / 1e12,
            "reference_equivalent_note": "the structured Sigma path replaces the reference's p(p+1)/2 pair contractions by "
                                         "trigonometric column sums (an identity on harmonic bases), so the step does the "
                                         "reference's work with fewer executed flops: this rate may exceed the FP64 peak; the "
                                         "per-kernel `achieved` figures count executed flops",
        }
        line.update(extras)
        if world == 1 and not args.no_extras:
            # the same workload with the structured Sigma path switched off: the dense packed-triangle contraction that
            # an arbitrary (non-harmonic) basis gets, rated on the reference algorithm's own flop count
            os.environ["VELA_B200_STRUCTURED"] = "0"
            try:
                dense = vb.Pulsar(c3.ds.model, c3.ds.toas, devices=[local_rank])
                dense.set_priors(c3.priors)
            finally:
                del os.environ["VELA_B200_STRUCTURED"]
            nd_steps = max(3, min(args.steps, 20))

            def dstep(i):
                dense.lnpost_device(c3.dev_sets[i % n_sets].data_ptr(), B, c3.dev_lnpr[i % n_sets].data_ptr(), c3.d_out.data_ptr(), stream)

            for i in range(3):
                dstep(i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(nd_steps):
                dstep(i)
            e1.record()
            torch.cuda.synchronize()
            dense_ms = e0.elapsed_time(e1) / nd_steps
            dense.set_stage_timing(True)
            dstage = np.zeros(5)
            for i in range(3):
                dstep(i)
                torch.cuda.synchronize()
                dstage += np.array(dense.last_stage_ms())
            dstage /= 3
            dtf = flops_sigma * B / (dstage[2] * 1e-3) / 1e12 if dstage[2] > 0 else None
            line["dense_sigma_path"] = {
                "value": B / (dense_ms * 1e-3), "unit": UNIT, "ms_per_step": dense_ms, "steps": nd_steps,
                "kernel": "sigma_contract_kernel (FP64 DMMA m8n8k4, packed lower triangle)", "kernel_ms": float(dstage[2]),
                "achieved": dtf, "peak": dmma_peak, "unit_roofline": "TFLOP/s", "frac": (dtf / dmma_peak) if dtf else None,
                "note": "VELA_B200_STRUCTURED=0; what a basis without harmonic structure runs"}
            del dense
        if not args.no_cpu_baseline and world == 1:
            md = vb.ModelDescriptor(c3.ds.model, c3.ds.toas)
            cores = host_threads()
            probe = c3.host_sets[0][: 4 * cores]
            t0 = time.perf_counter()
            O.lnlike_batch(md, probe, nthreads=cores)
            rate = len(probe) / (time.perf_counter() - t0)
            sample = int(max(cores, min(B, rate * 15.0) // cores * cores))
            t0 = time.perf_counter()
            ref = O.lnlike_batch(md, c3.host_sets[0][:sample], nthreads=cores) + c3.host_lnpr[0][:sample]
            cpu_dt = time.perf_counter() - t0
            got = psr.lnpost_vectorized(c3.host_sets[0][:sample])
            ok = np.isfinite(ref)
            rel = float(np.max(np.abs(got[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else None
            line["cpu_baseline"] = {"value": sample / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} walkers of one {WORKLOAD} batch, {cpu_dt:.1f} s "
                                              f"(C/OpenMP restatement of calc_lnpost_vectorized built {flags}; not Julia)"}
            line["parity"] = {"max_rel_lnpost_diff_vs_oracle": rel, "n": int(ok.sum()), "tolerance": 1e-9}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=ctl)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
