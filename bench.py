#!/usr/bin/env python
"""Benchmark of the RCHQ batch-selection hot path (BASELINE.json metric: recombination candidates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c1..c5]

One "step" = one full ``recombination(...)`` call (Nystrom block + all grouped passes + CAR + compaction) over one
batch of synthetic candidates.

Workload (``--workload auto``, what the driver runs):
  N = 1  BASELINE.json configs[1] (C2, Hartmann-6D shape: n_rec = 1e6, n_nys = 1000, batch = 200, Matern-5/2, f64); the
         line also carries a ``c5`` block (the north_star target config on this one GPU), a ``parity`` block (fast mode
         vs the CPU oracle on a 1e5-candidate sample of the workload: indices / weights / MMD), ``parity_mode`` (ms per
         step of the mode that reproduces the reference's op sequence) and ``reference_gpu`` (the reference algorithm's
         own PyTorch path -- oracle port -- executing on this GPU).
  N > 1  BASELINE.json configs[4] (C5, the north_star target: n_rec = 1e7 TOTAL, n_nys = 2000, batch = 1000), STRONG
         scaling: candidates row-sharded over the ranks, one all-reduce of the group sums per iteration; the line also
         carries ``c5_1gpu`` (the same config on rank 0 alone, measured in the same run) and ``weak_c2`` (C2 with 1e6
         candidates per GPU).
Rank 0 prints ONE JSON line.

``--impl reference`` times the reference algorithm's CPU path (the oracle restatement of SOBER/_rchq.py, which is
bit-identical to it; the reference itself is Python and /root/reference does not exist on the GPU box) with all host
threads: on the full workload when the host has the memory for the reference's materialised (E, L, S) Gram
(40 B x N x n_nys), else on a bounded sample of it (``extrapolated: true``).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_rec per GPU, d, n_nys, batch, family, lengthscale, description)
    "c1": (20_000, 2, 500, 100, "rbf", 1.0, "C1 Branin 2-D: n_rec=2e4, n_nys=500, batch=100, RBF"),
    "c2": (1_000_000, 6, 1000, 200, "matern", 0.5,
           "C2 Hartmann-6D: n_rec=1e6/GPU, n_nys=1000, batch=200, Matern-5/2, kernel mode"),
    "c3": (2_000_000, 24, 500, 100, "rbf", 2.0, "C3 Ising 24-D binary: n_rec=2e6, n_nys=500, batch=100, RBF(Hamming)"),
    "c4": (5_000_000, 1024, 1000, 500, "tanimoto", None,
           "C4 drug 1024-bit fingerprints: n_rec=5e6, n_nys=1000, batch=500, Tanimoto"),
    "c5": (10_000_000, 6, 2000, 1000, "matern", 0.5,
           "C5 Hartmann-6D scaling: n_rec=1e7 TOTAL (strong scaling), n_nys=2000, batch=1000, Matern-5/2"),
}
# algorithmic flops per kernel evaluation: 2d + 2 + c_k
C_K = {"rbf": 4, "matern": 12, "tanimoto": 6}


def synth(name, n, seed, device, generator_device=None):
    """Synthetic candidates of the workload's shape, landmarks = a random subset."""
    n_rec, d, L, b, fam, ls, _ = WORKLOADS[name]
    g = torch.Generator(device=device).manual_seed(seed)
    if name == "c3":
        X = (torch.rand(n, d, device=device, generator=g) < 0.5).to(torch.float64)
    elif name == "c4":
        X = (torch.rand(n, d, device=device, generator=g) < 0.05).to(torch.float64)
    elif name == "c1":
        X = torch.rand(n, d, dtype=torch.float64, device=device, generator=g) * 5 - 2
    else:
        X = torch.rand(n, d, dtype=torch.float64, device=device, generator=g)
    mu = torch.rand(n, dtype=torch.float64, device=device, generator=g)
    return X, mu


def make_kernel(name, device, pred_cov=False):
    from oracle import kernels as ok       # kernel OBJECTS only (gpytorch stand-ins); the product introspects them
    _, d, L, b, fam, ls, _ = WORKLOADS[name]
    cov = ok.make_kernel(fam, [ls] if ls is not None else 1.0, 1.0).to(device)
    if not pred_cov:
        return ok.Kernel(ok.BareModel(cov), mode="kernel")
    # the default Sober kernel: GP posterior predictive covariance with a synthetic GP of 200 observations
    x_obs = synth(name, 200, 5, torch.device("cpu"))[0].to(device)
    return ok.Kernel(ok.GPModel(cov, x_obs, None, noise=1e-4), mode="predictive_covariance")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_rate(name, sample_n, steps, warmup, threads):
    """The reference algorithm on the host cores: oracle/rchq.py (bit-identical to SOBER/_rchq.py on the CPU) on
    ``sample_n`` candidates of the workload.  Returns candidates/sec (best step) and the per-step seconds."""
    from oracle import rchq
    torch.set_num_threads(threads)
    _, d, L, b, fam, ls, _ = WORKLOADS[name]
    cpu = torch.device("cpu")
    X, mu = synth(name, sample_n, 0, cpu)
    mu /= mu.sum()
    Z = X[torch.randperm(sample_n, generator=torch.Generator().manual_seed(1))[:L]].clone()
    kern = make_kernel(name, cpu)
    times = []
    for it in range(warmup + steps):
        w0 = mu.clone()
        torch.manual_seed(7)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rchq.recombination(X, Z, b, kern, None, None, init_weights=w0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_n / min(times), times


def host_memory_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 0


def reference_arm(args, name, threads):
    """``--impl reference``: the reference algorithm's CPU path on this box's host cores, EXACTLY ``--steps`` timed and
    ``--warmup`` untimed steps.  Each step runs a bounded sample of the workload (full n_nys / batch / kernel; the
    candidate count is what is bounded) sized from a short probe so that the whole run ends within a few minutes
    (SOBER_B200_REF_BUDGET_S, default 240 s); the full workload is used when it fits that budget and the host's memory.
    The metric is a rate, and the reference's time is linear in the candidate count apart from the N-independent
    Nystrom / PSD-gate cost -- a smaller sample is therefore somewhat PESSIMISTIC for the CPU (measured in the build
    container, 8 threads, C2: 2.0e4 candidates/s on a 1.75e5 sample, 2.5e4 on the full 1e6)."""
    n_rec, d, L, b, fam, ls, desc = WORKLOADS[name]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("SOBER_B200_REF_BUDGET_S", "240"))
    # the reference materialises the (E, L, S) Gram and a few temporaries of its size: ~40 B x N x n_nys
    mem_cap = max(20_000, int((host_memory_available() - (16 << 30)) // (48 * L))) if host_memory_available() else 100_000
    if args.cpu_sample > 0:
        sample, probe = min(args.cpu_sample, n_rec), None
    else:
        n0 = min(n_rec, 20_000)
        n1 = min(n_rec, 60_000)
        _, t0 = cpu_reference_rate(name, n0, 1, 1, threads)          # the warm-up run also spins up the thread pool
        _, t1 = cpu_reference_rate(name, n1, 1, 0, threads)
        per_cand = max((t1[0] - t0[0]) / max(n1 - n0, 1), 1e-9) if n1 > n0 else t0[0] / n0
        fixed = max(t0[0] - per_cand * n0, 0.0)
        per_step = max(budget - (2 * t0[0] + t1[0]), 10.0) / (steps + warm)
        sample = int(max(per_step - fixed, 0.0) / per_cand)
        sample = max(min(sample, n_rec, mem_cap), min(n_rec, 20_000))
        if sample < n_rec:
            sample = max(1000, sample // 1000 * 1000)
        probe = {"n": [n0, n1], "seconds": [t0[0], t1[0]], "fixed_seconds": fixed, "seconds_per_candidate": per_cand}
    _, times = cpu_reference_rate(name, sample, steps, warm, threads)
    mean = sum(times) / len(times)
    rate = sample / mean
    what = ("oracle/rchq.py (bit-identical restatement of SOBER/_rchq.py) on %s of the workload's %d candidates per step, "
            "full n_nys/batch, torch CPU f64, %d threads, %d timed + %d warm-up steps; the reference materialises the "
            "(E,L,S) Gram so its memory grows as 40 B x N x n_nys" % ("ALL" if sample == n_rec else "%d" % sample, n_rec,
                                                                      threads, steps, warm))
    print(json.dumps({
        "impl": "reference", "metric": "recombination candidates/sec", "value": rate, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * mean,
        "higher_is_better": True, "scaling": "strong" if name == "c5" else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "n_nys": L, "batch": b, "kernel": fam, "sample_n_rec": sample,
                   "extrapolated": sample != n_rec,
                   "extrapolation": None if sample == n_rec else
                   "candidates/s measured on the sample; the reference's time is linear in N x n_nys (84 % kernel "
                   "evaluations, SURVEY.md section 6) plus an N-independent Nystrom / PSD-gate term, its memory too",
                   "sizing_probe": probe},
        "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank's device: data, kernel object, the timed loops."""

    def __init__(self, name, dev, rank, world, pred_cov=False, sharded=True, n_override=None):
        import sober_b200
        self.sb = sober_b200
        self.name, self.dev, self.rank, self.world = name, dev, rank, (world if sharded else 1)
        n_rec, self.d, self.L, self.b, self.fam, self.ls, self.desc = WORKLOADS[name]
        self.strong = name == "c5"
        self.n_local = n_override or (n_rec // self.world if self.strong else n_rec)
        self.n_total = self.n_local * self.world
        self.X, self.mu = synth(name, self.n_local, 100 + (rank if sharded else 0), dev)
        if self.world > 1:
            import torch.distributed as dist
            tot = self.mu.sum()
            dist.all_reduce(tot)
            self.mu /= tot
        else:
            self.mu /= self.mu.sum()
        # landmarks: replicated; drawn from rank 0's shard and broadcast
        gen = torch.Generator(device=dev).manual_seed(1)
        self.Z = self.X[torch.randperm(self.n_local, device=dev, generator=gen)[:self.L]].clone()
        if self.world > 1:
            dist.broadcast(self.Z, 0)
        self.kern = make_kernel(name, dev, pred_cov)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def step(self, weights, X=None):
        torch.manual_seed(7)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return self.sb.recombination(self.X if X is None else X, self.Z, self.b, self.kern, self.dev,
                                         torch.float64, init_weights=weights)

    def max_over_ranks(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    def timed(self, steps, warmup, flush, ops=None):
        """W untimed + K timed steps, inputs resident in HBM; CUDA events on the launching stream, barrier + synchronize
        on both sides, max over ranks.  Returns (ms total, idx, w)."""
        for _ in range(warmup):
            flush.zero_()
            idx, w = self.step(self.mu.clone())
        self.barrier()
        if ops is not None:
            ops.timing = {}
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t_start.record()
        for _ in range(steps):
            flush.zero_()
            idx, w = self.step(self.mu.clone())
        t_end.record()
        self.barrier()
        ms = self.max_over_ranks(t_start.elapsed_time(t_end))
        assert len(idx) <= self.b and abs(float(w.sum()) - 1.0) < 1e-9, "benchmark result failed its invariants"
        return ms, idx, w

    def timed_e2e(self, steps, flush):
        """The same metric through the public API with HOST buffers: pinned inputs, host->device copies of the
        candidates and weights and the device->host read of (idx, w) inside the timed region, every step."""
        Xh, muh = self.X.cpu().pin_memory(), self.mu.cpu().pin_memory()
        # recombination() turns its init_weights into the sparse solution IN PLACE, so every step gets its own pinned
        # copy of the input weights, made here, outside the timed region (a host-to-host copy is not part of the path)
        whs = [muh.clone().pin_memory() for _ in range(steps + 2)]
        for k in range(2):
            idx_e, w_e = self.step(whs[steps + k], Xh)
        self.barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for k in range(steps):
            flush.zero_()
            idx_e, w_e = self.step(whs[k], Xh)
            idx_host, w_host = idx_e.cpu(), w_e.cpu()
        e_ev.record()
        self.barrier()
        ms = self.max_over_ranks(s_ev.elapsed_time(e_ev))
        # the host-buffer path returns the same kind of answer: at most b points, weights summing to the input mass, and
        # the caller's host weight vector turned into that sparse solution in place (this rank's part of it)
        assert len(idx_host) <= self.b and abs(float(w_host.sum()) - 1.0) < 1e-9, "e2e result failed its invariants"
        assert int((whs[steps - 1] != 0).sum()) <= self.b, "e2e: init_weights was not turned into the sparse solution"
        return ms, Xh.numel() * 8 + muh.numel() * 8, idx_host.numel() * 8 + w_host.numel() * 8


def parity_block(runner, sample_n):
    """Fast mode (the benchmarked mode) against the CPU oracle on a sample of the workload: the oracle gets the same
    test matrix for the randomised range finder and the fast mode's null-space construction restated with LAPACK
    (oracle.projector_nullspace) -- SURVEY.md TL;DR 3 and 7: both are free choices of the reference's algorithm that
    decide which vertex it walks to.  Returns the block for the JSON line."""
    from oracle import rchq as oracle
    from sober_b200 import _nystrom
    name, dev = runner.name, runner.dev
    _, d, L, b, fam, ls, _ = WORKLOADS[name]
    cpu = torch.device("cpu")
    X, mu = synth(name, sample_n, 0, cpu)
    mu /= mu.sum()
    Z = X[torch.randperm(sample_n, generator=torch.Generator().manual_seed(1))[:L]].clone()
    R = torch.randn(L, b - 1, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    kern_cpu = make_kernel(name, cpu)
    orig = torch.randn
    torch.randn = lambda *a, **k: R.clone() if tuple(a[:2]) == tuple(R.shape) else orig(*a, **k)
    t0 = time.perf_counter()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m_o = mu.clone()
            idx_o, w_o = oracle.recombination(X, Z, b, kern_cpu, None, None, init_weights=m_o,
                                              nullspace=oracle.projector_nullspace)
    finally:
        torch.randn = orig
    t_oracle = time.perf_counter() - t0
    _nystrom._injected_test_matrix = R
    try:
        with warnings.catch_warnings(), runner.sb.configure(mode="fast"):
            warnings.simplefilter("ignore")
            m_g = mu.clone().to(dev)
            idx_g, w_g = runner.sb.recombination(X.to(dev), Z.to(dev), b, runner.kern, dev, torch.float64,
                                                 init_weights=m_g)
    finally:
        _nystrom._injected_test_matrix = None
    torch.cuda.synchronize(dev)
    same = bool(torch.equal(idx_g.cpu(), idx_o))
    common = sorted(set(idx_g.cpu().tolist()) & set(idx_o.tolist()))
    out = {"mode": "fast", "n": sample_n, "n_nys": L, "batch": b, "indices_identical": same,
           "points": int(len(idx_o)), "points_in_common": len(common),
           "oracle": "oracle/rchq.py on the CPU, same test matrix, nullspace=oracle.projector_nullspace",
           "oracle_seconds": t_oracle}
    if same:
        out["max_dw"] = float((w_g.cpu() - w_o).abs().max())
        out["max_dmu_in_place"] = float((m_g.cpu() - m_o).abs().max())
    # worst-case quadrature error of both batches, on the device with the kernel callable
    Xd, mud = X.to(dev), mu.to(dev)
    mmd_o = float(oracle.mmd_squared(runner.kern, Xd, mud, idx_o.to(dev), w_o.to(dev), chunk=8192))
    mmd_g = float(oracle.mmd_squared(runner.kern, Xd, mud, idx_g, w_g.to(torch.float64), chunk=8192))
    out["mmd2_oracle"], out["mmd2_ours"] = mmd_o, mmd_g
    out["mmd_rel"] = abs(mmd_g - mmd_o) / abs(mmd_o) if mmd_o != 0 else None
    out["ok"] = bool(same and out["max_dw"] < 1e-6 and out["mmd_rel"] is not None and out["mmd_rel"] < 1e-6)
    return out


def sharded_parity(dev, rank, world):
    """N > 1: the row-sharded run (NCCL all-reduce of the group sums) against the single-GPU run of the SAME global
    input on rank 0 -- a small C2-shaped fixture, before anything is timed.  Summation order differs between the two
    (per-rank partial sums), so weights agree to rounding and indices wherever no near-tie exists."""
    import torch.distributed as dist
    import sober_b200
    n_per, L, b = 20_000, 300, 64
    X, mu = synth("c2", n_per, 300 + rank, dev)
    tot = mu.sum()
    dist.all_reduce(tot)
    mu /= tot
    Z = X[:L].clone()
    dist.broadcast(Z, 0)
    kern = make_kernel("c2", dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        idx_s, w_s = sober_b200.recombination(X, Z, b, kern, dev, torch.float64, init_weights=mu.clone())
    xs = [torch.empty_like(X) for _ in range(world)]
    ms = [torch.empty_like(mu) for _ in range(world)]
    dist.all_gather(xs, X)
    dist.all_gather(ms, mu)
    out = None
    if rank == 0:
        sober_b200.set_communicator(None)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.manual_seed(7)
            idx_1, w_1 = sober_b200.recombination(torch.cat(xs), Z, b, kern, dev, torch.float64,
                                                  init_weights=torch.cat(ms))
        sober_b200.enable_sharding()
        same = bool(torch.equal(idx_s, idx_1))
        out = {"n_total": n_per * world, "n_nys": L, "batch": b, "indices_identical": same,
               "max_dw": float((w_s - w_1).abs().max()) if same else None,
               "what": "row-sharded run over %d ranks vs the same global input on rank 0 alone" % world}
    dist.barrier()
    return out


def reference_gpu_leg(runner, flush):
    """The reference algorithm's own PyTorch path (oracle port, op for op SOBER/_rchq.py) executing on THIS GPU on the
    full workload (the same-box comparison).  One warm-up on a small sample, one timed call."""
    from oracle import rchq as oracle
    dev = runner.dev
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.manual_seed(7)
            n_small = min(20_000, runner.n_local)
            oracle.recombination(runner.X[:n_small], runner.Z, runner.b, runner.kern, dev, None,
                                 init_weights=runner.mu[:n_small].clone())
            torch.cuda.synchronize(dev)
            flush.zero_()
            torch.manual_seed(7)
            w0 = runner.mu.clone()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            idx, w = oracle.recombination(runner.X, runner.Z, runner.b, runner.kern, dev, None, init_weights=w0)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
        peak = torch.cuda.max_memory_allocated(dev)
        return {"value": runner.n_local / dt, "unit": "candidates/s", "ms_per_step": 1e3 * dt, "steps": 1,
                "kind": "oracle/rchq.py (port of SOBER/_rchq.py) with device=cuda: torch ops -> cuBLAS / cuSOLVER / ATen",
                "peak_device_bytes": int(peak), "points": int(len(idx))}
    except Exception as err:            # e.g. out of memory for the materialised (E, L, S) Gram
        torch.cuda.empty_cache()
        return {"unavailable": "%s: %s" % (type(err).__name__, str(err).splitlines()[0][:160] if str(err) else "")}


def fp64_peak(ops):
    """FP64 FMA peak of this GPU (not in MEASURED_PEAKS.json): dependent-chain DFMA probe, 2 flop per FMA."""
    iters, blocks = 1 << 16, 148 * 8
    for _ in range(2):
        ops.fp64_probe(blocks, iters)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    ops.fp64_probe(blocks, iters)
    p1.record()
    torch.cuda.synchronize()
    return blocks * 256 * iters * 16 / (p0.elapsed_time(p1) * 1e-3) / 1e12


def popc_peak(ops):
    """Integer-pipe peak in 64-bit AND+POPC word-ops per second (the unit of the bit-packed K1 kernels)."""
    iters, blocks = 1 << 16, 148 * 8
    for _ in range(2):
        ops.popc_probe(blocks, iters)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    ops.popc_probe(blocks, iters)
    p1.record()
    torch.cuda.synchronize()
    return blocks * 256 * iters * 4 / (p0.elapsed_time(p1) * 1e-3) / 1e9


def measured_traffic(name, world, pred_cov):
    """DRAM bytes of the largest K1 launch from the committed ncu --set full capture of this shape
    (profiles/k1_traffic.json, written by tools/ncu_pick.py from the .ncu-rep of the round), or None."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        rec = table.get("%s_%dgpu%s" % (name, world, "_predcov" if pred_cov else ""))
        return (rec["dram_bytes"], rec["source"]) if rec else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="candidates of the CPU reference runs; 0 = automatic (1e5 for cpu_baseline / parity; the "
                         "reference arm takes the full workload when the host memory allows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c5 / parity / parity_mode / reference_gpu blocks")
    ap.add_argument("--pred-cov", action="store_true",
                    help="use Kernel(model, 'predictive_covariance') with a synthetic 200-observation GP (SURVEY 8d)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload if args.workload != "auto" else ("c2" if world == 1 else "c5")
    n_rec, d, L, b, fam, ls, desc = WORKLOADS[name]
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, name, threads)
        return

    # --------------------------------------------------------------------------------------------------
    import sober_b200
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        sober_b200.enable_sharding()
    from sober_b200 import _rchq, _linalg
    ops = _rchq._ops()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    sober_b200.options.set_mode(args.mode)
    shard_check = sharded_parity(dev, rank, world) if (world > 1 and not args.no_extras) else None
    run = Runner(name, dev, rank, world, args.pred_cov)

    # ---- timed region: K steps, inputs resident in HBM ----
    launches0 = ops.launches + _linalg.launches
    for _ in range(args.warmup):
        flush.zero_()
        run.step(run.mu.clone())
    launches0 = ops.launches + _linalg.launches
    with ClockSampler(local_rank) as clocks:
        elapsed_ms, idx, w = run.timed(args.steps, 0, flush, ops)
    timing = ops.timing_summary()
    uc_big = ops.timing_largest("update_compact")
    k1_big = ops.timing_largest("group_accumulate")
    ops.timing = None
    launches = ops.launches + _linalg.launches - launches0

    # ---- e2e: host (pinned) buffers in, host results out, copies inside the timed region ----
    e_steps = max(2, args.steps // 2)
    e2e_ms, h2d, d2h = run.timed_e2e(e_steps, flush)

    # ---- the other configuration of the scaling story, measured in the same run ----
    extras = {}
    if not args.no_extras and args.workload == "auto" and args.mode == "fast":
        if world > 1:
            wk = Runner("c2", dev, rank, world)
            ksteps = max(5, args.steps)
            ms, _, _ = wk.timed(ksteps, 3, flush)
            extras["weak_c2"] = {"workload": WORKLOADS["c2"][6], "scaling": "weak", "n_rec_total": wk.n_total,
                                 "value": wk.n_total * ksteps / (ms * 1e-3), "unit": "candidates/s",
                                 "ms_per_step": ms / ksteps, "steps": ksteps}
            del wk
            # the strong-scaling reference point: the same C5 call on rank 0 alone
            if rank == 0:
                sober_b200.set_communicator(None)
                one = Runner("c5", dev, 0, 1, sharded=False)
                ksteps = max(2, min(3, args.steps))
                ms, _, _ = one.timed(ksteps, 2, flush)
                extras["c5_1gpu"] = {"n_rec_total": one.n_total, "ms_per_step": ms / ksteps, "steps": ksteps,
                                     "value": one.n_total * ksteps / (ms * 1e-3), "unit": "candidates/s",
                                     "note": "same config on rank 0 alone while the other ranks wait"}
                del one
                sober_b200.enable_sharding()
            dist.barrier()
        else:
            one = Runner("c5", dev, 0, 1)
            ksteps = max(2, min(5, args.steps // 2))
            ms, _, _ = one.timed(ksteps, 2, flush, ops)
            t5 = ops.timing_summary()
            uc5, k15 = ops.timing_largest("update_compact"), ops.timing_largest("group_accumulate")
            ops.timing = None
            extras["c5"] = {"workload": WORKLOADS["c5"][6], "n_rec_total": one.n_total, "ms_per_step": ms / ksteps,
                            "steps": ksteps, "value": one.n_total * ksteps / (ms * 1e-3), "unit": "candidates/s",
                            "note": "the north_star target config (BASELINE configs[4]) on this one GPU",
                            "stage_ms_per_step": {k: v[1] / ksteps for k, v in t5.items()},
                            # the streaming pass over all 1e7 candidates (weight update + compaction + record move)
                            "update_compact_largest": {"ms": uc5[0], "bytes": uc5[1],
                                                       "GB/s": uc5[1] / (uc5[0] * 1e-3) / 1e9} if uc5 else None,
                            "k1_largest": {"ms": k15[0], "pairs": k15[1],
                                           "tflops_at_26_flop_per_pair": k15[1] * 26 / (k15[0] * 1e-3) / 1e12} if k15 else None}
            del one
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1) and of the streaming pass (HBM bound) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    flop_per_pair = 2 * d + 2 + C_K[fam]
    k1_calls, k1_ms, k1_pairs = timing.get("group_accumulate", (0, 0.0, 0))
    uc_calls, uc_ms, uc_bytes = timing.get("update_compact", (0, 0.0, 0))
    # HBM roofline of the streaming pass: the pass over ALL candidates (first iteration); later iterations halve
    uc_gbs = uc_big[1] / (uc_big[0] * 1e-3) / 1e9 if uc_big and uc_big[0] > 0 else None
    bits = d > 8 and fam in ("tanimoto", "rbf")          # the bit-packed K1 kernels (Tanimoto popcount / Hamming table)
    words = (d + 63) // 64
    step_ms = elapsed_ms / args.steps
    share = k1_ms / elapsed_ms if elapsed_ms else None
    if not bits:
        peak = fp64_peak(ops)
        k1_tflops = k1_pairs * flop_per_pair / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else None
        traffic, traffic_src = measured_traffic(name, world, args.pred_cov)
        roofline = {
            "kernel": "group_accumulate (K1: fused cross-kernel + weighted group sums, record layout)",
            "bound": "fp64",          # FP64 FMA pipe; the hbm|tensor enum has no entry for it
            "achieved": k1_tflops, "peak": peak, "unit": "TFLOP/s",
            "frac": (k1_tflops / peak) if k1_tflops else None,
            "traffic": traffic, "traffic_source": traffic_src,
            "traffic_unit": "bytes per largest launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "peak_source": "measured in this run: dependent-chain DFMA probe (sober_fp64_probe), 2 flop per FMA",
            "algorithmic_flop_per_pair": flop_per_pair, "pairs_per_step": k1_pairs / max(args.steps, 1),
            "launches": k1_calls, "ms_per_step": k1_ms / max(args.steps, 1), "share_of_step": share,
            "largest_launch": {"ms": k1_big[0], "pairs": k1_big[1],
                               "tflops": k1_big[1] * flop_per_pair / (k1_big[0] * 1e-3) / 1e12} if k1_big else None,
        }
    elif fam == "tanimoto" and d % 256 == 0 and d <= 1024:
        # bit-packed Tanimoto on the tensor cores (csrc/group_bits_mma.cu): <x, z> as an int8 GEMM over 0/1 bytes
        bf16 = peaks.get("bf16_tflops")
        peak = 2.0 * bf16 if bf16 else 4500.0
        rate = 2.0 * k1_pairs * d / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else None
        roofline = {
            "kernel": "group_accumulate (K1, bit-packed Tanimoto on tcgen05: kind::i8 MMA, landmark tile resident in TMEM as the A "
                      "operand, candidate fingerprints expanded to 0/1 bytes in shared memory, int32 accumulators in TMEM, FP64 "
                      "ratio epilogue)",
            "bound": "tensor", "achieved": rate, "peak": peak, "unit": "TOP/s (int8, 2 ops per bit pair)",
            "frac": (rate / peak) if rate else None, "traffic": None,
            "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops (int8 issues at twice the bf16 rate; no measured int8 entry)"
                            if bf16 else "nominal 4.5 POP/s dense int8"),
            "note": "the kernel is bound by instruction issue of the bit expansion and the FP64 epilogue, not by the MMA "
                    "(tensor pipe 31 % active): see profiles/r02_bits_tcgen05.txt; the popcount kernel it replaces ran at 122 G pairs/s",
            "pairs_per_step": k1_pairs / max(args.steps, 1), "pairs_per_second": k1_pairs / (k1_ms * 1e-3) if k1_ms > 0 else None,
            "launches": k1_calls, "ms_per_step": k1_ms / max(args.steps, 1), "share_of_step": share,
        }
    else:
        peak = popc_peak(ops)
        rate = k1_pairs * words / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None
        roofline = {
            "kernel": "group_accumulate (K1, bit-packed rows: popcount(x & z) + FP64 Tanimoto ratio, or popcount(x ^ z) "
                      "+ kernel-value table for a stationary kernel on {0,1}^d)",
            "bound": "int",           # integer pipe (AND/XOR + POPC)
            "achieved": rate, "peak": peak, "unit": "G word-op/s (64-bit AND+POPC)",
            "frac": (rate / peak) if rate else None, "traffic": None,
            "peak_source": "measured in this run: sober_popc_probe (independent popcount(x & z) chains, 64-bit words)",
            "algorithmic_word_ops_per_pair": words, "pairs_per_step": k1_pairs / max(args.steps, 1),
            "launches": k1_calls, "ms_per_step": k1_ms / max(args.steps, 1), "share_of_step": share,
        }

    out = {
        "metric": "recombination candidates/sec", "value": run.n_total * args.steps / (elapsed_ms * 1e-3),
        "unit": "candidates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong" if run.strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "n_rec_total": run.n_total, "n_rec_per_gpu": run.n_local, "n_nys": L, "batch": b,
                   "kernel": fam + (" / predictive_covariance (n_obs=200)" if args.pred_cov else " / kernel mode"),
                   "d": d, "mode": args.mode, "l2": "256 MiB buffer written between steps (flush)",
                   "parallelism": "row-sharded candidates x%d" % world},
        "e2e": {"value": run.n_total * e_steps / (e2e_ms * 1e-3), "unit": "candidates/s", "ms_per_step": e2e_ms / e_steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "bytes_are": "per rank (each rank copies its own shard of the candidates and weights)"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "roofline_stream": {
            "kernel": "update_compact (weight update + alive-list compaction, moves the record rows): "
                      "the launch over all candidates",
            "bound": "hbm", "achieved": uc_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": (uc_gbs / hbm_peak) if uc_gbs else None, "peak_source": hbm_src,
            "bytes": uc_big[1] if uc_big else None, "ms": uc_big[0] if uc_big else None,
            "all_launches_ms_per_step": uc_ms / max(args.steps, 1),
        },
        # CUDA-event time per step of each instrumented stage on the launching stream; car_step_graph = one replay of
        # the Caratheodory step (null space + elimination + survivor ranks), all_reduce = the NCCL collective
        "stage_ms_per_step": {k: v[1] / max(args.steps, 1) for k, v in timing.items()},
        "stage_calls_per_step": {k: v[0] / max(args.steps, 1) for k, v in timing.items()},
        "car": {name_: {"calls_per_step": timing[name_][0] / max(args.steps, 1),
                        "us_per_sequential_step": 1e3 * timing[name_][1] / timing[name_][2] if timing[name_][2] else None}
                for name_ in ("car_panel", "car_cols", "car_cluster", "car_eliminate", "car_step_graph")
                if name_ in timing},
    }
    unattributed = step_ms - sum(v for k, v in out["stage_ms_per_step"].items()
                                 if k not in ("car_panel", "car_cols", "car_cluster") or "car_step_graph" not in timing)
    out["stage_ms_per_step"]["other (range finder, projection GEMM, host syncs, launch gaps)"] = unattributed
    out.update(extras)
    if shard_check is not None:
        out["sharded_parity"] = shard_check
    if world == 1 and not args.no_extras and args.mode == "fast" and name in ("c1", "c2", "c3"):
        sample = min(args.cpu_sample if args.cpu_sample > 0 else 100_000, n_rec)
        out["parity"] = parity_block(run, sample)
        # the mode that executes the reference's own op sequence on this device (bitwise-asymmetric Gram from the
        # kernel object, the reference's PSD gate, torch.svd_lowrank, null space from the full torch.linalg.svd)
        sober_b200.options.set_mode("parity")
        psteps = 3
        pms, _, _ = run.timed(psteps, 1, flush)
        sober_b200.options.set_mode(args.mode)
        out["parity_mode"] = {"ms_per_step": pms / psteps, "steps": psteps,
                              "value": run.n_total * psteps / (pms * 1e-3), "unit": "candidates/s"}
        if name == "c2":
            out["reference_gpu"] = reference_gpu_leg(run, flush)
    if world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample if args.cpu_sample > 0 else 100_000, n_rec)
        rate, times = cpu_reference_rate(name, sample, 1, 0, threads)
        out["cpu_baseline"] = {"value": rate, "unit": "candidates/s", "cores": threads, "kind": "port",
                               "sample": "oracle/rchq.py on %d of the workload's candidates (full n_nys/batch), "
                                         "1 run of %.1f s" % (sample, times[0])}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
