#!/usr/bin/env python
"""Benchmark of the RCHQ batch-selection hot path (BASELINE.json metric: recombination candidates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5|c1|c3|c4]

One "step" = one full ``recombination(...)`` call (Nystrom block + all grouped passes + CAR + compaction) over one
batch of synthetic candidates.  Default workload = BASELINE.json configs[1] (Hartmann-6D shape: n_rec = 1e6 per
GPU, n_nys = 1000, batch = 200, Matern-5/2, float64).  With N > 1 (torchrun, one rank per GPU) the candidates are
row-sharded, n_rec = 1e6 PER GPU (weak scaling); the one collective per iteration is the all-reduce of the group
sums.  Rank 0 prints ONE JSON line.

``--impl reference`` times the reference algorithm's CPU path (the oracle restatement of SOBER/_rchq.py, which is
bit-identical to it; the reference itself is Python and /root/reference does not exist on the GPU box) on a
bounded sample of the same workload with all host threads.
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
# stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

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
    g = torch.Generator().manual_seed(5)
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
    """The reference algorithm on the host cores: oracle/rchq.py (bit-identical to SOBER/_rchq.py on the CPU) on a
    bounded sample of the workload.  Returns candidates/sec (best step) and the per-step seconds."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pred-cov", action="store_true",
                    help="use Kernel(model, 'predictive_covariance') with a synthetic 200-observation GP (SURVEY 8d)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    n_rec, d, L, b, fam, ls, desc = WORKLOADS[name]
    threads = os.cpu_count() or 1

    # --------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(args.cpu_sample, n_rec)
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        rate, times = cpu_reference_rate(name, sample, steps, warm, threads)
        print(json.dumps({
            "impl": "reference", "metric": "recombination candidates/sec", "value": rate, "unit": "candidates/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * min(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_nys": L, "batch": b, "kernel": fam, "sample_n_rec": sample},
            "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": threads, "kind": "port",
                             "sample": "oracle/rchq.py (bit-identical restatement of SOBER/_rchq.py) on %d of the "
                                       "workload's candidates, full n_nys/batch, torch CPU f64, %d threads; the "
                                       "reference materialises the (E,L,S) Gram so its memory grows as 40 B x N x "
                                       "n_nys" % (sample, threads)},
            "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # --------------------------------------------------------------------------------------------------
    import sober_b200
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        sober_b200.enable_sharding()
    strong = name == "c5"
    n_local = n_rec // world if strong else n_rec
    n_total = n_local * world
    X, mu = synth(name, n_local, 100 + rank, dev)
    if world > 1:
        tot = mu.sum()
        dist.all_reduce(tot)
        mu /= tot
    else:
        mu /= mu.sum()
    # landmarks: replicated; drawn from rank 0's shard and broadcast
    Z = X[torch.randperm(n_local, device=dev, generator=torch.Generator(device=dev).manual_seed(1))[:L]].clone()
    if world > 1:
        dist.broadcast(Z, 0)
    kern = make_kernel(name, dev, args.pred_cov)
    from sober_b200 import _rchq
    ops = _rchq._ops()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(weights):
        torch.manual_seed(7)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return sober_b200.recombination(X, Z, b, kern, dev, torch.float64, init_weights=weights)

    sober_b200.options.set_mode(args.mode)
    for _ in range(args.warmup):
        flush.zero_()
        idx, w = step(mu.clone())
    barrier()

    # ---- timed region: K steps, inputs resident in HBM ----
    ops.timing = {}
    from sober_b200 import _linalg
    launches0 = ops.launches + _linalg.launches
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for _ in range(args.steps):
            flush.zero_()
            idx, w = step(mu.clone())
        t_end.record()
        barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    timing = ops.timing_summary()
    uc_big = ops.timing_largest("update_compact")
    k1_big = ops.timing_largest("group_accumulate")
    ops.timing = None
    launches = ops.launches + _linalg.launches - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t)
    assert len(idx) <= b and abs(float(w.sum()) - 1.0) < 1e-9, "benchmark result failed its invariants"

    # ---- e2e: host (pinned) buffers in, host results out, copies inside the timed region ----
    Xh = X.cpu().pin_memory()
    muh = mu.cpu().pin_memory()
    wh = torch.empty_like(muh).pin_memory()
    for _ in range(2):
        wh.copy_(muh)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            idx_e, w_e = sober_b200.recombination(Xh, Z, b, kern, dev, torch.float64, init_weights=wh)
    barrier()
    e_steps = max(2, args.steps // 2)
    e0 = time.perf_counter()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    for _ in range(e_steps):
        flush.zero_()
        torch.manual_seed(7)
        wh.copy_(muh)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            idx_e, w_e = sober_b200.recombination(Xh, Z, b, kern, dev, torch.float64, init_weights=wh)
        idx_host, w_host = idx_e.cpu(), w_e.cpu()
    e_ev.record()
    barrier()
    e2e_ms = s_ev.elapsed_time(e_ev)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t)
    h2d = Xh.numel() * 8 + muh.numel() * 8
    d2h = idx_host.numel() * 8 + w_host.numel() * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1, FP64-pipe bound) and of the streaming pass (HBM bound) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # FP64 peak: not in MEASURED_PEAKS.json (HBM and bf16 only) -> measured here with a dependent-chain DFMA probe
    iters = 1 << 16
    blocks = 148 * 8
    for _ in range(2):
        ops.fp64_probe(blocks, iters)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    ops.fp64_probe(blocks, iters)
    p1.record()
    torch.cuda.synchronize()
    fp64_peak = blocks * 256 * iters * 16 / (p0.elapsed_time(p1) * 1e-3) / 1e12

    flop_per_pair = 2 * d + 2 + C_K[fam]
    k1_calls, k1_ms, k1_pairs = timing.get("group_accumulate", (0, 0.0, 0))
    k1_tflops = k1_pairs * flop_per_pair / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else None
    uc_calls, uc_ms, uc_bytes = timing.get("update_compact", (0, 0.0, 0))
    # HBM roofline of the streaming pass: the pass over ALL candidates (first iteration); later iterations halve
    uc_gbs = uc_big[1] / (uc_big[0] * 1e-3) / 1e9 if uc_big and uc_big[0] > 0 else None
    bits = fam == "tanimoto" and d > 8
    words = (d + 63) // 64
    car_calls, car_ms, car_steps = timing.get("car_eliminate", (0, 0.0, 0))
    step_ms = elapsed_ms / args.steps

    out = {
        "metric": "recombination candidates/sec", "value": n_total * args.steps / (elapsed_ms * 1e-3),
        "unit": "candidates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "n_rec_total": n_total, "n_rec_per_gpu": n_local, "n_nys": L, "batch": b,
                   "kernel": fam + (" / predictive_covariance (n_obs=200)" if args.pred_cov else " / kernel mode"),
                   "d": d, "mode": args.mode, "l2": "256 MiB buffer written between steps (flush)",
                   "parallelism": "row-sharded candidates x%d" % world},
        "e2e": {"value": n_total * e_steps / (e2e_ms * 1e-3), "unit": "candidates/s", "ms_per_step": e2e_ms / e_steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": ({
            "kernel": "group_accumulate (K1: fused cross-kernel + weighted group sums, record layout)",
            "bound": "fp64",          # FP64 FMA pipe; the hbm|tensor enum has no entry for it
            "achieved": k1_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (k1_tflops / fp64_peak) if k1_tflops else None,
            # DRAM bytes of the largest launch (10^6 candidates x 1000 landmarks) from the ncu --set full capture
            # profiles/r01_ncu_final_k1_and_car_cols.txt: 64.1 MB read + 7.6 MB written; the 64-byte records alone are
            # 64 MB, i.e. no re-reads.  Only quoted for the shape it was captured on.
            "traffic": (71.7e6 if (name == "c2" and world == 1 and not args.pred_cov) else None),
            "traffic_unit": "bytes per largest launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "peak_source": "measured in this run: dependent-chain DFMA probe (sober_fp64_probe), 2 flop per FMA",
            "algorithmic_flop_per_pair": flop_per_pair, "pairs_per_step": k1_pairs / max(args.steps, 1),
            "launches": k1_calls, "ms_per_step": k1_ms / max(args.steps, 1),
            "share_of_step": k1_ms / elapsed_ms if elapsed_ms else None,
            "largest_launch": {"ms": k1_big[0], "pairs": k1_big[1],
                               "tflops": k1_big[1] * flop_per_pair / (k1_big[0] * 1e-3) / 1e12} if k1_big else None,
        } if not bits else {
            "kernel": "group_accumulate (K1, bit-packed Tanimoto: popcount(x & z) + FP64 ratio)",
            "bound": "int",           # integer pipe (AND + POPC); no measured peak for it on this pool
            "achieved": k1_pairs * words / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None, "peak": None,
            "unit": "G word-op/s (64-bit AND+POPC)", "frac": None, "traffic": None,
            "algorithmic_word_ops_per_pair": words, "pairs_per_step": k1_pairs / max(args.steps, 1),
            "launches": k1_calls, "ms_per_step": k1_ms / max(args.steps, 1),
            "share_of_step": k1_ms / elapsed_ms if elapsed_ms else None,
        }),
        "roofline_stream": {
            "kernel": "update_compact (weight update + alive-list compaction, moves the record rows): "
                      "the launch over all candidates",
            "bound": "hbm", "achieved": uc_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": (uc_gbs / hbm_peak) if uc_gbs else None, "peak_source": hbm_src,
            "bytes": uc_big[1] if uc_big else None, "ms": uc_big[0] if uc_big else None,
            "all_launches_ms_per_step": uc_ms / max(args.steps, 1),
        },
        "stage_ms_per_step": {k: v[1] / max(args.steps, 1) for k, v in timing.items()},
        "car": {name_: {"calls_per_step": timing[name_][0] / max(args.steps, 1),
                        "us_per_sequential_step": 1e3 * timing[name_][1] / timing[name_][2] if timing[name_][2] else None}
                for name_ in ("car_cols", "car_cluster", "car_eliminate") if name_ in timing},
    }
    if world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample, n_rec)
        rate, times = cpu_reference_rate(name, sample, 1, 0, threads)
        out["cpu_baseline"] = {"value": rate, "unit": "candidates/s", "cores": threads, "kind": "port",
                               "sample": "oracle/rchq.py on %d of the workload's candidates (full n_nys/batch), "
                                         "1 run of %.1f s" % (sample, times[0])}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
