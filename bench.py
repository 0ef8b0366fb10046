#!/usr/bin/env python3
"""bench.py — headline benchmark of the crypto12381 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): G1 MSM points/s at n = 2^20 (configs[1]); one step = one G1 multi-scalar sum over 2^20
seeded synthetic (point, scalar) terms per GPU.  With N GPUs every rank owns 2^20 terms of one N*2^20-term sum
(weak scaling): per-rank partial -> one all-gather of the 96-byte partials over NCCL -> every rank adds them.
  value      whole-job points/s with inputs resident in HBM (device entries, CUDA events, max over ranks)
  e2e        the same through the host-pointer C-ABI call c12381_g1_msm (pinned host buffers, H2D + D2H inside)
  roofline   dominant kernel k_accumulate against the integer-multiply peak measured live by c12381_probe
  cpu_baseline   the reference's own CPU path (oracle/_ref: MIRACL ECP_muln via the unmodified bridge) on a
             bounded sample of the same points, all host threads, rank 0 only
  secondary  batched 4-pair pairing products (BASELINE configs[3]) as pairings/s, same run
`--impl reference` times only the reference CPU path (rank 0), same metric/unit/config."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "g1_msm_points_per_s"
UNIT = "points/s"
FP_MUL_PER_BUCKET_ADD = 10      # XYZZ mixed addition: 8 M + 2 S
MAC_PER_FP_MUL = 300            # 12x12 product + 12x12 reduction + 12 quotient digits
R_TOP = 0x73


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--pairing-instances", type=int, default=1 << 16)
    ap.add_argument("--cpu-sample-log-n", type=int, default=18)
    ap.add_argument("--g2-log-n", type=int, default=18)
    ap.add_argument("--bbs-log-b", type=int, default=16)
    ap.add_argument("--sweep-max-log-n", type=int, default=24)
    ap.add_argument("--no-secondary", action="store_true")
    return ap.parse_args()


def config(args, world):
    n = 1 << args.log_n
    return {"workload": f"G1 MSM sweep point n=2^{args.log_n} per GPU (BASELINE configs[1]), seeded random points k_i*G and scalars < r",
            "n_per_gpu": n, "n_total": n * world, "parallelism": f"points sharded over {world} GPU(s), one all-gather of 96-byte partials",
            "l2": "explicit 256 MiB flush write between timed steps; per-step working set (128 MiB inputs + ~0.9 GiB scratch) also exceeds the 126 MB L2"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None), "samples": len(rows), "reasons": reasons}


def rand_scalars(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] = rng.integers(0, R_TOP, size=n, dtype=np.uint8)   # top byte < 0x73 keeps every scalar below r
    return a


# ---- the reference arm ---------------------------------------------------------------------------------------------
def reference_points(n, seed):
    """n seeded points k_i*G made by the REFERENCE on the host (all threads): the CPU arm needs no GPU."""
    from oracle import ref
    ks = rand_scalars(n, seed).tobytes()
    return ref.g1_fixed_base_mul(ks, ref.hardware_threads())


def time_reference(points: bytes, scalars: bytes, steps: int, warmup: int):
    from oracle import ref
    threads = ref.hardware_threads()
    n = len(scalars) // 32
    for _ in range(warmup):
        ref.g1_msm(points, scalars, 0, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = ref.g1_msm(points, scalars, 0, threads)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, threads, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref12381.so was not built (needs /root/reference at build time)"}))
        return
    n = 1 << args.cpu_sample_log_n
    pts = reference_points(n, 1000)
    ss = rand_scalars(n, 2000).tobytes()
    v, dt, threads, _ = time_reference(pts, ss, max(1, args.steps), max(1, args.warmup))
    sample = f"2^{args.cpu_sample_log_n}-term G1 sum per step (bounded sample of the 2^{args.log_n} workload), sum_of_products -> MIRACL ECP_muln, chunked over {threads} host threads"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 (7x58-bit limbs, CPU)",
            "data": "synthetic", "config": config(args, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---- our arm ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from crypto12381_b200 import _lib, device as dv
    from crypto12381_b200.distributed import g1_msm_sharded
    _lib.init(local)
    lib = _lib.lib()
    n = 1 << args.log_n
    dev = torch.device("cuda", local)

    # synthetic seeded inputs: points k_i*G made on the GPU by the fixed-base kernel (setup, untimed)
    h_k = torch.from_numpy(rand_scalars(n, 1000 + rank)).reshape(-1)
    h_s = torch.from_numpy(rand_scalars(n, 2000 + rank)).reshape(-1).pin_memory()
    d_s = h_s.to(dev)
    d_p = dv.g1_fixed_base_mul_batch(h_k.to(dev))
    dv.sync_status()
    h_p = d_p.cpu().pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty(49, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return g1_msm_sharded(d_p, d_s)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    acc_ms, tot_ms = [], []
    l0 = int(lib.c12381_launch_count())
    barrier()
    t0 = time.time()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        res = step()
        b.record()
        b.synchronize()
        st = dv.last_msm_stats()
        acc_ms.append(st["accumulate_ms"])
        tot_ms.append(st["total_ms"])
    barrier()
    t1 = time.time()
    launches = int(lib.c12381_launch_count()) - l0
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    clocks = sampler.stop(t0, t1)
    result = bytes(res.cpu().numpy())
    stats = dv.last_msm_stats()

    # e2e: host-pointer C-ABI call, pinned host buffers, H2D + D2H inside the timed region (per rank, its own shard)
    import ctypes
    h_out = torch.empty(49, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        _lib.check(lib.c12381_g1_msm(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
    barrier()
    te = time.perf_counter()
    for _ in range(args.steps):
        _lib.check(lib.c12381_g1_msm(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
    e2e_ms = (time.perf_counter() - te) / args.steps * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    if world == 1:
        assert bytes(h_out.numpy()) == result, "host-pointer and device-pointer entries disagree"

    line = None
    if rank == 0:
        # integer-multiply peak, measured live on this GPU (probe kind 2: IMAD.WIDE 32x32+64 multiply-adds)
        probes = {k: dv.probe(i, 4000) for i, k in enumerate(["imad", "madc_pairs", "imad_wide", "fp_mul", "fp_sqr"])}
        peak_gmacs = max(probes["imad_wide"]["gops"], probes["madc_pairs"]["gops"] / 2.0)
        adds = stats["bucket_adds"]
        acc = statistics.mean(acc_ms)
        achieved = adds * FP_MUL_PER_BUCKET_ADD * MAC_PER_FP_MUL / (acc * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes of one k_accumulate launch from the committed ncu --set full capture (same n, same plan)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)["k_accumulate<Fp>"]
            if args.log_n == 20:
                traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
        except Exception:
            pass
        roofline = {"bound": "int32_mad", "kernel": "k_accumulate<Fp>", "achieved": achieved, "peak": peak_gmacs, "unit": "GMAC/s (32x32->64 multiply-adds)",
                    "frac": achieved / peak_gmacs, "traffic": traffic, "traffic_unit": "DRAM bytes per launch", "traffic_source": traffic_src,
                    "gather_bytes_algorithmic": adds * 96,
                    "peak_source": "measured live: c12381_probe kind 2 (mad.wide.u32 chains) / kind 1 (mad.lo.cc+madc.hi.cc pairs), same GPU, same run",
                    "algorithmic": f"{adds} bucket additions/launch x {FP_MUL_PER_BUCKET_ADD} Fp-mul x {MAC_PER_FP_MUL} MAC",
                    "kernel_ms": acc, "kernel_share_of_step": acc / statistics.mean(tot_ms), "window_bits": stats["window_bits"],
                    "fp_mul_gops": probes["fp_mul"]["gops"], "fp_sqr_gops": probes["fp_sqr"]["gops"], "probes": probes}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm = json.load(f)["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        passes = (stats["window_bits"] + 7) // 8
        ph = stats["phases_ms"]
        sort_bytes = adds * 8 * 2 * passes      # (key, index) pairs read and written once per radix pass
        roofline["phases_ms"] = ph
        roofline["hbm_phase"] = {"what": "segmented radix sort of the (bucket key, term index) pairs: bucket scatter", "algorithmic_bytes": sort_bytes,
                                 "ms": ph["sort"], "achieved_gbs": sort_bytes / (ph["sort"] * 1e-3) / 1e9 if ph["sort"] > 0 else None,
                                 "peak_gbs": hbm, "frac": (sort_bytes / (ph["sort"] * 1e-3) / 1e9 / hbm) if ph["sort"] > 0 else None,
                                 "peak_source": hbm_src, "radix_passes": passes}
        cpu = None
        try:
            from oracle import ref
            if ref.available():
                m = 1 << args.cpu_sample_log_n
                pts, ss = bytes(h_p[:96 * m].numpy()), bytes(h_s[:32 * m].numpy())
                v, dt, threads, cpu_out = time_reference(pts, ss, 1, 0)
                # same inputs, so the sample doubles as a parity check of the CUDA path against the reference
                got = bytes(dv.g1_msm(d_p[:96 * m], d_s[:32 * m]).cpu().numpy())
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "bit_exact_vs_gpu_on_sample": got == cpu_out,
                       "sample": f"first 2^{args.cpu_sample_log_n} terms of rank 0's inputs, one pass ({dt:.2f} s), sum_of_products -> MIRACL ECP_muln chunked over {threads} host threads (oracle/_ref, -O2)"}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        line = {"metric": METRIC, "value": n * world / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (12x32-bit limbs, Montgomery R=2^384)",
                "data": "synthetic", "config": config(args, world), "clocks": clocks,
                "e2e": {"value": n * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * 128, "d2h_bytes_per_step": 49, "ms_per_step": e2e_ms,
                        "call": "c12381_g1_msm (host pointers, pinned), per rank on its own shard"},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "result_hex": result.hex()}

    # secondary: batched 4-pair pairing products (BASELINE configs[3]); instances sharded over ranks, no collective
    if not args.no_secondary:
        B, k = max(1, args.pairing_instances // world), 4
        a = torch.from_numpy(rand_scalars(B * k, 3000 + rank)).reshape(-1).to(dev)
        b = torch.from_numpy(rand_scalars(B * k, 4000 + rank)).reshape(-1).to(dev)
        g1, g2 = dv.g1_fixed_base_mul_batch(a), dv.g2_fixed_base_mul_batch(b)
        gt = torch.empty(B * 576, dtype=torch.uint8, device=dev)
        dv.pairing_product_batch(g1, g2, k, gt)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dv.pairing_product_batch(g1, g2, k, gt)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item())
            sec = {"metric": "pairings_per_s", "value": B * world * k / (ms * 1e-3), "unit": "pairings/s", "products_per_s": B * world / (ms * 1e-3),
                   "instances": B * world, "pairs_per_instance": k, "ms": ms,
                   "fp_mul_per_instance_ref_count": 31600, "gmacs": B * 31600 * MAC_PER_FP_MUL / (ms * 1e-3) / 1e9}
            sec["frac_of_int32_mad_peak"] = sec["gmacs"] / line["roofline"]["peak"]
            if k == 4:      # the kernel bodies' own count (tools/count_fp_mul.py, host build with -DC12_COUNT_FP_MUL)
                sec["fp_mul_per_instance_kernel_count"] = 30751
                sec["frac_of_int32_mad_peak_kernel_count"] = sec["frac_of_int32_mad_peak"] * 30751 / 31600
            try:
                from oracle import ref
                if ref.available():
                    m = min(B, 4 * ref.hardware_threads())
                    p1, p2 = bytes(g1[:96 * k * m].cpu().numpy()), bytes(g2[:192 * k * m].cpu().numpy())
                    tc = time.perf_counter()
                    want = ref.pairing_product_batch(p1, p2, k, 1, ref.hardware_threads())
                    dtc = time.perf_counter() - tc
                    sec["cpu_baseline"] = {"value": m * k / dtc, "unit": "pairings/s", "cores": ref.hardware_threads(), "kind": "reference",
                                           "sample": f"{m} instances x {k} pairs", "bit_exact_vs_gpu_on_sample": want == bytes(gt[:576 * m].cpu().numpy())}
            except Exception as e:
                sec["cpu_baseline"] = {"value": None, "sample": f"unavailable: {e}"}
            line["secondary"] = sec
    # secondary 2: G2 MSM (BASELINE configs[2]), n = 2^18 per GPU, partial -> all-gather -> merge as for G1
    if not args.no_secondary:
        from crypto12381_b200.distributed import g2_msm_sharded
        n2 = 1 << args.g2_log_n
        k2 = torch.from_numpy(rand_scalars(n2, 5000 + rank)).reshape(-1).to(dev)
        s2 = torch.from_numpy(rand_scalars(n2, 6000 + rank)).reshape(-1).to(dev)
        p2 = dv.g2_fixed_base_mul_batch(k2)
        for _ in range(2):
            r2 = g2_msm_sharded(p2, s2)
        barrier()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        e0.record()
        for _ in range(reps):
            r2 = g2_msm_sharded(p2, s2)
        e1.record()
        barrier()
        st2 = dv.last_msm_stats()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item())
            g2 = {"metric": "g2_msm_points_per_s", "value": n2 * world / (ms * 1e-3), "unit": "points/s", "n_per_gpu": n2, "ms": ms,
                  "window_bits": st2["window_bits"], "phases_ms": st2["phases_ms"], "result_hex": bytes(r2.cpu().numpy()).hex()}
            try:
                from oracle import ref
                if ref.available():
                    m = 1 << 11
                    tc = time.perf_counter()
                    want = ref.g2_msm(bytes(p2[:192 * m].cpu().numpy()), bytes(s2[:32 * m].cpu().numpy()), ref.hardware_threads())
                    dtc = time.perf_counter() - tc
                    got = bytes(dv.g2_msm(p2[:192 * m], s2[:32 * m]).cpu().numpy())
                    g2["cpu_baseline"] = {"value": m / dtc, "unit": "points/s", "cores": ref.hardware_threads(), "kind": "reference",
                                          "sample": f"first 2^11 terms: per-term PAIR_G2mul + ECP2_add loop (g2_point.hpp:202-236) over {ref.hardware_threads()} host threads",
                                          "bit_exact_vs_gpu_on_sample": want == got}
            except Exception as e:
                g2["cpu_baseline"] = {"value": None, "sample": f"unavailable: {e}"}
            line["secondary_g2_msm"] = g2
    # secondary: the G1 sweep of BASELINE configs[1], n = 2^10 .. 2^24 (rank 0 only, device-resident, same pipeline)
    if not args.no_secondary and args.sweep_max_log_n >= 10 and rank == 0:
        gsw = torch.Generator(device=dev).manual_seed(9000)
        nmax = 1 << args.sweep_max_log_n
        kk = torch.randint(0, 256, (nmax, 32), dtype=torch.uint8, device=dev, generator=gsw)
        sw_s = torch.randint(0, 256, (nmax, 32), dtype=torch.uint8, device=dev, generator=gsw)
        kk[:, 0] %= R_TOP
        sw_s[:, 0] %= R_TOP
        sw_p = dv.g1_fixed_base_mul_batch(kk.reshape(-1))
        sw_s = sw_s.reshape(-1)
        del kk
        sweep = []
        for ln in range(10, args.sweep_max_log_n + 1, 2):
            m = 1 << ln
            reps = 10 if ln <= 18 else 3
            dv.g1_msm(sw_p[:96 * m], sw_s[:32 * m])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                dv.g1_msm(sw_p[:96 * m], sw_s[:32 * m])
            e1.record()
            torch.cuda.synchronize()
            ms_s = e0.elapsed_time(e1) / reps
            sweep.append({"log_n": ln, "ms": ms_s, "points_per_s": m / (ms_s * 1e-3), "window_bits": dv.last_msm_stats()["window_bits"]})
        dv.sync_status()
        line["secondary_g1_sweep"] = sweep
        del sw_p, sw_s
    # secondary: batched scalar multiplication: g^x from the fixed-base table and P^k, rank 0 only,
    # device-resident; CPU beside it: the reference's multiply (PAIR_G1mul) on a sample over all host threads
    if not args.no_secondary and args.sweep_max_log_n >= 10 and rank == 0:
        try:
            nb = 1 << 18
            xs_t = torch.from_numpy(rand_scalars(nb, 8100)).reshape(-1).to(dev)
            ks_t = torch.from_numpy(rand_scalars(nb, 8200)).reshape(-1).to(dev)

            def timed_ms(fn, reps=3):
                fn()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    out = fn()
                a1.record()
                torch.cuda.synchronize()
                return a0.elapsed_time(a1) / reps, out

            ms_fb, pts_t = timed_ms(lambda: dv.g1_fixed_base_mul_batch(xs_t))
            ms_mul, enc_t = timed_ms(lambda: dv.g1_mul_batch(pts_t, ks_t))
            ms_fb2, pts2_t = timed_ms(lambda: dv.g2_fixed_base_mul_batch(xs_t[:32 * (nb // 4)]))
            ms_mul2, _ = timed_ms(lambda: dv.g2_mul_batch(pts2_t, ks_t[:32 * (nb // 4)]))
            dv.sync_status()
            sm = {"g1_fixed_base_per_s": nb / (ms_fb * 1e-3), "g1_mul_per_s": nb / (ms_mul * 1e-3),
                  "g2_fixed_base_per_s": (nb // 4) / (ms_fb2 * 1e-3), "g2_mul_per_s": (nb // 4) / (ms_mul2 * 1e-3),
                  "batch": {"g1": nb, "g2": nb // 4}, "unit": "scalar multiplications/s",
                  "ms": {"g1_fixed_base": ms_fb, "g1_mul": ms_mul, "g2_fixed_base": ms_fb2, "g2_mul": ms_mul2}}
            try:
                from oracle import ref
                if ref.available():
                    m = 64 * ref.hardware_threads()
                    pb, kb = bytes(pts_t[:96 * m].cpu().numpy()), bytes(ks_t[:32 * m].cpu().numpy())
                    tc = time.perf_counter()
                    want = ref.g1_mul_batch(pb, kb, ref.hardware_threads())
                    dtc = time.perf_counter() - tc
                    sm["cpu_baseline"] = {"value": m / dtc, "unit": "G1 scalar multiplications/s", "cores": ref.hardware_threads(), "kind": "reference",
                                          "sample": f"{m} x multiply(point1&, big) -> PAIR_G1mul over {ref.hardware_threads()} host threads",
                                          "bit_exact_vs_gpu_on_sample": want == bytes(enc_t[:49 * m].cpu().numpy())}
            except Exception as e:
                sm["cpu_baseline"] = {"value": None, "sample": f"unavailable: {e}"}
            line["secondary_scalar_mul"] = sm
        except Exception as e:      # a secondary must never cost the headline line
            line["secondary_scalar_mul"] = {"error": str(e)}
    # secondary 3: BBS+ batch verification (BASELINE configs[4]): 2^16 signatures x 10 message blocks over all ranks,
    # instances split across ranks with no collective; the timed region is the whole device pipeline of bbs_plus.verify_batch_device
    if not args.no_secondary and args.bbs_log_b > 0:
        import numpy as np
        from crypto12381_b200 import bbs_plus
        Bs, nmsg = max(1, (1 << args.bbs_log_b) // world), 10
        R_ORD = bbs_plus.R
        rngb = np.random.default_rng(7000 + rank)
        gens = dv.g1_fixed_base_mul_batch(torch.from_numpy(rand_scalars(nmsg + 2, 7100)).reshape(-1).to(dev))        # g1, h0, h_j
        g2b = dv.g2_fixed_base_mul_batch(torch.from_numpy(rand_scalars(1, 7200)).reshape(-1).to(dev))
        gamma = int.from_bytes(rand_scalars(1, 7300).tobytes(), "big") % R_ORD
        wb = dv.g2_decompress_batch(dv.g2_mul_batch(g2b, torch.frombuffer(bytearray(gamma.to_bytes(32, "big")), dtype=torch.uint8).to(dev)))
        bases_g2 = torch.cat((wb, g2b))
        neg_g2 = torch.frombuffer(bytearray(bbs_plus._neg_g2(bytes(g2b.cpu().numpy()))), dtype=torch.uint8).to(dev)
        xs, rs = rand_scalars(Bs, 7400 + rank), rand_scalars(Bs, 7500 + rank)
        ms = rngb.integers(0, 256, size=(Bs, nmsg, 32), dtype=np.uint8)
        ms[:, :, 0] = 1                                        # encode_to<Zp>: 2^248 + a 31-byte block
        one = np.zeros((Bs, 1, 32), dtype=np.uint8)
        one[:, :, 31] = 1
        sc_g1 = torch.from_numpy(np.concatenate((one, rs.reshape(Bs, 1, 32), ms), axis=1).reshape(-1)).to(dev)
        sc_g2 = torch.from_numpy(np.concatenate((one, xs.reshape(Bs, 1, 32)), axis=1).reshape(-1)).to(dev)
        # sign on the GPU: A = (g1 h0^r prod h_j^m_j)^(1/(gamma + x)); the Zp inverse stays on the host as in the reference
        inv = b"".join(pow((gamma + int.from_bytes(x.tobytes(), "big")) % R_ORD, -1, R_ORD).to_bytes(32, "big") for x in xs)
        Bp = dv.g1_multi_fixed_base_batch(gens, sc_g1)
        sigA = dv.g1_mul_batch(Bp, torch.frombuffer(bytearray(inv), dtype=torch.uint8).to(dev))      # compressed 49 B
        v = bbs_plus.verify_batch_device(gens, bases_g2, neg_g2, sigA, sc_g1, sc_g2)
        dv.sync_status()
        all_ok = bool((v == 1).all().item())
        sigA_bad = sigA.clone()
        sigA_bad[49:98] = sigA[0:49]                            # signature 1 gets signature 0's A
        v_bad = bbs_plus.verify_batch_device(gens, bases_g2, neg_g2, sigA_bad, sc_g1, sc_g2)
        bad_ok = bool(v_bad[1].item() == 0 and (v_bad[2:] == 1).all().item())
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        e0.record()
        v = bbs_plus.verify_batch_device(gens, bases_g2, neg_g2, sigA, sc_g1, sc_g2)
        e1.record()
        barrier()
        # the same batch through the random-linear-combination check (one verdict per rank's shard; bbs_plus.verify_batch_aggregate)
        xi = [int.from_bytes(x.tobytes(), "big") for x in xs]
        rho_i = [int.from_bytes(r16.tobytes(), "big") | 1 for r16 in rngb.integers(0, 256, size=(Bs, 16), dtype=np.uint8)]
        tb = lambda vals: torch.frombuffer(bytearray(b"".join(v.to_bytes(32, "big") for v in vals)), dtype=torch.uint8).to(dev)
        rho_t, rhox_t, nrho_t = tb(rho_i), tb([r * x % R_ORD for r, x in zip(rho_i, xi)]), tb([R_ORD - r for r in rho_i])
        va = bbs_plus.verify_batch_aggregate_device(gens, bases_g2, sigA, sc_g1, rho_t, rhox_t, nrho_t)
        va_bad = bbs_plus.verify_batch_aggregate_device(gens, bases_g2, sigA_bad, sc_g1, rho_t, rhox_t, nrho_t)
        dv.sync_status()
        agg_ok = bool(va.item() == 1 and va_bad.item() == 0)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        bbs_plus.verify_batch_aggregate_device(gens, bases_g2, sigA, sc_g1, rho_t, rhox_t, nrho_t)
        a1.record()
        barrier()
        agg_ms = a0.elapsed_time(a1)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms_t = float(t.item())
            line["secondary_bbs_plus_verify"] = {"aggregate_check": {"what": "random-linear-combination batch check (an extension: one verdict per shard): B_i products, two G1 MSMs, one 2-pair pairing check",
                                                                     "ms_rank0": agg_ms, "signatures_per_s_rank0": Bs / (agg_ms * 1e-3), "valid_accepted_tampered_rejected": agg_ok},"metric": "bbs_plus_verifications_per_s", "value": Bs * world / (ms_t * 1e-3), "unit": "signatures/s",
                                                 "signatures": Bs * world, "message_blocks": nmsg, "ms": ms_t, "all_valid_accepted": all_ok,
                                                 "tampered_rejected": bad_ok,
                                                 "pipeline": "decompress A; w + x g2; g1 + r h0 + sum m_j h_j (window tables over the 12 shared bases); 2-pair pairing check"}
            try:   # the reference's arithmetic for the same verifications (bbs+.cpp:57-73) from its own bridge functions, all host threads
                from concurrent.futures import ThreadPoolExecutor
                from oracle import ref
                if ref.available():
                    mS, th = 4 * ref.hardware_threads(), ref.hardware_threads()
                    hb, h2, hA = bytes(gens.cpu().numpy()), bytes(bases_g2.cpu().numpy()), bytes(dv.g1_decompress_batch(sigA[:49 * mS]).cpu().numpy())
                    s1, s2 = bytes(sc_g1[:32 * 12 * mS].cpu().numpy()), bytes(sc_g2[:64 * mS].cpu().numpy())
                    ng = bytes(neg_g2.cpu().numpy())
                    tc = time.perf_counter()
                    with ThreadPoolExecutor(th) as ex:     # per signature: the live product loop (double_multiply pairs) and w * g2^x
                        Bc = list(ex.map(lambda i: ref.g1_msm(hb, s1[384 * i:384 * (i + 1)], 1, 1), range(mS)))
                        Wc = list(ex.map(lambda i: ref.g2_msm(h2, s2[64 * i:64 * (i + 1)], 1), range(mS)))
                    dt1 = time.perf_counter() - tc
                    Ba = bytes(dv.g1_decompress_batch(torch.frombuffer(bytearray(b"".join(Bc)), dtype=torch.uint8).to(dev)).cpu().numpy())
                    Wa = bytes(dv.g2_decompress_batch(torch.frombuffer(bytearray(b"".join(Wc)), dtype=torch.uint8).to(dev)).cpu().numpy())
                    p1 = b"".join(hA[96 * i:96 * i + 96] + Ba[96 * i:96 * i + 96] for i in range(mS))
                    p2 = b"".join(Wa[192 * i:192 * i + 192] + ng for i in range(mS))
                    tc = time.perf_counter()
                    gtc = ref.pairing_product_batch(p1, p2, 2, 1, th)
                    dtc = dt1 + time.perf_counter() - tc
                    unity = bytes(575) + b"\x01"
                    line["secondary_bbs_plus_verify"]["cpu_baseline"] = {
                        "value": mS / dtc, "unit": "signatures/s", "cores": th, "kind": "reference",
                        "sample": f"{mS} signatures: per signature the live G1 product loop (double_multiply), w * g2^x, pair_double_ate + final exponentiation "
                                  "through the unmodified bridge (point parsing excluded; the two affine conversions in between are not timed work of the reference)",
                        "all_accepted": all(gtc[576 * i:576 * i + 576] == unity for i in range(mS))}
            except Exception as e:
                line["secondary_bbs_plus_verify"]["cpu_baseline"] = {"value": None, "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
