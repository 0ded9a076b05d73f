#!/usr/bin/env python
"""bench.py -- Lloyd iterations/s (and sample-centroid distances/s) of the k-means hot path.

    python bench.py --gpus N --steps K --warmup W [--workload C3|C2|C1|C4|C5] [--impl ours|reference]

Contract (see the task statement): W untimed warm-up steps, then exactly K Lloyd iterations timed
with CUDA events between barriers, max over ranks, ONE JSON line from rank 0.

Workloads are the BASELINE.json configs; the default (and what the driver runs at every N) is
C3 = "KMeans fit n=100M d=64 k=256 fp32 row-sharded across 1/2/4/8 B200" -- the config the
metric and the north-star target are quoted on; it fits one B200 (25.6 GB), total work is fixed
as N grows ("scaling": "strong").  Inputs are ~200x larger than L2, so no explicit L2 flush.
C4 is the inference config: a "step" is one ML::kmeans::predict call (labels + inertia) over the rank's rows,
`value` = predict passes/s.  C5 additionally reports the k-means|| seeding time (`init` object).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, k, description)
    "C1": (1_000_000, 32, 16, "KMeans fit n=1M d=32 k=16 fp32 init=array max_iter=50 tol=0"),
    "C2": (10_000_000, 128, 1024, "KMeans fit n=10M d=128 k=1024 fp32"),
    "C3": (100_000_000, 64, 256, "KMeans fit n=100M d=64 k=256 fp32 row-sharded"),
    "C4": (50_000_000, 256, 4096, "KMeans predict n=50M d=256 k=4096 fp32 (fused distance+argmin inference)"),
    "C5": (200_000_000, 16, 64, "KMeans fit n=200M d=16 k=64 fp32 row-sharded"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant (fused) kernel at the full
# workload, from the committed `ncu --set full` captures (profiles/r01_ncu_full_c3.txt, r01_ncu_full_c2.txt)
TRAFFIC_NCU = {"C3": 25.602643e9 + 0.401213e9, "C2": 5.121371e9 + 0.042751e9}
METRIC = "kmeans_lloyd_iters_per_sec"
UNIT = "Lloyd iter/s"


def metric_unit(workload):
    # C4 is inference: one step = one predict pass over the rows
    return ("kmeans_predict_passes_per_sec", "predict passes/s") if workload == "C4" else (METRIC, UNIT)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        """block until nvidia-smi has printed its first sample (its start-up takes 0.1-1 s on a multi-GPU box, longer
        than a short timed region), so that the polling really is running when the timed region starts"""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout and self.proc.poll() is None:
            time.sleep(0.01)

    def inside(self, t_begin, t_end):
        """number of samples that arrived inside [t_begin, t_end] (host clock)"""
        return len([1 for (ts, _) in list(self.lines) if t_begin - 0.01 <= ts <= t_end + 0.03])

    def stop(self, t_begin=None, t_end=None):
        """clocks of the samples that arrived inside [t_begin, t_end] (host clock; all samples when not given)"""
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        time.sleep(0.15)
        self.proc.terminate()
        return self.stats(list(self.lines), t_begin, t_end)

    @staticmethod
    def stats(lines, t_begin=None, t_end=None):
        sm, mx, reasons = [], [], set()
        inside = [ln for (ts, ln) in lines
                  if t_begin is None or (t_begin - 0.01 <= ts <= t_end + 0.03)]
        if not inside:   # timed region shorter than one polling period: nearest samples around it
            inside = [ln for (ts, ln) in lines][-3:]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def gen_blobs_device(torch, n_local, d, k, row_offset, seed=1234, chunk=1 << 22):
    """isotropic blobs, centres ~ U(-10,10)^d, sigma 1, generated shard-locally on device"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    centres = torch.rand((k, d), device="cuda", generator=g) * 20.0 - 10.0   # same on every rank
    X = torch.empty((n_local, d), dtype=torch.float32, device="cuda")
    gs = torch.Generator(device="cuda").manual_seed(seed * 7919 + 1 + row_offset % (2**31))
    for s in range(0, n_local, chunk):
        e = min(n_local, s + chunk)
        lab = torch.randint(0, k, (e - s,), device="cuda", generator=gs)
        X[s:e] = centres[lab]
        X[s:e] += torch.randn((e - s, d), device="cuda", generator=gs)
    return X, centres


def cpu_reference_rate(n_full, d, k, budget_rows=None, iters=3):
    """reference CPU execution path (sklearn KMeans, cuML's _cpu_class_path) on a bounded sample of
    the workload; returns (full-workload iter/s, cores, sample description)."""
    import numpy as np
    from oracle import sklearn_ref
    cores = sklearn_ref.n_threads()
    # ~10-30 s of CPU work: n_s * k * d * 2 flop * iters at ~5 GFLOP/s/core
    if budget_rows is None:
        target_flop = 15.0 * 5e9 * max(cores, 1)
        budget_rows = int(max(50_000, min(n_full, target_flop / (2.0 * k * d * (iters + 1)))))
    rng = np.random.default_rng(1234)
    centres = rng.uniform(-10, 10, size=(k, d)).astype(np.float32)
    lab = rng.integers(0, k, size=budget_rows)
    X = centres[lab] + rng.standard_normal((budget_rows, d), dtype=np.float32)
    init = X[rng.choice(budget_rows, size=k, replace=False)].copy()   # throughput init: runs all iters
    # marginal cost per Lloyd iteration: t(1+iters) - t(1), so sklearn's fixed overhead (validation,
    # mean-centring, final E-step) is not charged to the iterations
    t1, n1 = sklearn_ref.time_fit(X, init, max_iter=1, reps=2)
    t2, n2 = sklearn_ref.time_fit(X, init, max_iter=1 + iters, reps=2)
    if n2 > n1 and t2 > t1:
        rate_sample = (n2 - n1) / (t2 - t1)
    else:
        rate_sample = n2 / t2
    rate_full = rate_sample * (budget_rows / n_full)
    sample = (f"sklearn {__import__('sklearn').__version__} KMeans(init=array, lloyd, tol=0) on "
              f"{budget_rows} of {n_full} rows (same d={d}, k={k}); marginal rate (t[{1 + iters} iters]-t[1 iter]), "
              f"best of 2 each; full-size rate = sample rate x rows ratio")
    return rate_full, cores, sample


def cpu_reference_predict_rate(n_full, d, k):
    """reference CPU path's predict on a bounded row sample -> (full-workload passes/s, cores, sample)"""
    import numpy as np
    from oracle import sklearn_ref
    cores = sklearn_ref.n_threads()
    target_flop = 5.0 * 5e9 * max(cores, 1)                    # ~5 s per repetition, 3 repetitions
    rows = int(max(4 * k, min(n_full, target_flop / (2.0 * k * d))))
    rng = np.random.default_rng(1234)
    centres = rng.uniform(-10, 10, size=(k, d)).astype(np.float32)
    X = centres[rng.integers(0, k, size=rows)] + rng.standard_normal((rows, d), dtype=np.float32)
    t = sklearn_ref.time_predict(X, centres, reps=3)
    sample = (f"sklearn {__import__('sklearn').__version__} KMeans.predict on {rows} of {n_full} rows (same d={d}, "
              f"k={k}), best of 3; full-size rate = sample rate x rows ratio")
    return (1.0 / t) * (rows / n_full), cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, k, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    rates = []
    for _ in range(max(1, min(args.steps, 3))):
        if args.workload == "C4":
            rate, cores, sample = cpu_reference_predict_rate(n, d, k)
        else:
            rate, cores, sample = cpu_reference_rate(n, d, k, iters=3)
        rates.append(rate)
    rate = max(rates)
    metric, unit = metric_unit(args.workload)
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "n": n, "d": d, "k": k},
        "dists_per_sec": rate * n * k,
        "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cuml_b200 import _lib
    from cuml_b200.cluster.kmeans_mg import comms_from_torch_distributed, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, d, k, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    lib = _lib.load()
    # a non-default stream shared by torch (events, data generation) and the library handle, so the
    # CUDA events below bracket exactly the library's launches
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream()
    assert stream.cuda_stream != 0
    if world > 1:
        h = comms_from_torch_distributed(stream=stream.cuda_stream)
    else:
        h = _lib.Handle(stream=stream.cuda_stream)

    lo, hi = shard_bounds(n, rank, world)
    n_local = hi - lo
    X, centres = gen_blobs_device(torch, n_local, d, k, lo)
    # throughput init = k data rows of rank 0's shard, identical on all ranks
    g = torch.Generator(device="cuda").manual_seed(42)
    C0 = X[torch.randperm(min(n_local, 1 << 20), device="cuda", generator=g)[:k]].clone()
    if world > 1:
        dist.broadcast(C0, src=0)
    Cd = C0.clone()
    labels = torch.zeros(n_local, dtype=torch.int32, device="cuda")
    engine = {"auto": 0, "simt": 1, "tc": 2}[args.engine]

    predict_only = args.workload == "C4"
    if predict_only:
        # ML::kmeans::predict through the C-ABI, int64 index overload (n*d > INT_MAX, reference kmeans.pyx:277-281)
        pp = _lib.default_params()
        pp.n_clusters = k
        labels = torch.zeros(n_local, dtype=torch.int64, device="cuda")
        pred_inertia = C.c_float()

    def step():
        if predict_only:
            _lib.check(lib.cuml_b200_kmeans_predict_f32_i64(h.ptr, C.byref(pp), Cd.data_ptr(), X.data_ptr(), n_local, d,
                                                            None, 1, labels.data_ptr(), C.byref(pred_inertia)))
            return
        _lib.check(lib.cuml_b200_kmeans_lloyd_step_f32(h.ptr, X.data_ptr(), n_local, d, None, k, Cd.data_ptr(),
                                                       None, None, None, engine))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # polling (20 ms) is already running when the timed region starts
        sampler.wait_first()
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    _lib.check(lib.cuml_b200_kernel_timing_enable(h.ptr, 1))
    lib.cuml_b200_launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_end = time.time()
    ms_total = ev0.elapsed_time(ev1)
    clock_window = (t_begin, t_end)
    launches = int(lib.cuml_b200_launch_count())
    f_ms, f_n, u_ms, u_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    _lib.check(lib.cuml_b200_kernel_timing_read(h.ptr, C.byref(f_ms), C.byref(f_n), C.byref(u_ms), C.byref(u_n)))
    _lib.check(lib.cuml_b200_kernel_timing_enable(h.ptr, 0))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # clocks / throttle reasons are those sampled DURING the timed region.  When that region was shorter than two
    # polling periods (multi-GPU runs: tens of ms) the same steps are repeated, untimed, for ~0.4 s right after it and
    # the samples of that identical load are reported instead (`clocks.window` says which).  Every rank repeats the same
    # number of steps (the step holds a collective); kernel timing and launch counting are already closed.
    probe = torch.tensor([1 if (rank == 0 and sampler.proc and sampler.inside(*clock_window) < 2) else 0],
                         dtype=torch.int32, device="cuda")
    if world > 1:
        dist.broadcast(probe, src=0)
    clock_note = "timed region"
    if int(probe.item()):
        extra = int(min(5000, max(args.steps, 400.0 / max(ms_per_step, 1e-3))))
        barrier()
        t_p0 = time.time()
        for _ in range(extra):
            step()
        barrier()
        clock_window = (t_p0, time.time())
        clock_note = f"{extra} untimed repetitions of the step right after the timed region (it was too short to sample)"
    clocks = None
    if rank == 0:
        clocks = sampler.stop(*clock_window)
        clocks["window"] = clock_note

    # ---- C5: k-means|| seeding time = fit(init=k-means||, max_iter=0) - fit(init=Array, max_iter=0), device X
    # (both calls end with the same final E-step + inertia pass; SURVEY 8d "report init time separately")
    init_info = None
    if args.workload == "C5":
        def fit0(init_method):
            p0 = _lib.default_params()
            p0.n_clusters, p0.init, p0.max_iter, p0.tol, p0.n_init = k, init_method, 0, 0.0, 1
            p0.oversampling_factor, p0.rng_seed = 2.0, 42
            Cs = C0.clone()
            inertia0, it0 = C.c_float(), C.c_int64()
            xp0 = (C.c_void_p * 1)(X.data_ptr())
            rows0 = (C.c_int64 * 1)(n_local)
            barrier()
            t0 = time.perf_counter()
            _lib.check(lib.cuml_b200_kmeans_fit_parts_f32(h.ptr, C.byref(p0), xp0, rows0, 1, d, None, Cs.data_ptr(),
                                                          C.byref(inertia0), C.byref(it0)))
            torch.cuda.synchronize()
            return time.perf_counter() - t0, float(inertia0.value)
        fit0(_lib.INIT_ARRAY)                                   # warm-up of the final pass
        t_arr, _ = fit0(_lib.INIT_ARRAY)
        t_seed, inertia_seed = fit0(_lib.INIT_KMEANS_PLUS_PLUS)
        tt = torch.tensor([t_arr, t_seed], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_arr, t_seed = [float(v) for v in tt.tolist()]
        init_info = {"method": "k-means|| (oversampling_factor 2.0)", "seconds": max(0.0, t_seed - t_arr),
                     "fit_max_iter0_seconds": t_seed, "final_pass_seconds": t_arr, "inertia_after_init": inertia_seed}

    # ---- end-to-end through the C-ABI fit with HOST buffers (H2D + K iterations + final predict pass)
    e2e = None
    if predict_only and not args.no_e2e:
        # predict with HOST rows: chunks of pinned host memory copied to a device buffer and predicted one by one
        # (the estimator's chunked host predict, reference _kmeans_predict_host_chunked kmeans.pyx:356-434)
        chunk = min(n_local, 4_000_000)
        Xh = torch.empty((chunk, d), dtype=torch.float32, pin_memory=True)
        Xh.copy_(X[:chunk])
        xb = torch.empty((chunk, d), dtype=torch.float32, device="cuda")
        lb = torch.zeros(chunk, dtype=torch.int64, device="cuda")
        lab_host = torch.empty(chunk, dtype=torch.int64, pin_memory=True)
        n_chunks = (n_local + chunk - 1) // chunk
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_chunks):                             # every chunk re-sends the same pinned rows
            xb.copy_(Xh, non_blocking=True)
            _lib.check(lib.cuml_b200_kmeans_predict_f32_i64(h.ptr, C.byref(pp), Cd.data_ptr(), xb.data_ptr(), chunk, d,
                                                            None, 1, lb.data_ptr(), C.byref(pred_inertia)))
            lab_host.copy_(lb, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": (n_chunks * chunk / n_local) / dt, "unit": "predict passes/s",
               "h2d_bytes_per_step": int(n_chunks * chunk * d * 4), "d2h_bytes_per_step": int(n_chunks * chunk * 8),
               "seconds": dt, "call": f"cuml_b200_kmeans_predict_f32_i64 on {n_chunks} host chunks of {chunk} rows"}
        del Xh, xb, lb
    elif not args.no_e2e:
        Xh = torch.empty((n_local, d), dtype=torch.float32, pin_memory=True)
        Xh.copy_(X)
        del X, labels
        torch.cuda.empty_cache()
        p = _lib.default_params()
        p.n_clusters, p.init, p.max_iter, p.tol = k, _lib.INIT_ARRAY, args.steps, 0.0
        Ce = C0.clone()
        inertia, n_iter = C.c_float(), C.c_int64()
        xp = (C.c_void_p * 1)(Xh.data_ptr())
        rows = (C.c_int64 * 1)(n_local)
        barrier()
        t0 = time.perf_counter()
        _lib.check(lib.cuml_b200_kmeans_fit_parts_f32(h.ptr, C.byref(p), xp, rows, 1, d, None, Ce.data_ptr(),
                                                      C.byref(inertia), C.byref(n_iter)))
        centers_host = Ce.cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": n_iter.value / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * d * 4 / max(1, n_iter.value)),
               "d2h_bytes_per_step": int((k * d * 4 + 8) / max(1, n_iter.value)),
               "seconds": dt, "call": "cuml_b200_kmeans_fit_parts_f32(host X, init=Array, max_iter=steps, tol=0)",
               "inertia": float(inertia.value)}
        del Xh

    if rank == 0:
        peaks = measured_peaks()
        fused_ms = f_ms.value / max(1, f_n.value)
        update_ms = u_ms.value / max(1, u_n.value)
        flop = 2.0 * n_local * k * d
        fused_bytes = 4.0 * n_local * d + 4.0 * n_local            # X read once + labels written
        tf_achieved = flop / (fused_ms * 1e-3) / 1e12 if fused_ms > 0 else None
        gb_achieved = fused_bytes / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else None
        # fp32-equivalent tensor roofline from the measured dense bf16 rate (MEASURED_PEAKS.json).  TF32 runs at half
        # the bf16 rate.  3xTF32 issues 3 tf32 MMAs per algorithmic product: peak(2nkd) = bf16/2/3.  The CTA-pair
        # kernel with bf16 correction terms issues 1 tf32 + 2 bf16 MMAs: time per product = 1/(bf16/2) + 2/bf16, so
        # peak(2nkd) = bf16/4 (the denominator is larger, the fraction therefore lower, than under the 3xTF32 rule).
        variant = int(lib.cuml_b200_kmeans_estep_variant(h.ptr, d, k))
        mma_cost = 4.0 if variant in (3, 5) else 6.0
        tf_peak = peaks["bf16_tflops_sustained"] / mma_cost
        # which roofline binds the fused kernel: time at the tensor peak vs time at the HBM peak
        tensor_bound = (flop / (tf_peak * 1e12)) >= (fused_bytes / (peaks["hbm_gbs"] * 1e9))
        if tensor_bound:
            roof = {"bound": "tensor", "achieved": tf_achieved, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": (tf_achieved / tf_peak) if tf_achieved else None}
        else:
            roof = {"bound": "hbm", "achieved": gb_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (gb_achieved / peaks["hbm_gbs"]) if gb_achieved else None}
        variant_name = {0: "CUDA-core fp32", 1: "tcgen05 1-CTA 3xTF32", 2: "tcgen05 CTA-pair 3xTF32",
                        3: "tcgen05 CTA-pair tf32 + 2 bf16 correction terms", 4: "tcgen05 A-in-TMEM 3xTF32",
                        5: "tcgen05 1-CTA tf32 + 2 bf16 correction terms"}[variant]
        roof.update({
            "traffic": TRAFFIC_NCU.get(args.workload) if world == 1 and not args.n else None,
            "kernel": f"fused_l2_argmin ({variant_name}: distance + argmin)", "kernel_ms": fused_ms,
            "algorithmic_flops_per_launch": flop, "algorithmic_bytes_per_launch": fused_bytes,
            "algorithmic_tflops": tf_achieved,
            "issued_mma_products_per_algorithmic": 3,
            "hbm_gbs_fused": gb_achieved,
            "peak_note": f"{peaks['source']}: tensor peak = bf16_tflops_sustained / {mma_cost:g} "
                         f"({'1 tf32 + 2 bf16 MMAs' if variant in (3, 5) else '3 tf32 MMAs'} per algorithmic product, "
                         f"tf32 = bf16/2); hbm peak = hbm_gbs (copy)",
            "update_kernel": "accumulate_owner (TMA ring, label-class ownership) + reduce_partials",
            "update_kernel_ms": update_ms,
            "update_kernel_hbm_gbs": 4.0 * n_local * (d + 1) / (update_ms * 1e-3) / 1e9 if update_ms > 0 else None,
            "update_kernel_frac_of_hbm_peak": (4.0 * n_local * (d + 1) / (update_ms * 1e-3) / 1e9 / peaks["hbm_gbs"])
                                              if update_ms > 0 else None,
            "hbm_peak_gbs": peaks["hbm_gbs"], "tensor_peak_tflops": tf_peak})
        metric, unit = metric_unit(args.workload)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (split-precision tensor-core contraction: tf32 + bf16 corrections or 3xTF32; fp32/fp64 reductions)",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "n": n, "d": d, "k": k, "rows_per_gpu": n_local,
                       "parallelism": (f"row-sharded x{world}, rank-local predict (no collective)" if predict_only else
                                       f"row-sharded x{world}, 1 allreduce of (k*d+k+1) f64 per iteration"),
                       "l2": "inputs_exceed_l2", "engine": args.engine, "init": "array (k data rows)"},
            "dists_per_sec": value * n * k,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": roof,
        }
        if init_info is not None:
            line["init"] = init_info
        if world == 1 and not args.no_cpu:
            rate, cores, sample = (cpu_reference_predict_rate if predict_only else cpu_reference_rate)(n, d, k)
            line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": "reference", "sample": sample}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the row count (debugging only)")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
