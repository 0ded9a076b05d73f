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
# roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant (fused) kernel at the
# full single-GPU workload, taken from the committed `ncu --set full` captures (bytes cannot be counted inside a timed
# run); roofline.traffic_source names the file.  Absent capture => null.
TRAFFIC_NCU = {"C3": (25.602092e9 + 0.401808e9, "profiles/r02_ncu_full_c3.txt"),
               "C2": (5.121371e9 + 0.042751e9, "profiles/r01_ncu_full_c2.txt"),
               # C4 streams the X K-blocks once per centroid tile (16 tiles): 820 GB reach the SMs, L2 serves 78 % of it
               "C4": (368.469337e9 + 0.216847e9, "profiles/r02_ncu_full_c4.txt"),
               "C5": (12.800274e9 + 0.746665e9, "profiles/r02_ncu_full_c5.txt")}
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


GEN_CHUNK = 1 << 20   # rows per generator chunk (global row index // GEN_CHUNK keys the stream)


def blob_centres(torch, d, k, seed=1234):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand((k, d), device="cuda", generator=g) * 20.0 - 10.0


def gen_blobs_device(torch, lo, hi, d, k, seed=1234):
    """rows [lo, hi) of the synthetic matrix: isotropic blobs, centres ~ U(-10,10)^d, sigma 1.
    The generator is keyed by the GLOBAL row index (chunk c = rows [c*2^20, (c+1)*2^20) has its own seeded stream), so
    every sharding of the rows -- N = 1, 2, 4, 8 ranks -- sees the same matrix."""
    centres = blob_centres(torch, d, k, seed)
    X = torch.empty((hi - lo, d), dtype=torch.float32, device="cuda")
    gs = torch.Generator(device="cuda")
    for c in range(lo // GEN_CHUNK, (max(hi, lo + 1) - 1) // GEN_CHUNK + 1):
        gs.manual_seed(seed * 7919 + 1 + c)
        lab = torch.randint(0, k, (GEN_CHUNK,), device="cuda", generator=gs)
        blk = centres[lab]
        blk += torch.randn((GEN_CHUNK, d), device="cuda", generator=gs)
        g0 = c * GEN_CHUNK
        a, b = max(lo, g0), min(hi, g0 + GEN_CHUNK)
        if b > a:
            X[a - lo:b - lo] = blk[a - g0:b - g0]
        del blk, lab
    return X, centres


def throughput_init(torch, n, d, k, seed=42):
    """k distinct data rows among the first min(n, 2^20) global rows: every rank regenerates chunk 0 and picks the
    same rows (a poor start: every implementation runs all the iterations; not for centroid parity)"""
    head, _ = gen_blobs_device(torch, 0, min(n, GEN_CHUNK), d, k)
    g = torch.Generator(device="cuda").manual_seed(seed)
    return head[torch.randperm(head.shape[0], device="cuda", generator=g)[:k]].clone()


def parity_init(torch, centres, seed=42, jitter=0.5):
    """true centres + N(0, 0.5^2): one centroid per blob, unique stable fixed point (SURVEY 8c regime 1)"""
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    return (centres + jitter * torch.randn(centres.shape, device="cuda", generator=g)).contiguous()


def all_host_threads():
    """every host core for the CPU arm at every N: torchrun exports OMP_NUM_THREADS=1 to its workers, which made round
    1's N > 1 reference lines single-threaded.  Must run before numpy / scikit-learn are imported."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[v] = str(n)
    return n


def host_blobs(n_rows, d, k):
    """the first n_rows rows of the synthetic matrix on the HOST (numpy) + the throughput init.  On a GPU box they come
    from the same device generator as our arm (identical blobs); without a GPU (build container) from numpy."""
    import numpy as np
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        X = np.empty((n_rows, d), dtype=np.float32)
        step = 8 * GEN_CHUNK
        for lo in range(0, n_rows, step):
            hi = min(n_rows, lo + step)
            Xd, _ = gen_blobs_device(torch, lo, hi, d, k)
            X[lo:hi] = Xd.cpu().numpy()
            del Xd
        init = throughput_init(torch, n_rows, d, k).cpu().numpy()
        torch.cuda.empty_cache()
        return X, init, "device generator (same rows as the GPU arm)"
    rng = np.random.default_rng(1234)
    centres = rng.uniform(-10, 10, size=(k, d)).astype(np.float32)
    X = np.empty((n_rows, d), dtype=np.float32)
    for lo in range(0, n_rows, GEN_CHUNK):
        hi = min(n_rows, lo + GEN_CHUNK)
        X[lo:hi] = centres[rng.integers(0, k, size=hi - lo)] + rng.standard_normal((hi - lo, d), dtype=np.float32)
    init = X[rng.choice(min(n_rows, GEN_CHUNK), size=k, replace=False)].copy()
    return X, init, "numpy generator (no GPU visible)"


def cpu_reference_rate(n_full, d, k, budget_s=90.0, warm_iters=1, timed_iters=2, max_timed=2):
    """reference CPU execution path (sklearn KMeans, cuML's _cpu_class_path, kmeans.pyx:604) with all host cores.
    Marginal Lloyd-iteration rate: t(fit, max_iter = warm + timed) - t(fit, max_iter = warm), so sklearn's fixed work
    (validation, centring, the final labelling pass) is not charged to the iterations.  Runs on ALL n_full rows when a
    short calibration says that fits `budget_s` and host memory; otherwise on the largest row prefix that does, scaled
    by the row ratio and labelled an estimate.  Returns (full-workload iter/s, cores, sample text, rows, timed iters)."""
    import numpy as np
    from threadpoolctl import threadpool_limits
    from oracle import sklearn_ref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with threadpool_limits(limits=cores):
        cores_eff = min(cores, sklearn_ref.n_threads())
        # calibration: seconds per row and iteration on a small prefix
        n_cal = int(min(n_full, max(20 * k, 200_000)))
        Xc, initc, _ = host_blobs(n_cal, d, k)
        sklearn_ref.time_fit(Xc, initc, max_iter=1, reps=1)
        t1, _ = sklearn_ref.time_fit(Xc, initc, max_iter=1, reps=1)
        t3, _ = sklearn_ref.time_fit(Xc, initc, max_iter=3, reps=1)
        per_row_iter = max((t3 - t1) / 2.0, 1e-9) / n_cal
        per_row_fixed = max(t1 - (t3 - t1) / 2.0, 0.0) / n_cal
        del Xc
        # two fits: warm and warm + timed iterations, each with its fixed part
        cost_per_row = per_row_iter * (2 * warm_iters + timed_iters) + 2 * per_row_fixed
        rows = int(min(n_full, max(20 * k, budget_s / max(cost_per_row, 1e-12))))
        try:
            import psutil
            avail = psutil.virtual_memory().available
            rows = int(min(rows, max(20 * k, 0.4 * avail / (4.0 * d))))   # X + sklearn's working copies
        except Exception:
            pass
        if rows >= 0.97 * n_full:
            rows = n_full
        # cheap workloads: time more iterations (up to the requested steps) so the difference of two fits is not noise
        timed_iters = int(max(timed_iters, min(max_timed, 4.0 / max(per_row_iter * rows, 1e-9))))
        X, init, how = host_blobs(rows, d, k)
        big = X.nbytes > (2 << 30)      # no second host copy of a multi-GB matrix (sklearn then centres X in place)
        ta, na = sklearn_ref.time_fit(X, init, max_iter=warm_iters, reps=1, copy_x=not big)
        tb, nb = sklearn_ref.time_fit(X, init, max_iter=warm_iters + timed_iters, reps=1, copy_x=not big)
    if nb > na and tb > ta:
        rate_rows, timed = (nb - na) / (tb - ta), nb - na
    else:
        rate_rows, timed = nb / tb, nb
    rate_full = rate_rows * (rows / n_full)
    import sklearn
    sample = (f"sklearn {sklearn.__version__} KMeans(init=array (k data rows), lloyd, tol=0, n_init=1), {cores_eff} threads, on "
              f"{rows} of {n_full} rows ({how}); marginal rate (t[max_iter={warm_iters + timed_iters}] - "
              f"t[max_iter={warm_iters}] = {tb - ta:.2f} s for {timed} iterations, one fit each)"
              + ("" if rows == n_full else "; ESTIMATED: full-size rate = sample rate x rows ratio"))
    return rate_full, cores_eff, sample, rows, timed


def cpu_reference_predict_rate(n_full, d, k, budget_s=30.0):
    """reference CPU path's predict on a bounded row prefix -> (full-workload passes/s, cores, sample, rows)"""
    import numpy as np
    from threadpoolctl import threadpool_limits
    from oracle import sklearn_ref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with threadpool_limits(limits=cores):
        cores_eff = min(cores, sklearn_ref.n_threads())
        n_cal = int(min(n_full, 50_000))
        Xc, _, _ = host_blobs(n_cal, d, k)
        centres = Xc[:k].copy()
        tc = sklearn_ref.time_predict(Xc, centres, reps=2)
        rows = int(min(n_full, max(4 * k, (budget_s / 3.0) / max(tc / n_cal, 1e-12))))
        X, _, how = host_blobs(rows, d, k)
        t = sklearn_ref.time_predict(X, centres, reps=2)
    import sklearn
    sample = (f"sklearn {sklearn.__version__} KMeans.predict, {cores_eff} threads, on {rows} of {n_full} rows ({how}), best of 2"
              + ("" if rows == n_full else "; ESTIMATED: full-size rate = sample rate x rows ratio"))
    return (1.0 / t) * (rows / n_full), cores_eff, sample, rows


def workload_config(workload, n, d, k):
    """the `config` object: identical in both arms (it names the workload, not how an arm runs it)"""
    return {"workload": f"{workload}: {WORKLOADS[workload][3]}", "n": n, "d": d, "k": k, "l2": "inputs_exceed_l2",
            "init": "array (k data rows)", "data_seed": 1234}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (sklearn, kmeans.pyx:604) on this box's host
    cores.  Under torchrun only rank 0 works; the line reports the steps it really ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, k, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    metric, unit = metric_unit(args.workload)
    if args.workload == "C4":
        rate, cores, sample, rows = cpu_reference_predict_rate(n, d, k, budget_s=args.cpu_budget or 60.0)
        warm, timed = 0, 2
    else:
        rate, cores, sample, rows, timed = cpu_reference_rate(n, d, k, budget_s=args.cpu_budget or 110.0,
                                                              max_timed=max(2, args.steps))
        warm = 1
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus,
        "steps": timed, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n, d, k),
        "dists_per_sec": rate * n * k,
        "estimated": rows != n,
        "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def parity_fit(torch, dist, _lib, lib, h, args, n, d, k, lo, hi, world, rank, stream, barrier):
    """The multi-GPU contract of the reference (cpp/tests/mg/kmeans_test.cu:116-137, test_dask_kmeans.py:54-126): the
    model fitted on N row shards equals the one fitted on one GPU.  All ranks fit their shards of the SAME matrix from
    the parity init (one centroid per blob: stable fixed point) through the C-ABI; at N > 1 rank 0 then regenerates the
    whole matrix and repeats the fit alone on a communicator-less handle.  Bars: inertia 1e-5, centroids 1e-4."""
    iters = max(1, min(args.steps, 10))
    Xs, centres = gen_blobs_device(torch, lo, hi, d, k)
    Cp = parity_init(torch, centres)

    def fit(handle, Xt, rows):
        p = _lib.default_params()
        p.n_clusters, p.init, p.max_iter, p.tol, p.n_init = k, _lib.INIT_ARRAY, iters, 0.0, 1
        Cf = Cp.clone()
        inertia, n_iter = C.c_float(), C.c_int64()
        xp = (C.c_void_p * 1)(Xt.data_ptr())
        rows_a = (C.c_int64 * 1)(rows)
        _lib.check(lib.cuml_b200_kmeans_fit_parts_f32(handle.ptr, C.byref(p), xp, rows_a, 1, d, None, Cf.data_ptr(),
                                                      C.byref(inertia), C.byref(n_iter)))
        torch.cuda.synchronize()
        return Cf, float(inertia.value)

    C_n, in_n = fit(h, Xs, hi - lo)
    out = {"init": "true centres + N(0, 0.5^2)", "iters": iters, "inertia": in_n,
           "centroid_abs_sum": float(C_n.double().abs().sum().item())}
    del Xs
    torch.cuda.empty_cache()
    if world > 1:
        if rank == 0:
            h1 = _lib.Handle(stream=stream.cuda_stream)
            Xf, _ = gen_blobs_device(torch, 0, n, d, k)
            C_1, in_1 = fit(h1, Xf, n)
            h1.close()
            del Xf
            torch.cuda.empty_cache()
            cerr = float(((C_n - C_1).abs().max() / C_1.abs().max()).item())
            irel = abs(in_n - in_1) / abs(in_1)
            out["vs_n1"] = {"inertia_n1": in_1, "inertia_rel": irel, "centroid_err_max_over_max": cerr,
                            "ok": bool(irel <= 1e-5 and cerr <= 1e-4),
                            "how": "rank 0 refits the whole matrix alone after the timed regions"}
        barrier()
    return out


def bind_to_gpu_numa_node(torch, local_rank):
    """pin this rank's host threads (and therefore the pages of the pinned buffers it allocates next) to the CPUs NVML
    reports as local to its GPU.  Round 1's 8-rank end-to-end runs copied at 21 GB/s per rank against 53 GB/s alone: all
    eight pinned buffers had been allocated from wherever the launcher happened to run."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        hdl = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {bus} ({cpus[0]}-{cpus[-1]})"
    except Exception as e:   # NVML absent or affinity not permitted: placement stays as launched
        return f"not bound ({type(e).__name__})"
    return "not bound"


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
    # pinned host buffers should come from the GPU's own NUMA node at every N (a single-rank run that landed on the far
    # node measured 31 instead of 55 GB/s of H2D: e2e 16 instead of 24.5 iter/s); the CPU baseline gets every core back
    full_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local_rank)
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
        h = comms_from_torch_distributed(stream=stream.cuda_stream, backend=args.comm)
    else:
        h = _lib.Handle(stream=stream.cuda_stream)

    lo, hi = shard_bounds(n, rank, world)
    n_local = hi - lo
    X, centres = gen_blobs_device(torch, lo, hi, d, k)
    C0 = throughput_init(torch, n, d, k)     # identical on all ranks by construction
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
        # the first seeded call also pays the one-off device allocations of the seeding work buffers (per-row minimum
        # distances, candidate lists: hundreds of MB at this size, from a cold stream-ordered pool); the reference draws
        # them from a warm RMM pool, so the steady call is the number to compare -- both are reported
        t_seed_first, _ = fit0(_lib.INIT_KMEANS_PLUS_PLUS)
        t_seed, inertia_seed = fit0(_lib.INIT_KMEANS_PLUS_PLUS)
        tt = torch.tensor([t_arr, t_seed, t_seed_first], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_arr, t_seed, t_seed_first = [float(v) for v in tt.tolist()]
        init_info = {"method": "k-means|| (oversampling_factor 2.0)", "seconds": max(0.0, t_seed - t_arr),
                     "first_call_seconds": max(0.0, t_seed_first - t_arr),
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

    parity = None
    if not predict_only and not args.no_parity:
        if "X" in dir():
            del X
        torch.cuda.empty_cache()
        parity = parity_fit(torch, dist, _lib, lib, h, args, n, d, k, lo, hi, world, rank, stream, barrier)

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
        fused_upd = (not predict_only) and bool(int(lib.cuml_b200_kmeans_fused_update(h.ptr, d, k))) and n_local >= 2
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
        # second yardstick for the fused kernel: the north-star's 3xTF32 rule (three tf32 MMAs per product:
        # bf16_sustained / 6) whatever scheme the kernel runs, next to the scheme's own peak above
        tf_peak_3x = peaks["bf16_tflops_sustained"] / 6.0
        # the whole iteration against its one-pass floor: max(tensor time at the scheme's peak, X read once + labels
        # written at the HBM peak); this is the fraction the north-star target (>= 0.70) is quoted on
        floor_ms = max(flop / (tf_peak * 1e12), fused_bytes / (peaks["hbm_gbs"] * 1e9)) * 1e3
        roof.update({
            "frac_vs_3xtf32_peak": (tf_achieved / tf_peak_3x) if (tf_achieved and tensor_bound) else None,
            "iteration": {"ms": ms_per_step, "one_pass_floor_ms": floor_ms, "frac": floor_ms / ms_per_step,
                          "note": "floor = max(2nkd / tensor peak, (4nd + 4n) / hbm peak) for this rank's rows"},
            "traffic": TRAFFIC_NCU[args.workload][0] if (args.workload in TRAFFIC_NCU and world == 1 and not args.n) else None,
            "traffic_source": (TRAFFIC_NCU[args.workload][1] + " (ncu --set full, one launch; not measured in this run)")
                              if (args.workload in TRAFFIC_NCU and world == 1 and not args.n) else None,
            "kernel": f"fused_l2_argmin ({variant_name}: distance + argmin)", "kernel_ms": fused_ms,
            "algorithmic_flops_per_launch": flop, "algorithmic_bytes_per_launch": fused_bytes,
            "algorithmic_tflops": tf_achieved,
            "issued_mma_products_per_algorithmic": 3,
            "hbm_gbs_fused": gb_achieved,
            "peak_note": f"{peaks['source']}: tensor peak = bf16_tflops_sustained / {mma_cost:g} "
                         f"({'1 tf32 + 2 bf16 MMAs' if variant in (3, 5) else '3 tf32 MMAs'} per algorithmic product, "
                         f"tf32 = bf16/2); hbm peak = hbm_gbs (copy)",
            "update_kernel": ("none: centroid sums / counts come out of the fused kernel (one pass over X); time below = reduce_partials"
                              if fused_upd else
                              "accumulate (TMA ring; label-class ownership or lane = column tables) + reduce_partials"),
            "update_kernel_ms": update_ms,
            "update_kernel_hbm_gbs": 4.0 * n_local * (d + 1) / (update_ms * 1e-3) / 1e9 if (update_ms > 0 and not fused_upd) else None,
            "update_kernel_frac_of_hbm_peak": (4.0 * n_local * (d + 1) / (update_ms * 1e-3) / 1e9 / peaks["hbm_gbs"])
                                              if (update_ms > 0 and not fused_upd) else None,
            "hbm_peak_gbs": peaks["hbm_gbs"], "tensor_peak_tflops": tf_peak})
        metric, unit = metric_unit(args.workload)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (split-precision tensor-core contraction: tf32 + bf16 corrections or 3xTF32; fp32/fp64 reductions)",
            "data": "synthetic",
            "config": workload_config(args.workload, n, d, k),
            "run": {"rows_per_gpu": n_local, "engine": args.engine, "communicator": getattr(h, "comm_kind", None),
                    "host_threads": numa, "e2e_host_memory": "pinned (cudaHostAlloc through torch)",
                    "parallelism": (f"row-sharded x{world}, rank-local predict (no collective)" if predict_only else
                                    f"row-sharded x{world}, 1 all-reduce of (k*d+k+1) f64 per iteration")},
            "dists_per_sec": value * n * k,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": roof,
        }
        if init_info is not None:
            line["init"] = init_info
        if parity is not None:
            line["parity_fit"] = parity
        if world == 1 and not args.no_cpu:
            try:
                os.sched_setaffinity(0, full_affinity)   # the CPU baseline runs on every core again
            except OSError:
                pass
            if predict_only:
                rate, cores, sample, _ = cpu_reference_predict_rate(n, d, k, budget_s=args.cpu_budget or 20.0)
            else:
                rate, cores, sample, _, _ = cpu_reference_rate(n, d, k, budget_s=args.cpu_budget or 25.0)
            line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": "reference", "sample": sample}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity and not parity.get("vs_n1", {}).get("ok", True):
        print("bench.py: the N-rank fit does not match the 1-rank fit: %r" % (parity["vs_n1"],), file=sys.stderr)
        sys.exit(3)


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
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank vs 1-rank parity fit")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="seconds of CPU work for the sklearn arm (0 = default)")
    ap.add_argument("--comm", default=None, choices=["nccl", "peer"], help="communicator of the N > 1 runs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        all_host_threads()
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
