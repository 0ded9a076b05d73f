"""Which role of the fused kernel limits a shape?  Times the E-step with parts disabled
(CUML_B200_DBG_SKIP bitmask; results are wrong by design when a part is skipped)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--run":
    sys.path.insert(0, ROOT)
    import torch, ctypes as C
    from cuml_b200 import _lib
    lib = _lib.load(); h = _lib.Handle()
    n, d, k = [int(x) for x in sys.argv[2:5]]
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((n, d), device="cuda", generator=g)
    Cd = torch.randn((k, d), device="cuda", generator=g)
    labels = torch.zeros(n, dtype=torch.int32, device="cuda")
    _lib.check(lib.cuml_b200_kernel_timing_enable(h.ptr, 1))
    for _ in range(6):
        _lib.check(lib.cuml_b200_kmeans_assign_f32(h.ptr, X.data_ptr(), n, d, k, Cd.data_ptr(), labels.data_ptr(), 2))
    f_ms, f_n, u_ms, u_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    _lib.check(lib.cuml_b200_kernel_timing_read(h.ptr, C.byref(f_ms), C.byref(f_n), C.byref(u_ms), C.byref(u_n)))
    print("RES", f_ms.value / f_n.value)
else:
    for shape in [(16000000, 16, 64), (8000000, 64, 256), (2000000, 128, 1024), (8000000, 32, 16)]:
        out = {}
        for skip in (0, 1, 2, 4, 3, 6, 5, 7):
            env = dict(os.environ, CUML_B200_DBG_SKIP=str(skip))
            r = subprocess.run([sys.executable, __file__, "--run"] + [str(x) for x in shape], capture_output=True, text=True, env=env, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
            out[skip] = round(float(line[0].split()[1]), 3) if line else r.stderr[-300:]
        print(shape, out, flush=True)
