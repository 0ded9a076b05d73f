"""Where does the end-to-end fit (host X) spend its time?  H2D rate, allocation, phases (CUML_B200_TRACE)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUML_B200_TRACE"] = "1"
import torch
from cuml_b200 import _lib
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000, 64, 256
lib = _lib.load()
torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream()
h = _lib.Handle(stream=st.cuda_stream)
Xh = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
for i in range(0, n, 1 << 22):
    Xh[i:i + (1 << 22)].normal_()
torch.cuda.synchronize()
# raw H2D rate into a fresh torch allocation
t0 = time.perf_counter(); Xd = torch.empty((n, d), dtype=torch.float32, device="cuda"); torch.cuda.synchronize(); t1 = time.perf_counter()
Xd.copy_(Xh, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
Xd.copy_(Xh, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
print(f"torch alloc {1e3*(t1-t0):.1f} ms, first H2D {1e3*(t2-t1):.1f} ms ({n*d*4/(t2-t1)/1e9:.1f} GB/s), second {1e3*(t3-t2):.1f} ms ({n*d*4/(t3-t2)/1e9:.1f} GB/s)")
C0 = Xd[:k].clone()
del Xd; torch.cuda.empty_cache()
p = _lib.default_params(); p.n_clusters, p.init, p.max_iter, p.tol = k, _lib.INIT_ARRAY, 20, 0.0
for rep in range(2):
    Ce = C0.clone(); inertia, n_iter = C.c_float(), C.c_int64()
    xp = (C.c_void_p * 1)(Xh.data_ptr()); rows = (C.c_int64 * 1)(n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(lib.cuml_b200_kmeans_fit_parts_f32(h.ptr, C.byref(p), xp, rows, 1, d, None, Ce.data_ptr(), C.byref(inertia), C.byref(n_iter)))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"fit rep {rep}: {dt:.3f} s -> {n_iter.value/dt:.2f} iter/s", flush=True)
