"""C4-shaped predict (d=256, k=4096) at reduced n: timing + parity of a row sample vs fp64."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
from cuml_b200 import _lib
from oracle import lloyd
lib = _lib.load(); h = _lib.Handle()
n, d, k = 2_000_000, 256, 4096
g = torch.Generator(device="cuda").manual_seed(1)
cent = torch.rand((k, d), device="cuda", generator=g) * 20 - 10
lab = torch.randint(0, k, (n,), device="cuda", generator=g)
X = cent[lab] + torch.randn((n, d), device="cuda", generator=g)
Cd = (cent + 0.3 * torch.randn((k, d), device="cuda", generator=g)).contiguous()
labels = torch.zeros(n, dtype=torch.int32, device="cuda")
_lib.check(lib.cuml_b200_kernel_timing_enable(h.ptr, 1))
for _ in range(3):
    _lib.check(lib.cuml_b200_kmeans_assign_f32(h.ptr, X.data_ptr(), n, d, k, Cd.data_ptr(), labels.data_ptr(), 2))
f_ms, f_n, u_ms, u_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
_lib.check(lib.cuml_b200_kernel_timing_read(h.ptr, C.byref(f_ms), C.byref(f_n), C.byref(u_ms), C.byref(u_n)))
ms = f_ms.value / f_n.value
print(f"C4-shape assign n={n}: {ms:.2f} ms -> {2.0*n*k*d/ms/1e9:.1f} TFLOP/s algorithmic ({3*2.0*n*k*d/ms/1e9:.0f} issued); full 50M rows would take {ms*25/1e3:.2f} s")
idx = np.arange(0, n, 997)
Xs = X[idx].cpu().numpy(); Cn = Cd.cpu().numpy()
agree, bad = lloyd.label_disagreements_ok(Xs, Cn, labels[idx].cpu().numpy(), 2.0 ** -20)
print("label agreement on sample", agree, "inexcusable", bad)
# transform (squared distances) at the same shape on a slice the output of which fits (n_t x k fp32)
n_t = 1_000_000
out = torch.empty((n_t, k), dtype=torch.float32, device="cuda")
p = _lib.default_params(); p.n_clusters = k
torch.cuda.synchronize(); 
for rep in range(3):
    t0 = time.perf_counter()
    _lib.check(lib.cuml_b200_kmeans_transform_f32_i64(h.ptr, C.byref(p), Cd.data_ptr(), X.data_ptr(), n_t, d, out.data_ptr()))
    h.sync(); dt = time.perf_counter() - t0
print(f"C4-shape transform n={n_t}: {dt*1e3:.1f} ms -> {2.0*n_t*k*d/dt/1e12:.1f} TFLOP/s; output {n_t*k*4/dt/1e9:.0f} GB/s")
