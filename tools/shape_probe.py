"""Time Lloyd steps of an arbitrary shape through the C-ABI hook (E-step / M-step kernel times from the library's own
CUDA events).  usage: python tools/shape_probe.py n d k [steps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cuml_b200 import _lib  # noqa: E402

n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
lib = _lib.load()
torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream()
h = _lib.Handle(stream=st.cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1)
cent = torch.rand((k, d), device="cuda", generator=g) * 20 - 10
X = torch.empty((n, d), dtype=torch.float32, device="cuda")
ch = 1 << 22
for s0 in range(0, n, ch):
    e0 = min(n, s0 + ch)
    X[s0:e0] = cent[torch.randint(0, k, (e0 - s0,), device="cuda", generator=g)]
    X[s0:e0] += torch.randn((e0 - s0, d), device="cuda", generator=g)
Cd = X[torch.randperm(min(n, 1 << 20), device="cuda", generator=g)[:k]].clone()
labels = torch.zeros(n, dtype=torch.int32, device="cuda")


def step():
    _lib.check(lib.cuml_b200_kmeans_lloyd_step_f32(h.ptr, X.data_ptr(), n, d, None, k, Cd.data_ptr(), labels.data_ptr(),
                                                   None, None, 0))


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.check(lib.cuml_b200_kernel_timing_enable(h.ptr, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
f_ms, f_n, u_ms, u_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
_lib.check(lib.cuml_b200_kernel_timing_read(h.ptr, C.byref(f_ms), C.byref(f_n), C.byref(u_ms), C.byref(u_n)))
ms = e0.elapsed_time(e1) / steps
print(f"n={n} d={d} k={k}: {ms:.3f} ms/step (fused {f_ms.value / max(1, f_n.value):.3f} ms, update {u_ms.value / max(1, u_n.value):.3f} ms) "
      f"variant {int(lib.cuml_b200_kmeans_estep_variant(h.ptr, d, k))} fused_update {int(lib.cuml_b200_kmeans_fused_update(h.ptr, d, k))} "
      f"hbm floor {(4.0 * n * d + 4.0 * n) / 6457.1e6:.3f} ms")
