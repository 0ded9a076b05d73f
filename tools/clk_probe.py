"""Role-level wait/total cycles of the single-CTA fused kernel (CTA 0) for a few shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CUML_B200_DBG_CLK"] = "1"
import torch
from cuml_b200 import _lib
lib = _lib.load(); h = _lib.Handle()
for (n, d, k) in [(16_000_000, 16, 64), (8_000_000, 32, 16), (8_000_000, 64, 128), (4_000_000, 128, 128)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((n, d), device="cuda", generator=g)
    Cd = torch.randn((k, d), device="cuda", generator=g)
    labels = torch.zeros(n, dtype=torch.int32, device="cuda")
    print("shape", n, d, k, flush=True)
    for _ in range(2):
        _lib.check(lib.cuml_b200_kmeans_assign_f32(h.ptr, X.data_ptr(), n, d, k, Cd.data_ptr(), labels.data_ptr(), 2))
    h.sync()
