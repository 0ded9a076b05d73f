"""CPU model of the kernel's split-precision operands (oracle/split_precision.py): the representation error of both
schemes sits at fp32 level for every BASELINE feature width, the folded half-norm pieces are exact, and an argmin
over the modelled scores only departs from the exact argmin below the gap tolerance the GPU parity tests allow."""
import numpy as np
import pytest

from oracle import blobs, lloyd
from oracle import split_precision as sp

FP32_GAP_TOL = 2.0 ** -20      # tests/test_kmeans_gpu.py


def test_rounding_helpers_are_what_the_instructions_do():
    a = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -3.14159274, 1e-30, 65504.0], dtype=np.float32)
    t = sp.tf32_truncate(a)
    assert np.all(np.abs(t) <= np.abs(a)) and np.all((t.view(np.uint32) & 0x1FFF) == 0)
    r = sp.tf32_round(a)
    assert np.all((r.view(np.uint32) & 0x1FFF) == 0) and np.all(np.abs(r - a) <= np.abs(a) * 2.0 ** -11)
    assert r[1] == np.float32(1.0 + 2.0 ** -10)                          # the tie goes away from zero (rna)
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 100
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(sp.bf16_round(x), want)                        # rn-even, as torch's own conversion


@pytest.mark.parametrize("d", [16, 32, 64, 128, 256, 1024])
def test_both_schemes_reach_fp32_level(d):
    rng = np.random.default_rng(d)
    X = (rng.standard_normal((300, d)) * 5).astype(np.float32)
    C = (rng.standard_normal((64, d)) * 5).astype(np.float32)
    e3 = sp.relative_dot_error(sp.dots_3xtf32(X, C), X, C)
    eb = sp.relative_dot_error(sp.dots_tf32_bf16(X, C), X, C)
    e1 = sp.relative_dot_error(sp.tf32_round(X).astype(np.float64) @ sp.tf32_round(C).astype(np.float64).T, X, C)
    # single tf32 is 4e-5 .. 3e-4 here; 3xTF32 models at 4e-8 .. 4e-7 and tf32 + 2 bf16 at 9e-8 .. 1.2e-6 (largest at
    # d = 16, where few terms average out): two to three orders better, and inside the 4e-6 bar of the GPU test
    assert e3 < 1e-6 and eb < 2e-6 and e1 > 20 * max(e3, eb)


def test_half_norm_pieces_are_exact_and_tf32_representable():
    rng = np.random.default_rng(1)
    C = rng.uniform(-10, 10, size=(500, 64)).astype(np.float32)
    pieces, hn = sp.half_norm_pieces(C)
    assert np.all((pieces.view(np.uint32) & 0x1FFF) == 0)                   # read unchanged by kind::tf32
    assert np.array_equal(pieces.astype(np.float64).sum(1), hn.astype(np.float64))   # 11 + 11 + 2 bits: nothing lost


@pytest.mark.parametrize("n,d,k", [(4000, 64, 256), (3000, 128, 300), (6000, 16, 64)])
def test_modelled_argmin_departs_only_below_the_gap_tolerance(n, d, k):
    X, centres, _ = blobs.make_blobs(n, d, k)
    # regime 2 centres (k data rows): many near-ties, the hard case for the label bar
    C = X[np.random.default_rng(42).choice(n, size=k, replace=False)].copy()
    pieces, _ = sp.half_norm_pieces(C)
    for dots in (sp.dots_3xtf32(X, C), sp.dots_tf32_bf16(X, C)):
        score = pieces.astype(np.float64).sum(1)[None, :] + dots            # x.c - 1/2 ||c||^2, larger is nearer
        lab = score.argmax(1).astype(np.int32)
        agree, bad = lloyd.label_disagreements_ok(X, C, lab, FP32_GAP_TOL)
        assert bad == 0 and agree >= 0.9999
