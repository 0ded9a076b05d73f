"""numpy model of the split-precision contraction of the fused distance + argmin kernel.  TEST INFRASTRUCTURE ONLY.

The CUDA kernel (cuml_b200/csrc/fused_l2_argmin_sm100.cu) computes ``x.c - 1/2 ||c||^2`` on
the tensor cores from low-precision operand pieces.  This module restates the *operand preparation* of both schemes
bit-exactly (what is rounded to what) and evaluates the products in fp64, so that the error it reports is the
scheme's own representation error -- the fp32 accumulation order inside the tensor core is not modelled and adds the
usual ~sqrt(d) * 2^-24 on top.  It answers, on the CPU, the question the GPU test
``test_tensor_core_dot_accuracy`` answers on hardware: is the contraction good to fp32 level, so that label
disagreements with the exact argmin can only occur below the stated gap tolerance 2^-20 (||x||^2 + ||c||^2)?
"""
from __future__ import annotations

import numpy as np


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def tf32_truncate(a):
    """fp32 with the 13 low mantissa bits cleared (what ``kind::tf32`` reads from an fp32 word)"""
    return (_bits(a) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_round(a):
    """round to the nearest tf32 (10 mantissa bits), ties away from zero -- ``cvt.rna.tf32.f32``"""
    return ((_bits(a) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_round(a):
    """round to the nearest bf16 (7 mantissa bits), ties to even -- ``cvt.rn.bf16x2.f32``; returned as fp32"""
    b = _bits(a)
    return ((b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(np.float32)


def dots_3xtf32(X, C):
    """3xTF32: hi = truncated tf32, lo = x - hi (exact in fp32, read truncated by the MMA); lo.lo dropped"""
    xh, ch = tf32_truncate(X), tf32_truncate(C)
    xl, cl = tf32_truncate(X - xh), tf32_truncate(C - ch)
    f = np.float64
    return xl.astype(f) @ ch.astype(f).T + xh.astype(f) @ cl.astype(f).T + xh.astype(f) @ ch.astype(f).T


def dots_tf32_bf16(X, C):
    """tf32 main term + two bf16 correction terms: hi = nearest tf32, lo = x - hi; the corrections use
    bf16(lo) x bf16(hi)"""
    xh, ch = tf32_round(X), tf32_round(C)
    xl, cl = (X - xh).astype(np.float32), (C - ch).astype(np.float32)
    f = np.float64
    return (bf16_round(xl).astype(f) @ bf16_round(ch).astype(f).T + bf16_round(xh).astype(f) @ bf16_round(cl).astype(f).T
            + xh.astype(f) @ ch.astype(f).T)


def half_norm_pieces(C):
    """-1/2 ||c||^2 (fp32) as three tf32-exact pieces (11 + 11 + 2 mantissa bits) folded into the accumulator by a
    ones x pieces MMA; returns (pieces [k, 3] fp32, the fp32 value they represent)"""
    hn = (-(0.5 * (C.astype(np.float64) ** 2).sum(1)).astype(np.float32)).astype(np.float32)   # fp64 row sum, rounded once
    p1 = tf32_truncate(hn)
    r1 = (hn - p1).astype(np.float32)
    p2 = tf32_truncate(r1)
    p3 = (r1 - p2).astype(np.float32)                      # <= 2 significant bits left: already tf32-exact
    return np.stack([p1, p2, p3], axis=1), hn


def relative_dot_error(dots, X, C):
    """max |dots - x.c| / (||x|| ||c||) against fp64"""
    f = np.float64
    ref = X.astype(f) @ C.astype(f).T
    scale = np.sqrt((X.astype(f) ** 2).sum(1))[:, None] * np.sqrt((C.astype(f) ** 2).sum(1))[None, :]
    return float((np.abs(dots - ref) / scale).max())
