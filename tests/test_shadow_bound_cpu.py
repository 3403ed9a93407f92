"""The decision rule of the bf16 pre-filter (kernels.cuh scan_claim_shadow), checked on the CPU against the oracle's margins:
whenever |m~ + c| > rel(d) * A~  (m~ = f32 dot of the normal with the bf16-rounded row, A~ = sum |n||x~|, c = bias / extra-dim
term), the sign of m~ + c must be the sign of the margin the reference computes from the f32 row in its own summation order.
The inputs are built to sit ON the decision boundary: rows orthogonal to the normal up to rounding, mixed magnitudes, d from 64 to
8192. (The kernel's arithmetic for m~ and A~ is ordinary f32 in another order than numpy's; the bound covers any order.)"""
import numpy as np
import pytest

import oracle


def bf16_rn(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def shadow_rel(d):
    return np.float32(0.001962) + np.float32(d) * np.float32(1.6e-7)


@pytest.mark.parametrize("metric,d", [("cosine", 64), ("cosine", 768), ("euclidean", 96), ("manhattan", 200), ("dot-product", 768), ("cosine", 8192)])
def test_certain_rows_have_the_reference_sign(metric, d):
    m = oracle.METRICS[metric]
    rng = np.random.default_rng(d * 7 + len(metric))
    n = 3000 if d <= 768 else 300
    normal = (rng.standard_normal(d) * rng.choice([1e-3, 1.0, 30.0], size=d)).astype(np.float32)
    rows = (rng.standard_normal((n, d)) * rng.choice([1e-2, 1.0, 100.0], size=(n, 1))).astype(np.float32)
    # two thirds of the rows: remove the component along the normal (margin ~ rounding noise), then add a tiny multiple back
    nn = normal.astype(np.float64)
    proj = (rows.astype(np.float64) @ nn) / (nn @ nn)
    eps = rng.choice([0.0, 1e-7, -1e-7, 1e-4, -1e-4, 3e-3, -3e-3], size=n)
    near = rng.random(n) < 0.67
    rows[near] = (rows[near].astype(np.float64) - np.outer(proj[near] - eps[near] * np.abs(proj[near] + 1e-3), nn)).astype(np.float32)
    bias = np.float32(rng.standard_normal() * 0.01) if metric in ("euclidean", "manhattan") else np.float32(0)
    nh = (float(bias), 0.0)
    if metric == "dot-product":
        nh = (float(np.float32(0.37)), 0.0)            # normal.extra_dim
    ih0 = np.abs(rng.standard_normal(n)).astype(np.float32) if metric == "dot-product" else np.zeros(n, dtype=np.float32)
    rel = shadow_rel(d)
    certain = wrong = 0
    xs = bf16_rn(rows)
    mt_all = (xs * normal[None, :]).sum(axis=1, dtype=np.float32)
    a_all = (np.abs(xs) * np.abs(normal)[None, :]).sum(axis=1, dtype=np.float32)
    for i in range(n):
        ref = oracle.margin(m, normal, nh, rows[i], (float(ih0[i]), 0.0))
        if metric == "cosine":
            mt = mt_all[i]
        elif metric == "dot-product":
            mt = np.float32(mt_all[i] + np.float32(np.float32(nh[0]) * ih0[i]))
        else:
            mt = np.float32(bias + mt_all[i])
        if abs(mt) > rel * a_all[i]:
            certain += 1
            ref_right = (not np.isnan(ref)) and not np.signbit(np.float32(ref))
            if (mt > 0) != ref_right:
                wrong += 1
    assert wrong == 0
    assert 0 < certain < n          # both outcomes occur: some rows are decided, the constructed near-plane rows are not
