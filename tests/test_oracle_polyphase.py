"""DESIGN.md section 5, "bilinear resize folded into the consumer convolution": the arithmetic behind the decision not to build it,
checked against the oracle's own conv2d_same(resize2x(x)) (both pinned by the OpenCV TF-importer vectors)."""
import numpy as np
import pytest
import torch

from oracle import polyphase as PP
from oracle import splitvae_oracle as O


@pytest.mark.parametrize("k", [4, 6])
@pytest.mark.parametrize("n", [4, 8, 9])
def test_one_dimensional_identity(k, n):
    rng = np.random.default_rng(k * 100 + n)
    w, x = rng.normal(size=k), rng.normal(size=n)
    direct = PP.conv_same_1d(w, PP.upsample_matrix(n) @ x)
    assert np.allclose(PP.polyphase_1d(w, x), direct, atol=1e-12)
    # the clamped up-sampling differs from the zero-extended one only by +-x/4 at positions {-1, 0} and {2n-1, 2n}
    Rp = PP.upsample_matrix(n, clamp=False) @ x                        # positions -2 .. 2n+1
    Uz = np.concatenate([np.zeros(2), PP.upsample_matrix(n) @ x, np.zeros(2)])
    D = Uz - Rp
    expect = np.zeros(2 * n + 4)
    expect[1], expect[2], expect[2 * n + 1], expect[2 * n + 2] = -x[0] / 4, x[0] / 4, x[-1] / 4, -x[-1] / 4
    assert np.allclose(D, expect, atol=1e-12)


def test_tap_counts_and_mac_ratio():
    """6 taps -> phases of 5 and 4 low-resolution taps, 4 taps -> 4 and 3: 0.5625 / 0.766 of the MACs, not (k/2 + 1)^2 / k^2"""
    rng = np.random.default_rng(1)
    for k, taps in ((6, (5, 4)), (4, (4, 3))):
        got = tuple(sorted((len(v) for v, _ in PP.phase_filters(rng.normal(size=k))), reverse=True))
        assert got == taps                                                     # (6 taps: the even phase is the long one; 4 taps: the odd one)
        ratio = (sum(taps) / 2 / k) ** 2
        assert abs(ratio - {6: 0.5625, 4: 0.765625}[k]) < 1e-12
    # step-level: d3 (4x4), d4 / d5 (6x6) carry 14.50 / 38.65 / 17.18 GFLOP per pass at C2: 1.65x fewer MACs overall
    g = {"d3": (4, 14.50), "d4": (6, 38.65), "d5": (6, 17.18)}
    before = sum(v for _, v in g.values())
    after = sum(v * {4: 0.765625, 6: 0.5625}[k] for k, v in g.values())
    assert 1.6 < before / after < 1.7


@pytest.mark.parametrize("k,ci,co,n", [(6, 3, 2, 6), (4, 2, 3, 5)])
def test_two_dimensional_identity_against_the_oracle(k, ci, co, n):
    """uniform 2-D phase convolutions (outer products of the 1-D phase filters, per (ci, co)) + border correction == the oracle;
    and away from the border rows / columns the uniform part alone is already exact"""
    rng = np.random.default_rng(5)
    x = rng.normal(size=(1, n, n, ci))
    w = rng.normal(size=(k, k, ci, co))
    ref = O.conv2d_same(O.resize2x(torch.from_numpy(x)), torch.from_numpy(w), None, 1).numpy()[0]
    # separable check per rank-1 kernel: w = sum over (a) of e_a (x) w[a, :] - use linearity: apply the 1-D identity along x for every
    # filter row, then along y
    R = PP.upsample_matrix(n)
    U = np.einsum("yi,xj,ijc->yxc", R, R, x[0])                                    # the resize as a Kronecker product
    assert np.allclose(U, O.resize2x(torch.from_numpy(x)).numpy()[0], atol=1e-12)
    # uniform part: zero-extended interpolation on both axes, then a zero-padded convolution over the positions -2 .. 2n+1
    Rp = PP.upsample_matrix(n, clamp=False)
    Up = np.einsum("yi,xj,ijc->yxc", Rp, Rp, x[0])                                 # [2n+4, 2n+4, ci]
    pl = (k - 1) // 2
    big = np.zeros((2 * n + 4 + k, 2 * n + 4 + k, ci))
    big[:2 * n + 4, :2 * n + 4] = Up
    uni = np.zeros((2 * n, 2 * n, co))
    for oy in range(2 * n):
        for ox in range(2 * n):
            for a in range(k):
                for b in range(k):
                    py, px = oy + a - pl + 2, ox + b - pl + 2
                    if 0 <= py < 2 * n + 4 and 0 <= px < 2 * n + 4:
                        uni[oy, ox] += big[py, px] @ w[a, b]
    sup = PP.correction_support(k, n)
    interior = [o for o in range(2 * n) if o not in sup]
    assert len(sup) == k + 1 or 2 * n <= k + 1                                   # pl + 1 rows at the start, k - pl at the end
    if interior:
        ii = np.ix_(interior, interior)
        assert np.allclose(uni[ii], ref[ii], atol=1e-10)                          # no correction needed away from the border
    diff = ref - uni
    mask = np.ones((2 * n, 2 * n), bool)
    mask[np.ix_(interior, interior)] = False
    assert np.abs(diff[~mask]).max(initial=0.0) < 1e-10 and np.abs(diff[mask]).max() > 1e-3    # ... and it IS needed on it
