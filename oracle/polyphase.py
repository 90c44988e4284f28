"""Design check for DESIGN.md section 5 (NOT on the product path, test infrastructure like the rest of oracle/): the decoder's
`Conv2D(k, 'same')(tf.image.resize(x, 2x))` (vae/model.py:163-167) as four phase convolutions of the LOW-resolution tensor plus a
border correction.  Written to settle the arithmetic before anyone rewrites the N-stacked kernels:

* 1-D: U = R x with the half-pixel bilinear matrix R (edge clamp), y = conv_same(w, U).  Let R' be the same interpolation of the
  ZERO-extended signal (no clamp) and U' = R' xz.  Then U_zero-padded = U' + D with D supported on up-sampled positions {-1, 0} and
  {2n-1, 2n}: D[-1] = -x[0]/4, D[0] = +x[0]/4, D[2n-1] = +x[n-1]/4, D[2n] = -x[n-1]/4.
* conv_same(w, U') is exactly two phase filters on xz: output 2i + p = sum_j v_p[j] xz[i + j - q_p]  (phase_filters below); for k = 6
  the phases have 5 and 4 taps, for k = 4: 3 and 4 - so the MAC ratio is ((5 + 4) / 2 / 6)^2 = 0.5625 resp. ((4 + 3) / 2 / 4)^2 = 0.766,
  not the (ceil(k/2)+1)^2 / k^2 of the round-1 note.
* 2-D: the operators are Kronecker products, (R' + Delta) x (R' + Delta), so the correction is Delta x R' + R' x Delta + Delta x Delta:
  it touches the outermost pl + 1 / k - pl output rows and columns only.
tests/test_oracle_polyphase.py checks all of it against oracle.conv2d_same(oracle.resize2x(x))."""
import numpy as np


def upsample_matrix(n, clamp=True):
    """[2n + 2*m, n] rows for up-sampled positions -m .. 2n-1+m with m = 0 (clamp) - or the zero-extended interpolation R' evaluated on
    positions -2 .. 2n+1 (clamp=False; rows outside [0, 2n) are what SAME padding would have had to contain for the uniform phase form)."""
    if clamp:
        R = np.zeros((2 * n, n))
        for i in range(n):
            R[2 * i, max(i - 1, 0)] += 0.25
            R[2 * i, i] += 0.75
            R[2 * i + 1, i] += 0.75
            R[2 * i + 1, min(i + 1, n - 1)] += 0.25
        return R
    R = np.zeros((2 * n + 4, n))          # positions -2 .. 2n+1, row index = position + 2
    for pos in range(-2, 2 * n + 2):
        i, p = divmod(pos, 2)
        for src, wgt in (((i - 1, 0.25), (i, 0.75)) if p == 0 else ((i, 0.75), (i + 1, 0.25))):
            if 0 <= src < n:
                R[pos + 2, src] += wgt
    return R


def phase_filters(w):
    """1-D kernel w (k taps, SAME: pad_left = (k - 1) // 2) -> [(v_p, q_p)] for p = 0, 1 with
    y[2i + p] = sum_j v_p[j] * xz[i + j - q_p]  for the zero-extended low-resolution signal xz (uniform part, no border term)."""
    k = len(w)
    pl = (k - 1) // 2
    out = []
    for p in (0, 1):
        taps = {}
        for t in range(k):
            pos = p + t - pl                       # up-sampled position relative to 2i
            i, ph = divmod(pos, 2)
            for src, wgt in (((i - 1, 0.25), (i, 0.75)) if ph == 0 else ((i, 0.75), (i + 1, 0.25))):
                taps[src] = taps.get(src, 0.0) + wgt * w[t]
        lo, hi = min(taps), max(taps)
        out.append((np.array([taps.get(s, 0.0) for s in range(lo, hi + 1)]), -lo))
    return out


def conv_same_1d(w, u):
    k = len(w)
    pl = (k - 1) // 2
    uz = np.concatenate([np.zeros(pl), u, np.zeros(k - 1 - pl)])
    return np.array([np.dot(w, uz[o:o + k]) for o in range(len(u))])


def polyphase_1d(w, x):
    """uniform phase filters on the zero-extended x  +  the border correction; equals conv_same_1d(w, R x)"""
    n, k = len(x), len(w)
    pl = (k - 1) // 2
    y = np.zeros(2 * n)
    for p, (v, q) in enumerate(phase_filters(w)):
        xz = np.concatenate([np.zeros(len(v)), x, np.zeros(len(v))])
        for i in range(n):
            y[2 * i + p] = np.dot(v, xz[len(v) + i - q:len(v) + i - q + len(v)])
    # D[-1] = -x0/4, D[0] = +x0/4, D[2n-1] = +x_last/4, D[2n] = -x_last/4, pushed through the k taps
    for pos, val in ((-1, -0.25 * x[0]), (0, 0.25 * x[0]), (2 * n - 1, 0.25 * x[-1]), (2 * n, -0.25 * x[-1])):
        for t in range(k):
            o = pos - t + pl
            if 0 <= o < 2 * n:
                y[o] += w[t] * val
    return y


def correction_support(k, n):
    """output positions the border term can touch (per axis)"""
    pl = (k - 1) // 2
    s = set()
    for pos in (-1, 0, 2 * n - 1, 2 * n):
        for t in range(k):
            o = pos - t + pl
            if 0 <= o < 2 * n:
                s.add(o)
    return sorted(s)
