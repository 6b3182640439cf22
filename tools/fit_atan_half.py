"""Fits the polynomial used by mixture_pairs (forward): atan(q) = q + q s P(s), s = q^2, |q| <= tan(asin(0.7)) = 0.98020,
minimising the maximum ABSOLUTE error of atan (Remez-style exchange on a dense grid via iteratively reweighted least squares),
then checks the fp32 Horner evaluation.  Prints the coefficients, highest power first."""
import sys
import numpy as np

QMAX = np.tan(np.arcsin(0.7)) * 1.0005


def fit(deg):
    q = np.linspace(1e-4, QMAX, 20001)
    s = q * q
    # target: (atan(q) - q) / (q s) = P(s); absolute error of atan = q s |P - target|
    tgt = (np.arctan(q) - q) / (q * s)
    wgt = q * s
    V = np.vander(s, deg + 1, increasing=True)
    w = np.ones_like(q)
    best = None
    for it in range(200):
        A = V * (wgt * w)[:, None]
        b = tgt * wgt * w
        co, *_ = np.linalg.lstsq(A, b, rcond=None)
        err = np.abs((V @ co - tgt) * wgt)
        m = err.max()
        if best is None or m < best[0]:
            best = (m, co.copy())
        w *= (1 + 4 * err / m) / 3
        w /= w.mean()
    return best


def horner32(co_hi_first, q32):
    s = (q32.astype(np.float64) * q32).astype(np.float32)
    p = np.full_like(s, np.float32(co_hi_first[0]))
    for c in co_hi_first[1:]:
        p = (p.astype(np.float64) * s + np.float64(np.float32(c))).astype(np.float32)
    ps = (p.astype(np.float64) * s).astype(np.float32)
    return (ps.astype(np.float64) * q32 + q32).astype(np.float32)


if __name__ == "__main__":
    for deg in (6, 7, 8):
        m, co = fit(deg)
        hi_first = co[::-1]
        q32 = np.linspace(0, QMAX, 2_000_001).astype(np.float32)
        e32 = np.abs(horner32(hi_first, q32).astype(np.float64) - np.arctan(q32.astype(np.float64))).max()
        print(f"deg {deg}: fit max err {m:.3e}, fp32 Horner max err {e32:.3e}")
        print("   ", ", ".join(f"{np.float32(c)!r}" for c in hi_first))
