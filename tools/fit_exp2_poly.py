"""Degree-3 minimax polynomial for 2^f on [-0.5, 0.5] and an fp32 emulation of csrc/attn.cu:ex2_poly (the VBX_EXP_POLY
experiment: part of the attention exponentials on the FMA pipe instead of MUFU).  CPU only (numpy + scipy).

    python tools/fit_exp2_poly.py        # prints the coefficients and the emulated max relative error
"""
import numpy as np

MAGIC = np.float32(12582912.0)  # 1.5 * 2^23: adding it rounds to the nearest integer and leaves that integer in the low mantissa bits
COEFFS = (0.9999280571937561, 0.6932609677314758, 0.2426111251115799, 0.0551716685295105)  # c0..c3 as used in attn.cu


def fit(deg=3, n=20001):
    from scipy.optimize import minimize
    f = np.linspace(-0.5, 0.5, n)
    y = 2.0 ** f
    V = np.vander(f, deg + 1, increasing=True)
    c = np.linalg.lstsq(V / y[:, None], np.ones_like(f), rcond=None)[0]  # least-squares start, then minimise the max error
    obj = lambda c: np.max(np.abs(V @ c / y - 1))  # noqa: E731
    for _ in range(5):
        c = minimize(obj, c, method='Nelder-Mead', options=dict(xatol=1e-12, fatol=1e-14, maxiter=20000)).x
    return c, obj(c)


def ex2_poly_fp32(x, coeffs=COEFFS):
    """Bit-faithful numpy restatement of the device function (every intermediate rounded to fp32)."""
    c = np.asarray(coeffs, dtype=np.float32)
    x = np.maximum(np.asarray(x, dtype=np.float32), np.float32(-125.0))
    t = (x + MAGIC).astype(np.float32)
    f = (x - (t - MAGIC).astype(np.float32)).astype(np.float32)
    p = (f * c[3] + c[2]).astype(np.float32)  # fmaf rounds once; the double rounding here is below 1 ulp and irrelevant at 7e-5
    p = (p * f + c[1]).astype(np.float32)
    p = (p * f + c[0]).astype(np.float32)
    return (p.view(np.int32) + (t.view(np.int32) << 23)).view(np.float32)


if __name__ == '__main__':
    c, e = fit(3)
    print('minimax degree 3:', [float(np.float32(v)) for v in c], 'max rel err', e)
    x = -np.abs(np.random.default_rng(0).normal(0, 8, 4_000_000)).astype(np.float32)
    ref = np.exp2(np.maximum(x, -125).astype(np.float64))
    print('fp32 emulation of ex2_poly: max rel err', float(np.max(np.abs(ex2_poly_fp32(x) / ref - 1))))
