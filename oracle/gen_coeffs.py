"""ORACLE tooling: derives the polynomial coefficients of the canonical
atan / asin kernels used by BOTH oracle/ref_exact.c and
se3ds_b200/csrc/canon_math.cuh, and measures their error against float64.

Canonical forms (every operation individually rounded to float32; `fma` is a
single rounding):
  atan on [0,1]:  s = t*t;  p = Horner_fma(A[n..0], s);  r = fma(p*s, t, t)
  asin on [0,.5]: s = x*x;  p = Horner_fma(B[n..0], s);  r = fma(p*s, x, x)

Run:  python oracle/gen_coeffs.py        (prints C initialisers + error stats)
The printed coefficients are pasted verbatim (as hex floats) into the two
files; tests/test_canon_math.py checks that both copies agree bit for bit.
"""
import numpy as np

F32 = np.float32


def fit_lawson(fn, lo, hi, deg, n=4001, iters=60):
  """Weighted least squares with Lawson re-weighting ~ minimax of RELATIVE-to-|f| error
  of g(s) = (f(sqrt s) - sqrt s) / (sqrt s)^3 on s in [lo, hi]."""
  k = np.arange(n)
  s = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / n)
  s = np.sort(s)
  x = np.sqrt(s)
  g = np.where(x > 1e-12, (fn(x) - x) / np.maximum(x, 1e-300)**3, 0.0)
  # scale of the final result is f(x) ~ x; error contribution of g is x^3 * dg  => weight x^2
  wgt_fix = x**2 / np.maximum(fn(x) / np.maximum(x, 1e-300), 1e-300)
  V = np.vander(s, deg + 1, increasing=True)
  w = np.ones(n)
  for _ in range(iters):
    W = w * wgt_fix
    c, *_ = np.linalg.lstsq(V * W[:, None], g * W, rcond=None)
    err = np.abs((V @ c - g) * wgt_fix)
    w = w * (err / err.max() + 1e-3)
    w /= w.sum()
  return c


def eval_f32(coefs, x, kind):
  """Emulates the canonical float32 evaluation with numpy (fma emulated in f64 then rounded:
  exact for the products of two f32 plus an f32 up to double rounding, good enough for STATS)."""
  x = x.astype(F32)
  s = (x * x).astype(F32)
  p = np.full_like(s, F32(coefs[-1]))
  for c in coefs[-2::-1]:
    p = (p.astype(np.float64) * s.astype(np.float64) + np.float64(F32(c))).astype(F32)
  ps = (p * s).astype(F32)
  return (ps.astype(np.float64) * x.astype(np.float64) + x.astype(np.float64)).astype(F32)


def ulp_err(approx, truth64):
  truth32 = truth64.astype(F32)
  ulp = np.spacing(np.abs(truth32)).astype(np.float64)
  return np.abs(approx.astype(np.float64) - truth64) / ulp


def main():
  rng = np.random.default_rng(0)
  for name, fn, hi, deg in (('ATAN', np.arctan, 1.0, 8), ('ASIN', np.arcsin, 0.25, 5)):
    c = fit_lawson(fn, 0.0, hi, deg)
    c32 = c.astype(F32)
    x = np.concatenate([rng.uniform(0, np.sqrt(hi), 2_000_000), np.linspace(0, np.sqrt(hi), 200_001)]).astype(F32)
    e = ulp_err(eval_f32(c32, x, name), fn(x.astype(np.float64)))
    print(f'// {name}: degree {deg} in s, max err {e.max():.3f} ulp, mean {e.mean():.3f} ulp')
    print(f'static const float k{name}[{deg + 1}] = {{')
    for v in c32:
      print(f'  {float(v).hex()}f,  // {v:.9e}')
    print('};')


if __name__ == '__main__':
  main()
