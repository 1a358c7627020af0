"""Numerical check of the FP64 primitives of csrc/common.cuh against 50-digit arithmetic (algorithm emulation on the
CPU: fma emulated in 80-bit long double, the MUFU seeds emulated as float32-accurate values)."""
import numpy as np
from decimal import Decimal, getcontext
getcontext().prec = 50
ld = np.longdouble
fma = lambda a, b, c: np.float64(ld(a) * ld(b) + ld(c))
B, DEG = 11, 2          # csrc/common.cuh defaults (SOBER_EXP_TAB_BITS / SOBER_EXP_DEG)
MAGIC = 6755399441055744.0

def exp_neg(s, B=B, DEG=DEG):
    T = 1 << B
    TAB = np.array([float(Decimal(2) ** (Decimal(j) / T)) for j in range(T)])
    LN2_T, L2ET = 0.69314718055994530942 / T, T / 0.69314718055994530942
    kd = fma(s, -L2ET, MAGIC)
    k = (kd.view(np.int64) & 0xffffffff).astype(np.int64)
    k = np.where(k >= 2 ** 31, k - 2 ** 32, k)
    kf = kd - MAGIC
    r = fma(kf, -LN2_T, -s)
    coef = [1 / 120, 1 / 24, 1 / 6, 0.5]
    q = np.float64(coef[5 - DEG]) + 0 * r
    for c in coef[5 - DEG + 1:]:
        q = fma(q, r, c)
    q = fma(q, r, 1.0); p = fma(q, r, 1.0)
    return np.ldexp(TAB[k & (T - 1)] * p, (k >> B).astype(np.int32))

def sqrt_pos(a):
    y = np.float64(1 / np.sqrt(a)).astype(np.float32).astype(np.float64) * (1 + 2 ** -22.0)
    g = a * y; h = 0.5 * y
    return fma(g, fma(-g, h, 0.5), g)

rng = np.random.default_rng(0)
s = np.concatenate([rng.random(20000) * 60, rng.random(5000) * 700, rng.random(5000) * 1e-3])
want = np.array([float((-Decimal(float(x))).exp()) for x in s])
for (b_, d_) in [(6, 5), (6, 4), (8, 3), (11, 2)]:
    rel = np.abs(exp_neg(s, b_, d_) - want) / want
    print("exp_neg  table 2^%-2d degree %d: max rel err %.3e   max rel err / (1 + s) %.3e" % (b_, d_, rel.max(), (rel / (1 + s)).max()))
a = np.concatenate([rng.random(100000) * 100, 10.0 ** rng.uniform(-30, 4, 100000)])
print("sqrt_pos max rel err %.3e" % (np.abs(sqrt_pos(a) - np.sqrt(a)) / np.sqrt(a)).max())
