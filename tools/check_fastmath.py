"""Numerical check of the FP64 primitives of csrc/common.cuh against 50-digit arithmetic (algorithm emulation on the
CPU: fma emulated in 80-bit long double, the MUFU seeds emulated as float32-accurate values)."""
import numpy as np
from decimal import Decimal, getcontext
getcontext().prec = 50
ld = np.longdouble
fma = lambda a, b, c: np.float64(ld(a) * ld(b) + ld(c))
TAB = np.array([float(Decimal(2) ** (Decimal(j) / 64)) for j in range(64)])
LN2_64, L2E64, MAGIC = 0.010830424696249145, 92.33248261689366, 6755399441055744.0

def exp_neg(s):
    kd = fma(s, -L2E64, MAGIC)
    k = (kd.view(np.int64) & 0xffffffff).astype(np.int64)
    k = np.where(k >= 2 ** 31, k - 2 ** 32, k)
    kf = kd - MAGIC
    r = fma(kf, -LN2_64, -s)
    q = fma(1 / 120, r, 1 / 24); q = fma(q, r, 1 / 6); q = fma(q, r, 0.5); q = fma(q, r, 1.0); p = fma(q, r, 1.0)
    return np.ldexp(TAB[k & 63] * p, (k >> 6).astype(np.int32))

def sqrt_pos(a):
    y = np.float64(1 / np.sqrt(a)).astype(np.float32).astype(np.float64) * (1 + 2 ** -22.0)
    g = a * y; h = 0.5 * y
    return fma(g, fma(-g, h, 0.5), g)

rng = np.random.default_rng(0)
s = np.concatenate([rng.random(20000) * 60, rng.random(5000) * 700, rng.random(5000) * 1e-3])
want = np.array([float((-Decimal(float(x))).exp()) for x in s])
rel = np.abs(exp_neg(s) - want) / want
print("exp_neg  max rel err %.3e   max rel err / (1 + s) %.3e" % (rel.max(), (rel / (1 + s)).max()))
a = np.concatenate([rng.random(100000) * 100, 10.0 ** rng.uniform(-30, 4, 100000)])
print("sqrt_pos max rel err %.3e" % (np.abs(sqrt_pos(a) - np.sqrt(a)) / np.sqrt(a)).max())
