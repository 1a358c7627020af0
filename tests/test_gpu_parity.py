"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the oracle
and the reference-generated golden fixtures.

Tolerances (BASELINE.json north_star): kernel / feature matrices 1e-10 relative in f64; weights 1e-6 (asserted
much tighter); indices identical; integer/index work bit-exact.
"""
import ctypes as C
import os
import warnings

import pytest
import torch

from oracle import kernels as ok
from oracle import rchq as oracle
from _cases import CASES, LOOP_CASES, Case, projector_nullspace

pytestmark = pytest.mark.gpu
STABLE = [c for c in CASES if c != "rbf2d_branin"]


@pytest.fixture(scope="module")
def ops(cuda_device):
    from sober_b200._ops import CudaOps
    return CudaOps(cuda_device)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def scaled(spec, Z, dev):
    """centre / reciprocal lengthscale (family constant folded in) the way Recombiner.run builds them."""
    from sober_b200 import _lib
    d = Z.shape[1]
    if spec.stationary:
        return Z.mean(0).contiguous(), (spec.inv_ls.to(dev) * _lib.FAMILY_SCALE[spec.family]).expand(d).contiguous()
    return torch.zeros(d, dtype=torch.float64, device=dev), torch.ones(d, dtype=torch.float64, device=dev)


def enable_bits(rec, spec, d, dev):
    """Select the bit-packed K1 path the way Recombiner.run does for {0,1} inputs with d > 8."""
    from sober_b200 import _lib
    from sober_b200._rchq import _family_values
    rec._bits, rec._lut = False, None
    if d <= _lib.RECORD_MAX_D:
        return False
    if spec.family == _lib.TANIMOTO:
        rec._bits = "tanimoto"
    elif spec.stationary and spec.inv_ls.numel() == 1:
        rec._bits = "hamming"
        step = (float(spec.inv_ls.reshape(-1)[0]) * _lib.FAMILY_SCALE[spec.family]) ** 2
        rec._lut = _family_values(spec.family, torch.arange(d + 1, dtype=torch.float64, device=dev) * step).contiguous()
    return bool(rec._bits)


def spec_tables(rec, case, dev):
    """Landmark table / prepared points for a fixture, built the way Recombiner does."""
    from sober_b200._kernel_spec import introspect
    kern = case.kernel()
    spec = introspect(kern)
    center, inv_ls = scaled(spec, case.Z, dev)
    return kern, spec, center, inv_ls


# ------------------------------------------------------------------------------------------------------------
# P0: Gram matrices, every family, both K1 variants
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1])     # 0: record/register kernel when d <= 8, 1: tiled kernel
@pytest.mark.parametrize("name", ["matern6d_rest", "rbf2d_branin", "rbf_ard5d", "ising24_hamming", "tanimoto256"])
def test_gram_matches_oracle(ops, cuda_device, name, variant):
    from sober_b200 import Recombiner, configure
    case = Case(name, cuda_device)
    with configure(k1_variant=variant) as opts:
        rec = Recombiner(ops, opts=opts)
        ops.variant = variant
        kern, spec, center, inv_ls = spec_tables(rec, case, cuda_device)
        if variant == 0:
            enable_bits(rec, spec, case.X.shape[1], cuda_device)   # binary fixtures with d > 8: bit-packed kernels
        try:
            table = rec._table(case.Z, spec, center, inv_ls)
            got_zz = rec._gram_T(rec._points(case.Z, spec, center, inv_ls), table).T
            sub = case.X[:777].contiguous()
            got_zx = rec._gram_T(rec._points(sub, spec, center, inv_ls), table).T
        finally:
            ops.variant = 0
    # north_star tolerance: 1e-10 relative.  Measured: ~1e-12 (the 4-instruction sqrt of csrc/common.cuh carries a
    # relative error of 1.3e-13, which enters exp(-r) multiplied by r)
    assert rel(got_zz, case.K_raw) < 1e-10
    assert float((got_zz - case.K_raw).abs().max()) < 1e-11
    want = kern(case.Z, sub)
    assert float((got_zx - want).abs().max()) < 1e-11 and rel(got_zx, want) < 1e-10


@pytest.mark.parametrize("nu", [0.5, 1.5])
def test_gram_other_matern_orders(ops, cuda_device, nu):
    from sober_b200 import Recombiner
    g = torch.Generator().manual_seed(1)
    X = torch.rand(500, 3, dtype=torch.float64, generator=g).to(cuda_device)
    Z = X[:40].clone()
    cov = ok.ScaleKernel(ok.MaternKernel(nu, [0.4]), 2.0).to(cuda_device)
    kern = ok.Kernel(ok.BareModel(cov), mode="kernel")
    from sober_b200._kernel_spec import introspect
    spec = introspect(kern)
    rec = Recombiner(ops)
    center, inv_ls = scaled(spec, Z, cuda_device)
    got = rec._gram_T(rec._points(X, spec, center, inv_ls), rec._table(Z, spec, center, inv_ls)).T
    want = kern(Z, X)
    # nu = 1/2 is not differentiable at r = 0: the expanded distance's rounding noise (1e-16 in d2 -> 1e-8 in r)
    # shows up on exact duplicates, in the oracle as much as here; compare away from the diagonal
    mask = want < (2.0 - 1e-6)
    assert float(((got - want).abs() * mask).max()) < 1e-11


def test_gspace_post_kernel_gram_matches_oracle(ops, cuda_device):
    """POST mode of K1 (non-linear in the posterior covariance, SOBER/BASQ/_scale_mmlt.py:256-275): the Gram
    expm1(cov_h(z, x)) against the oracle's gspace_kernel divided by its mu_g factors, 1e-10 relative."""
    from sober_b200 import Recombiner
    from sober_b200._kernel_spec import introspect
    case = Case("gspace_matern4d", cuda_device)
    kern = case.kernel()
    spec = introspect(kern)
    assert spec is not None and spec.gspace
    rec = Recombiner(ops)
    center, inv_ls = scaled(spec, case.Z, cuda_device)
    lm = rec._landmarks(case.Z, spec, center, inv_ls)
    X = case.X[:777]
    pts = ops.prepare_points(X, center, inv_ls)
    kx = rec._gram_T(pts, lm["table_obs"])
    at, _ = ops.group_accumulate(pts, lm["table_z"], None, None, len(X), 0, 0, len(X), post=(kx, lm["k_zo_w"].contiguous()))
    owner = kern.__self__
    want = owner.gspace_kernel(case.Z, X) / (owner.gspace_mean_predict(case.Z).unsqueeze(1) *
                                              owner.gspace_mean_predict(X).unsqueeze(0))
    assert rel(at.T, want) < 1e-10
    assert rel(lm["m_z"], owner.gspace_mean_predict(case.Z)) < 1e-10


# ------------------------------------------------------------------------------------------------------------
# P1: group sums per iteration (remainder quirk included), against the oracle's trace
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", LOOP_CASES)
def test_group_accumulate_matches_oracle_trace(ops, cuda_device, name, variant):
    from sober_b200 import Recombiner, configure
    from sober_b200._rchq import Alive
    case = Case(name)                                   # oracle on the CPU
    if case.objective is not None:
        pytest.skip("objective branch covered end-to-end")
    if case.mode in ("weighted_predictive_covariance", "gspace"):
        pytest.skip("the K1 weights carry m(x) inside Recombiner.run: covered end-to-end")
    groups, updates = [], []

    def trace(stage, payload):
        if stage == "group":
            groups.append(payload)
        elif stage == "update":
            updates.append(payload)
    mu0 = torch.full((len(case.X),), 1.0 / len(case.X), dtype=torch.float64) if case.mu is None else case.mu.clone()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        oracle.recombination(case.X, case.Z, case.b, case.kernel(), None, None, init_weights=mu0.clone(), trace=trace)
    alive = [torch.nonzero(mu0 != 0).reshape(-1)] + [u["alive"] for u in updates]
    mass = [mu0[alive[0]]] + [u["mass_alive"] for u in updates]

    dcase = Case(name, cuda_device)
    n = case.U.shape[0]
    S = 2 * (n + 1)
    ops.variant = variant
    try:
        with configure(k1_variant=variant) as opts:
            rec = Recombiner(ops, opts=opts, basis=dcase.U)
            kern, spec, center, inv_ls = spec_tables(rec, dcase, cuda_device)
            if variant == 0:
                enable_bits(rec, spec, dcase.X.shape[1], cuda_device)
            U, Uext, table = rec._nystrom(dcase.Z, case.b - 1, kern, spec, center, inv_ls)
            records = rec._use_records(spec, dcase.X.shape[1])
            st = {"spec": spec, "table": table, "pts": None if records else rec._points(dcase.X, spec, center, inv_ls)}
            for t, grp in enumerate(groups):
                idx = alive[t].to(cuda_device, torch.int32)
                m = mass[t].to(cuda_device)
                R = len(idx)
                E = R // S
                al = Alive(idx, m, ops.make_records(dcase.X, center, inv_ls, idx, m).rec if records else None)
                at, totw = rec._accumulate(st, al, R, 0, E * S, S)
                if spec.mode == "kernel":
                    assert rel(at.T.cpu(), grp["A"]) < 1e-10
                # projected (unnormalised) barycentres incl. the second count of the remainder
                bary = at @ Uext.T
                if R > E * S:
                    tail_at, tail_tw = rec._accumulate(st, al.tail(E * S), R - E * S, 0, R - E * S, 1)
                    bary[S - 1] += (tail_at @ Uext.T)[0]
                    totw[S - 1] += tail_tw[0]
                assert rel(totw.cpu(), grp["totw"]) < 1e-13
                assert rel(bary.cpu(), grp["Xt_unnormalised"]) < 1e-10
    finally:
        ops.variant = 0


# ------------------------------------------------------------------------------------------------------------
# P2: CAR elimination on the reference's own null-space bases: bit-identical
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("force_global", [False, True])
@pytest.mark.parametrize("name", STABLE + ["rbf2d_branin"])
def test_car_elimination_bitwise(ops, cuda_device, name, force_global):
    case = Case(name, cuda_device)
    if case.n_car == 0:
        pytest.skip("no CAR call in this fixture")
    if force_global:
        os.environ["SOBER_B200_CAR_FORCE_GLOBAL"] = "1"
    try:
        for i in range(case.n_car):
            phi = case.car(i, "Phi")                     # (N x k), the reference's Vh tail transposed
            mass = case.car(i, "mu").clone().contiguous()
            ops.car_eliminate(phi.T.contiguous(), mass)
            torch.cuda.synchronize()
            keep = mass > 0
            assert torch.equal(torch.nonzero(keep).reshape(-1), case.car(i, "idx"))
            assert torch.equal(mass[keep], case.car(i, "w"))
    finally:
        os.environ.pop("SOBER_B200_CAR_FORCE_GLOBAL", None)


@pytest.mark.parametrize("name", STABLE + ["rbf2d_branin"])
def test_car_cluster_elimination_bitwise(ops, cuda_device, name):
    """The cluster-resident kernel in EXACT mode on the reference's own null-space bases: bit-identical."""
    case = Case(name, cuda_device)
    if case.n_car == 0:
        pytest.skip("no CAR call in this fixture")
    for i in range(case.n_car):
        phi = case.car(i, "Phi")
        assert ops.car_cluster_fits(phi.shape[0], phi.shape[0] - phi.shape[1], True) > 0
        mass = case.car(i, "mu").clone().contiguous()
        ops.car_cluster(mass, basis_rows=phi.T.contiguous(), exact=True)
        torch.cuda.synchronize()
        keep = mass > 0
        assert torch.equal(torch.nonzero(keep).reshape(-1), case.car(i, "idx"))
        assert torch.equal(mass[keep], case.car(i, "w"))


@pytest.mark.parametrize("name", STABLE + ["rbf2d_branin"])
def test_car_cols_elimination_bitwise(ops, cuda_device, name):
    """The column-distributed cluster kernel in EXACT mode on the reference's own null-space bases: bit-identical."""
    case = Case(name, cuda_device)
    if case.n_car == 0:
        pytest.skip("no CAR call in this fixture")
    for i in range(case.n_car):
        phi = case.car(i, "Phi")
        assert ops.car_cols_fits(phi.shape[0], phi.shape[1]) > 0
        mass = case.car(i, "mu").clone().contiguous()
        ops.car_cols(phi.T.contiguous(), mass, exact=True)
        torch.cuda.synchronize()
        keep = mass > 0
        assert torch.equal(torch.nonzero(keep).reshape(-1), case.car(i, "idx"))
        assert torch.equal(mass[keep], case.car(i, "w"))


@pytest.mark.parametrize("S,n_prime", [(400, 200), (448, 200), (200, 100), (96, 41), (33, 7), (401, 199)])
def test_car_cols_projector_basis(ops, cuda_device, S, n_prime):
    """Fast mode at bench size: projector null space + column-distributed elimination vs the same basis through the
    oracle's elimination (support identical, weights 1e-10, moments preserved)."""
    from sober_b200 import _car
    g = torch.Generator().manual_seed(S + n_prime)
    feats = torch.randn(S, n_prime - 1, dtype=torch.float64, generator=g) * \
        torch.logspace(0, -4, n_prime - 1, dtype=torch.float64)
    mass = torch.rand(S, dtype=torch.float64, generator=g)
    mass /= mass.sum()
    design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1)
    got = _car.caratheodory(ops, feats.to(cuda_device), mass.to(cuda_device), "projector").cpu()
    want = mass.clone()
    oracle.eliminate(projector_nullspace(design), want, oracle.Factory())
    assert int((got > 0).sum()) <= n_prime
    assert torch.equal(got > 0, want > 0)
    assert float((got - want).abs().max()) < 1e-10
    assert float((design.T @ got - design.T @ mass).abs().max()) < 1e-12


@pytest.mark.parametrize("S,n_prime", [(48, 24), (400, 200), (96, 41), (200, 100), (33, 7)])
def test_car_cluster_fused_qr(ops, cuda_device, S, n_prime):
    """QR + null space + elimination in one kernel vs LAPACK QR + the oracle's elimination: same support, same
    weights, moments [1 X]^T w preserved to rounding."""
    g = torch.Generator().manual_seed(S * 1000 + n_prime)
    feats = torch.randn(S, n_prime - 1, dtype=torch.float64, generator=g) * \
        torch.logspace(0, -5, n_prime - 1, dtype=torch.float64)            # columns spanning 5 decades
    mass = torch.rand(S, dtype=torch.float64, generator=g)
    mass /= mass.sum()
    design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1)
    assert ops.car_cluster_fits(S, n_prime, False) > 0
    got = mass.clone().to(cuda_device)
    ops.car_cluster(got, design=design.to(cuda_device).contiguous())
    torch.cuda.synchronize()
    got = got.cpu()
    want = mass.clone()
    phi = torch.linalg.qr(design, mode="complete").Q[:, n_prime:]
    oracle.eliminate(phi.clone(), want, oracle.Factory())
    assert int((got > 0).sum()) <= n_prime
    assert torch.equal(got > 0, want > 0)
    assert float((got - want).abs().max()) < 1e-10
    assert float((design.T @ got - design.T @ mass).abs().max()) < 1e-13


def test_car_early_stop_guard(ops, cuda_device):
    """No positive entry in the leading null vector -> stop (SOBER/_rchq.py:241-242)."""
    rows = -torch.ones((3, 8), dtype=torch.float64, device=cuda_device)
    mass = torch.full((8,), 0.125, dtype=torch.float64, device=cuda_device)
    piv, steps = ops.car_eliminate(rows.clone(), mass, want_pivots=True)
    assert int(steps) == 0 and bool((piv == -1).all())
    assert torch.equal(mass, torch.full_like(mass, 0.125))


# ------------------------------------------------------------------------------------------------------------
# streaming passes: bit-exact
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 1_000_003])
def test_compact_nonzero_exact(ops, cuda_device, n):
    g = torch.Generator().manual_seed(n)
    mu = torch.rand(n, dtype=torch.float64, generator=g)
    mu[torch.rand(n, generator=g) < 0.3] = 0.0
    mu = mu.to(cuda_device)
    idx, out, cnt = ops.compact_nonzero(mu)
    want = torch.nonzero(mu != 0).reshape(-1)
    assert cnt == len(want)
    assert torch.equal(idx.long(), want) and torch.equal(out, mu[want])


@pytest.mark.parametrize("S,R,pos0", [(48, 4000, 0), (48, 3072, 0), (400, 1_000_000, 0), (10, 1003, 0),
                                      (48, 4000, 1234), (2000, 5_000_123, 77)])
def test_update_compact_exact(ops, cuda_device, S, R, pos0):
    from sober_b200._rchq import KeepMap
    g = torch.Generator().manual_seed(S + R)
    E = R // S
    ES = E * S
    n_local = (R - pos0) // 2 if pos0 else R          # a middle shard when pos0 > 0
    kept = torch.rand(S, generator=g) < 0.5
    kept[S - 1] = bool(R % 2)
    wstar = torch.where(kept, torch.rand(S, dtype=torch.float64, generator=g) + 0.1, torch.zeros(S, dtype=torch.float64))
    totw = torch.rand(S, dtype=torch.float64, generator=g) + 0.5
    idx = torch.sort(torch.randperm(2 * R, generator=g)[:n_local])[0].to(torch.int32)
    mu = torch.rand(n_local, dtype=torch.float64, generator=g)
    km = KeepMap(kept.tolist(), S, ES)
    new_pos0 = km.before(pos0)
    n_out = km.before(pos0 + n_local) - new_pos0
    rank = (torch.cumsum(kept.int(), 0) - kept.int()).to(torch.int32)
    d = lambda t: t.to(cuda_device)
    dim = 6 if S % 2 == 0 else 3                       # record strides 8 and 6 (one with a pad slot)
    ldr = (dim + 3) // 2 * 2
    rec = torch.rand(n_local, ldr, dtype=torch.float64, generator=g)
    rec[:, dim + 1] = mu
    idx_o, mu_o, rec_o = ops.update_compact(d(idx), d(mu), n_local, pos0, ES, S, d(wstar), d(totw), d(rank), km.K,
                                            km.tail_keep, new_pos0, n_out, rec=d(rec), d=dim)
    pos = pos0 + torch.arange(n_local)
    grp = torch.where(pos < ES, pos % S, torch.full_like(pos, S - 1))
    live = torch.where(pos < ES, kept[grp], torch.full_like(pos, km.tail_keep, dtype=torch.bool))
    assert int(live.sum()) == n_out
    assert torch.equal(idx_o.cpu(), idx[live])
    want_mu = (mu[live] * wstar[grp[live]]) / totw[grp[live]]
    assert torch.equal(mu_o.cpu(), want_mu)
    want_rec = rec[live].clone()
    want_rec[:, dim + 1] = want_mu
    assert torch.equal(rec_o.cpu(), want_rec)
    # the device-driven variant (K, tail_keep, new_pos0 from the cumulative kept-count on the device): same bits
    summary = torch.cat([torch.cumsum(kept.int(), 0).to(torch.int32), torch.ones(1, dtype=torch.int32)])
    idx_d, mu_d, rec_d = ops.update_compact_dev(d(idx), d(mu), n_local, pos0, ES, S, d(wstar), d(totw), d(rank), d(summary),
                                                rec=d(rec), d=dim)
    assert torch.equal(idx_d[:n_out], idx_o) and torch.equal(mu_d[:n_out], mu_o) and torch.equal(rec_d[:n_out], rec_o)


@pytest.mark.parametrize("n,d", [(1000, 6), (333, 2), (257, 24), (100, 300)])
def test_prepare_points_and_norms(ops, cuda_device, n, d):
    g = torch.Generator().manual_seed(d)
    X = torch.randn(n, d, dtype=torch.float64, generator=g).to(cuda_device)
    c = torch.randn(d, dtype=torch.float64, generator=g).to(cuda_device)
    s = (torch.rand(d, dtype=torch.float64, generator=g) + 0.5).to(cuda_device)
    pts = ops.prepare_points(X, c, s)
    want = (X - c) * s
    assert torch.equal(pts.rows[:, :d], want)
    assert float((pts.xn - (want * want).sum(-1)).abs().max()) < 1e-12 * d
    raw = ops.raw_points(X)
    assert float((raw.xn - (X * X).sum(-1)).abs().max()) < 1e-12 * d


@pytest.mark.parametrize("n,d", [(5000, 6), (999, 3), (1234, 8), (77, 1)])
def test_make_records(ops, cuda_device, n, d):
    g = torch.Generator().manual_seed(n)
    X = torch.randn(n, d, dtype=torch.float64, generator=g).to(cuda_device)
    c = torch.randn(d, dtype=torch.float64, generator=g).to(cuda_device)
    s = (torch.rand(d, dtype=torch.float64, generator=g) + 0.5).to(cuda_device)
    idx = torch.sort(torch.randperm(n, generator=g)[:n // 2])[0].to(cuda_device, torch.int32)
    mu = torch.rand(n // 2, dtype=torch.float64, generator=g).to(cuda_device)
    ps = ops.make_records(X, c, s, idx, mu)
    want = (X[idx.long()] - c) * s
    assert ps.rec.shape == (n // 2, (d + 3) // 2 * 2)
    assert torch.equal(ps.rec[:, :d], want)
    assert float((ps.rec[:, d] - (want * want).sum(-1)).abs().max()) < 1e-12 * d
    assert torch.equal(ps.rec[:, d + 1], mu)
    assert bool((ps.rec[:, d + 2:] == 0).all())
    full = ops.make_records(X, c, s)
    assert full.rec.shape[0] == n and bool((full.rec[:, d + 1] == 1).all())


@pytest.mark.parametrize("n,d", [(1000, 24), (513, 64), (777, 100), (2000, 256), (300, 1024), (100, 2048), (50, 1500)])
def test_pack_bits(ops, cuda_device, n, d):
    import numpy as np
    g = torch.Generator().manual_seed(d)
    X = (torch.rand(n, d, generator=g) < 0.1).to(torch.float64)
    words, popc, ok = ops.pack_bits(X.to(cuda_device))
    assert ok and words.shape[0] == n and words.shape[1] * 64 >= d
    bits = np.unpackbits(words.cpu().numpy().view(np.uint8), axis=1, bitorder="little")
    assert np.array_equal(bits[:, :d], X.numpy().astype(np.uint8))
    assert not bits[:, d:].any()
    assert torch.equal(popc.cpu(), X.sum(-1))
    X[n // 2, d // 3] = 0.5
    assert not ops.pack_bits(X.to(cuda_device))[2]


@pytest.mark.parametrize("d,density", [(1024, 0.05), (2048, 0.02), (512, 0.3), (167, 0.2)])
def test_tanimoto_popcount_gram_matches_oracle(ops, cuda_device, d, density):
    """Bit-packed Tanimoto (popcount) against the oracle's float Tanimoto: exact integer dot products, so the only
    difference is the rounding of the division."""
    from sober_b200 import Recombiner
    from sober_b200._kernel_spec import introspect
    g = torch.Generator().manual_seed(d)
    X = (torch.rand(3000, d, generator=g) < density).to(torch.float64).to(cuda_device)
    X[5] = 0                                                     # an all-zero fingerprint: eps keeps it finite
    Z = X[torch.randperm(3000, generator=g)[:333].to(cuda_device)].clone()
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("tanimoto", 1.0, 1.9).to(cuda_device)), mode="kernel")
    rec = Recombiner(ops)
    rec._bits = "tanimoto"
    spec = introspect(kern)
    center, inv_ls = scaled(spec, Z, cuda_device)
    got = rec._gram_T(rec._points(X, spec, center, inv_ls), rec._table(Z, spec, center, inv_ls)).T
    want = kern(Z, X)
    # north_star tolerance: 1e-10.  One Newton step on the MUFU reciprocal seed (csrc/common.cuh): 3.3e-13 measured
    assert rel(got, want) < 1e-12


@pytest.mark.parametrize("fam,nu,d", [("rbf", None, 24), ("matern", 2.5, 24), ("matern", 1.5, 100), ("rbf", None, 640)])
def test_hamming_lut_gram_matches_oracle(ops, cuda_device, fam, nu, d):
    """Stationary kernels with one lengthscale on {0,1}^d: k = lut[popcount(x ^ z)] vs the oracle's float kernel."""
    from sober_b200 import Recombiner
    from sober_b200._kernel_spec import introspect
    g = torch.Generator().manual_seed(d)
    X = (torch.rand(2500, d, generator=g) < 0.5).to(torch.float64).to(cuda_device)
    Z = X[torch.randperm(2500, generator=g)[:260].to(cuda_device)].clone()
    cov = ok.ScaleKernel(ok.RBFKernel([3.0]) if fam == "rbf" else ok.MaternKernel(nu, [3.0]), 0.7).to(cuda_device)
    kern = ok.Kernel(ok.BareModel(cov), mode="kernel")
    spec = introspect(kern)
    rec = Recombiner(ops)
    assert enable_bits(rec, spec, d, cuda_device) and rec._bits == "hamming"
    center, inv_ls = scaled(spec, Z, cuda_device)
    got = rec._gram_T(rec._points(X, spec, center, inv_ls), rec._table(Z, spec, center, inv_ls)).T
    want = kern(Z, X)
    assert rel(got, want) < 1e-11


@pytest.mark.parametrize("m,q", [(400, 200), (1000, 199), (37, 33), (513, 256), (8, 1)])
def test_trsm_right_upper(ops, cuda_device, m, q):
    from sober_b200._linalg import solve_right_upper
    g = torch.Generator().manual_seed(m + q)
    a = torch.randn(q + 5, q, dtype=torch.float64, generator=g)
    r = torch.linalg.cholesky(a.T @ a).T.contiguous().to(cuda_device)        # well-conditioned upper factor
    y = torch.randn(m, q, dtype=torch.float64, generator=g).to(cuda_device)
    got = solve_right_upper(r, y)
    want = torch.linalg.solve_triangular(r, y, upper=True, left=False)
    assert rel(got, want) < 1e-11
    assert rel(got @ r, y) < 1e-12
    # non-contiguous upper view of a lower factor, as the callers pass it
    low = r.T.contiguous()
    assert rel(solve_right_upper(low.mH, y), want) < 1e-11


@pytest.mark.parametrize("q", [1, 2, 31, 32, 33, 64, 65, 100, 160, 199, 200, 224])
def test_cholesky_upper_cluster_kernel(ops, cuda_device, q):
    """2-CTA register-resident Cholesky (csrc/chol_pair.cu) vs LAPACK: same factor to rounding, R^T R = G, exact zeros
    below the diagonal, info = 0."""
    from sober_b200._linalg import cholesky_upper
    g = torch.Generator().manual_seed(q)
    a = torch.randn(2 * q + 3, q, dtype=torch.float64, generator=g)
    gram = a.T @ a
    want = torch.linalg.cholesky(gram).T
    r, info = cholesky_upper(gram.to(cuda_device))
    assert int(info) == 0
    assert torch.equal(torch.tril(r, -1), torch.zeros_like(r))
    assert rel(r.cpu(), want) < 1e-12
    assert rel((r.T @ r).cpu(), gram) < 1e-14
    # padded leading dimension (a view into a wider matrix)
    wide = torch.zeros(q, q + 7, dtype=torch.float64, device=cuda_device)
    wide[:, :q] = gram.to(cuda_device)
    from sober_b200 import _lib
    import ctypes as C
    out = torch.full((q, q + 3), 7.0, dtype=torch.float64, device=cuda_device)
    inf2 = torch.ones((), dtype=torch.int32, device=cuda_device)
    _lib.check(_lib.load().sober_cholesky_upper(C.c_void_p(wide.data_ptr()), q + 7, q, C.c_void_p(out.data_ptr()), q + 3,
                                                C.c_void_p(inf2.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "chol")
    assert int(inf2) == 0 and torch.equal(out[:, :q], r) and bool((out[:, q:] == 7.0).all())


@pytest.mark.parametrize("q,bad", [(200, 137), (64, 0), (33, 32), (100, 99)])
def test_cholesky_upper_reports_first_bad_pivot(ops, cuda_device, q, bad):
    """Not positive definite: info = 1 + first failing column (LAPACK's convention), NaN from that column on."""
    from sober_b200._linalg import cholesky_upper
    g = torch.Generator().manual_seed(q + bad)
    a = torch.randn(2 * q, q, dtype=torch.float64, generator=g)
    gram = a.T @ a
    gram[bad, bad] = -1.0 if bad == 0 else gram[bad, bad] - 1e6
    _, want_info = torch.linalg.cholesky_ex(gram)
    r, info = cholesky_upper(gram.to(cuda_device))
    assert int(info) == int(want_info) == bad + 1
    assert bool(torch.isnan(r[bad, bad])) and bool(torch.isfinite(r[:bad]).all())


@pytest.mark.parametrize("S,Lp,n", [(400, 1000, 199), (48, 96, 23), (200, 530, 99), (33, 7, 5), (2000, 300, 999)])
def test_project_design_dmma(ops, cuda_device, S, Lp, n):
    """Projection + tail + barycentres + ones column (DMMA kernel) vs the same in torch."""
    g = torch.Generator().manual_seed(S + Lp + n)
    at = torch.randn(S, Lp, dtype=torch.float64, generator=g).to(cuda_device)
    uext = torch.randn(n, Lp, dtype=torch.float64, generator=g).to(cuda_device)
    totw = (torch.rand(S, dtype=torch.float64, generator=g) + 0.5).to(cuda_device)
    tail = torch.randn(Lp, dtype=torch.float64, generator=g).to(cuda_device)
    tail_tw = torch.tensor([0.37], dtype=torch.float64, device=cuda_device)
    design, tw = ops.project_design(at, uext, totw, tail=tail, tail_tw=tail_tw)
    at2 = at.clone()
    at2[S - 1] += tail
    tw2 = totw.clone()
    tw2[S - 1] += 0.37
    want = torch.cat([torch.ones(S, 1, dtype=torch.float64, device=cuda_device), (at2 @ uext.T) / tw2.unsqueeze(1)], 1)
    assert torch.equal(tw, tw2)
    assert rel(design, want) < 1e-12
    design0, tw0 = ops.project_design(at, uext, totw)
    assert torch.equal(tw0, totw) and rel(design0[:, 1:], (at @ uext.T) / totw.unsqueeze(1)) < 1e-12


def test_scatter_result(ops, cuda_device):
    dst = torch.rand(1000, dtype=torch.float64, device=cuda_device)
    idx = torch.tensor([3, 17, 999], device=cuda_device)
    w = torch.tensor([0.2, 0.3, 0.5], dtype=torch.float64, device=cuda_device)
    ops.scatter_result(dst, idx, w)
    assert float(dst.sum()) == 1.0 and torch.equal(dst[idx], w)


# ------------------------------------------------------------------------------------------------------------
# P3: end to end
# ------------------------------------------------------------------------------------------------------------
def _golden_nullspace(case):
    calls = {"i": 0}

    def ns(design):
        phi = case.car(calls["i"], "Phi")
        calls["i"] += 1
        assert phi.shape[0] == design.shape[0]
        return phi.T.contiguous()
    return ns


@pytest.mark.parametrize("name", STABLE)
def test_end_to_end_against_reference_fixture(ops, cuda_device, name):
    """The reference's Nystrom basis and null-space bases (CPU LAPACK) injected; everything else -- Gram
    evaluations, group sums, projection, elimination, weight updates, compaction -- on the GPU."""
    from sober_b200 import Recombiner, configure
    case = Case(name, cuda_device)
    mu = None if case.mu is None else case.mu.clone()
    with warnings.catch_warnings(), configure(mode="parity") as opts:
        warnings.simplefilter("ignore")
        rec = Recombiner(ops, opts=opts, basis=case.U, nullspace=_golden_nullspace(case))
        idx, w = rec.run(case.X, case.Z, case.b, case.kernel(), init_weights=mu, calc_obj=case.objective)
    assert idx.is_cuda and idx.dtype == torch.int64
    assert torch.equal(idx, case.idx)
    assert float((w - case.w).abs().max()) < 1e-9
    if mu is not None:
        want = torch.from_numpy(case.raw["mu_after"]).to(cuda_device)
        assert float((mu - want).abs().max()) < 1e-9


@pytest.mark.parametrize("name", ["matern6d_rest", "matern6d_pow2", "rbf_ard5d", "ising24_hamming", "tanimoto256",
                                  "predcov_matern6d", "wpredcov_matern6d", "gspace_matern4d"])
def test_parity_mode_equals_oracle_on_same_device(ops, cuda_device, name):
    """parity mode vs the oracle executing the reference's op sequence on the SAME device (same CUDA generator
    draw for svd_lowrank, same cuSOLVER SVD for the null space): identical points, weights to 1e-6."""
    import sober_b200
    case = Case(name, cuda_device)
    kern = case.kernel()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mu_o = None if case.mu is None else case.mu.clone()
        torch.manual_seed(11)
        idx_o, w_o = oracle.recombination(case.X, case.Z, case.b, kern, cuda_device, None, init_weights=mu_o)
        mu_g = None if case.mu is None else case.mu.clone()
        torch.manual_seed(11)
        with sober_b200.configure(mode="parity"):
            idx_g, w_g = sober_b200.recombination(case.X, case.Z, case.b, kern, cuda_device, torch.float64,
                                                  init_weights=mu_g)
    assert torch.equal(idx_g, idx_o)
    assert float((w_g - w_o).abs().max()) < 1e-6
    if mu_g is not None:
        assert float((mu_g - mu_o).abs().max()) < 1e-6


@pytest.mark.parametrize("name", ["matern6d_rest", "matern6d_pow2", "rbf_ard5d", "tanimoto256", "predcov_matern6d",
                                  "ising24_hamming", "wpredcov_matern6d", "gspace_matern4d", "objective_matern4d"])
def test_fast_mode_equals_cpu_oracle_with_projector_nullspace(ops, cuda_device, name):
    """fast mode (CUDA Gram, Cholesky gate, projector null space, cluster elimination kernel) vs the CPU oracle fed
    the same test matrix and the same null-space construction restated with LAPACK (tests/_cases.py)."""
    import sober_b200
    from sober_b200 import _nystrom
    cpu = Case(name)
    R = torch.randn(cpu.Z.shape[0], cpu.b - 1, dtype=torch.float64, generator=torch.Generator().manual_seed(5))

    orig = torch.randn
    torch.randn = lambda *a, **k: R.clone() if tuple(a[:2]) == tuple(R.shape) else orig(*a, **k)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mu_o = None if cpu.mu is None else cpu.mu.clone()
            idx_o, w_o = oracle.recombination(cpu.X, cpu.Z, cpu.b, cpu.kernel(), None, None, init_weights=mu_o,
                                              nullspace=projector_nullspace, calc_obj=cpu.objective)
    finally:
        torch.randn = orig
    gpu = Case(name, cuda_device)
    _nystrom._injected_test_matrix = R
    try:
        with warnings.catch_warnings(), sober_b200.configure(mode="fast"):
            warnings.simplefilter("ignore")
            mu_g = None if gpu.mu is None else gpu.mu.clone()
            idx_g, w_g = sober_b200.recombination(gpu.X, gpu.Z, gpu.b, gpu.kernel(), None, None, init_weights=mu_g,
                                                  calc_obj=gpu.objective)
    finally:
        _nystrom._injected_test_matrix = None
    assert torch.equal(idx_g.cpu(), idx_o)
    assert float((w_g.cpu() - w_o).abs().max()) < 1e-6
    kern = cpu.kernel()
    mu_full = torch.full((len(cpu.X),), 1.0 / len(cpu.X), dtype=torch.float64) if cpu.mu is None else cpu.mu
    mmd_o = oracle.mmd_squared(kern, cpu.X, mu_full, idx_o, w_o)
    mmd_g = oracle.mmd_squared(kern, cpu.X, mu_full, idx_g.cpu(), w_g.cpu())
    assert abs(float(mmd_g - mmd_o)) <= 1e-6 * abs(float(mmd_o)) + 1e-15


@pytest.mark.parametrize("name", ["matern6d_rest", "predcov_matern6d"])
def test_generic_callable_path_on_gpu(ops, cuda_device, name):
    import sober_b200
    case = Case(name, cuda_device)
    kern = case.kernel()
    res = []
    for fuse in (True, False):
        with warnings.catch_warnings(), sober_b200.configure(mode="parity", fuse=fuse, generic_chunk=1000):
            warnings.simplefilter("ignore")
            torch.manual_seed(3)
            mu = None if case.mu is None else case.mu.clone()
            res.append(sober_b200.recombination(case.X, case.Z, case.b, kern, None, None, init_weights=mu))
    assert torch.equal(res[0][0], res[1][0])
    assert float((res[0][1] - res[1][1]).abs().max()) < 1e-8


def test_host_tensors_accepted_and_mutated(ops, cuda_device):
    """CPU inputs (the e2e leg of bench.py): copied to the device, result on the device, weights mutated on host."""
    import sober_b200
    case = Case("matern6d_rest")
    kern = Case("matern6d_rest", cuda_device).kernel()
    mu = case.mu.clone()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        idx, w = sober_b200.recombination(case.X, case.Z, case.b, kern, None, None, init_weights=mu)
    assert idx.is_cuda and int((mu != 0).sum()) == len(idx)
    assert torch.equal(mu[idx.cpu()], w.cpu())


# ------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: N = 1e6, L = 1000, b = 200, Matern-5/2, 6-D)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_cand", [1_000_000, 400 * 2 ** 11])
def test_full_size_invariants(ops, cuda_device, n_cand):
    import sober_b200
    from sober_b200 import Recombiner, configure
    g = torch.Generator(device=cuda_device).manual_seed(0)
    X = torch.rand(n_cand, 6, dtype=torch.float64, device=cuda_device, generator=g)
    Z = X[torch.randperm(n_cand, device=cuda_device, generator=g)[:1000]].clone()
    mu = torch.rand(n_cand, dtype=torch.float64, device=cuda_device, generator=g)
    mu /= mu.sum()
    mu0 = mu.clone()
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("matern", [0.5], 1.0).to(cuda_device)), mode="kernel")
    b = 200
    with warnings.catch_warnings(), configure(mode="fast") as opts:
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        rec = Recombiner(ops, opts=opts)
        U, _, _ = rec._nystrom(Z, b - 1, kern, None, None, None)     # basis (same seed -> same as in run)
        torch.manual_seed(7)
        idx, w = rec.run(X, Z, b, kern, init_weights=mu)
    assert len(idx) <= b and bool((idx[1:] > idx[:-1]).all()) and bool((w > 0).all())
    assert abs(float(w.sum()) - 1.0) < 1e-12
    assert int((mu != 0).sum()) == len(idx) and torch.equal(mu[idx], w)
    # Nystrom feature means: preserved to rounding when N = S * 2^k; when a remainder occurs the reference's
    # double count (SOBER/_rchq.py:128-136 + 153-164) breaks it -- reproduced, so only a loose bound applies
    feats_sel = U @ kern(Z, X[idx])
    full = torch.zeros(U.shape[0], dtype=torch.float64, device=cuda_device)
    for s in range(0, n_cand, 1 << 17):
        full += U @ (kern(Z, X[s:s + (1 << 17)]) @ mu0[s:s + (1 << 17)])
    err = float((feats_sel @ w - full).abs().max())
    if n_cand == 400 * 2 ** 11:
        assert err < 1e-11
    else:
        assert err < 0.2


# ------------------------------------------------------------------------------------------------------------
# streams and graphs: scheduling only, results must not move
# ------------------------------------------------------------------------------------------------------------
def test_partition_stream_runs_kernels(ops, cuda_device):
    """sober_partition_stream: a stream confined to fewer SMs than the device has; kernels launched on it run and
    order with the main stream through events."""
    from sober_b200._linalg import cholesky_upper
    part = ops.partition_stream()
    if part is None:
        pytest.skip("driver cannot partition the device (no green contexts)")
    sms = C.c_int()
    from sober_b200 import _lib
    _lib.check(_lib.load().sober_sm_count(C.byref(sms)), "sm_count")
    assert 0 < ops.partition_sms < sms.value
    g = torch.Generator().manual_seed(3)
    a = torch.randn(300, 120, dtype=torch.float64, generator=g).to(cuda_device)
    gram = a.T @ a
    part.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(part):
        r, info = cholesky_upper(gram)
        r.record_stream(torch.cuda.current_stream())
    torch.cuda.current_stream().wait_stream(part)
    assert int(info) == 0 and rel(r.T @ r, gram) < 1e-13


@pytest.mark.parametrize("name", ["matern6d_rest", "predcov_matern6d", "tanimoto256"])
def test_overlap_and_graphs_do_not_change_results(ops, cuda_device, name):
    """The SM-partitioned first K1 pass and the CUDA-graph replay of the Caratheodory steps are pure scheduling: the
    same indices and weights as the plain single-stream, eager run (three calls each, so that graphs get captured
    and replayed)."""
    import sober_b200
    case = Case(name, cuda_device)
    kern = case.kernel()

    def run(**kw):
        outs = []
        with warnings.catch_warnings(), sober_b200.configure(mode="fast", **kw):
            warnings.simplefilter("ignore")
            for _ in range(3):
                torch.manual_seed(21)
                mu = None if case.mu is None else case.mu.clone()
                outs.append(sober_b200.recombination(case.X, case.Z, case.b, kern, None, None, init_weights=mu))
        return outs

    plain = run(overlap=False, graphs=False)
    fancy = run(overlap=True, graphs=True)
    # (cuBLAS may pick another GEMM algorithm under capture: weights to rounding, indices exactly)
    for (i0, w0), (i1, w1) in zip(plain, fancy):
        assert torch.equal(i0, i1) and float((w0 - w1).abs().max()) < 1e-12
    assert torch.equal(plain[0][0], plain[2][0]) and torch.equal(plain[0][1], plain[2][1])


def test_full_size_overlap_and_graphs(ops, cuda_device):
    """Same at BASELINE configs[1] size (N = 1e6, where the first K1 pass really runs beside the range finder)."""
    import sober_b200
    g = torch.Generator(device=cuda_device).manual_seed(0)
    n_cand = 1_000_000
    X = torch.rand(n_cand, 6, dtype=torch.float64, device=cuda_device, generator=g)
    Z = X[torch.randperm(n_cand, device=cuda_device, generator=g)[:1000]].clone()
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("matern", [0.5], 1.0).to(cuda_device)), mode="kernel")
    res = {}
    for fancy in (False, True):
        with warnings.catch_warnings(), sober_b200.configure(mode="fast", overlap=fancy, graphs=fancy):
            warnings.simplefilter("ignore")
            for _ in range(2):
                torch.manual_seed(5)
                res[fancy] = sober_b200.recombination(X, Z, 200, kern, None, None)
    # same points; the weights agree to summation-order rounding (the first K1 pass on the SM-partitioned stream picks its
    # row-split count for the partition's SM count, so its partial sums are grouped differently)
    assert torch.equal(res[False][0], res[True][0]) and float((res[False][1] - res[True][1]).abs().max()) < 1e-10


def test_pipelined_upload_matches_device_resident_input(ops, cuda_device):
    """Candidates in pinned host memory are copied in row chunks on a side stream while the first K1 pass consumes the
    chunks that have landed (_rchq.run / _ops.upload_chunks).  Same selection as with the candidates already on the
    device, weights to rounding (the chunked group sums differ in summation order only), result written back into the
    caller's host weight vector."""
    import sober_b200
    from sober_b200 import _nystrom
    g = torch.Generator().manual_seed(11)
    N, d, L, b = 300_000, 6, 300, 40
    X = torch.rand(N, d, dtype=torch.float64, generator=g)
    mu = torch.rand(N, dtype=torch.float64, generator=g) + 0.1
    mu /= mu.sum()
    Z = X[torch.randperm(N, generator=g)[:L]].clone().to(cuda_device)
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("matern", [0.7] * d, 1.3).to(cuda_device)), mode="kernel")
    R = torch.randn(L, b - 1, dtype=torch.float64, generator=g).to(cuda_device)
    out = {}
    for where in ("device", "pinned"):
        Xin = X.to(cuda_device) if where == "device" else X.clone().pin_memory()
        w_in = mu.clone().to(cuda_device) if where == "device" else mu.clone().pin_memory()
        _nystrom._injected_test_matrix = R
        try:
            with warnings.catch_warnings(), sober_b200.configure(mode="fast"):
                warnings.simplefilter("ignore")
                idx, w = sober_b200.recombination(Xin, Z, b, kern, None, None, init_weights=w_in)
        finally:
            _nystrom._injected_test_matrix = None
        out[where] = (idx.cpu(), w.cpu(), w_in.cpu())
    assert torch.equal(out["device"][0], out["pinned"][0])
    assert float((out["device"][1] - out["pinned"][1]).abs().max()) < 1e-9
    assert float((out["device"][2] - out["pinned"][2]).abs().max()) < 1e-9
    assert abs(float(out["pinned"][2].sum()) - 1.0) < 1e-9


def test_fbgp_kernel_through_the_generic_path_on_gpu(ops, cuda_device):
    """``FullyBayesianGP.marginal_predictive_covariance`` (SOBER/FBGP/_fully_Bayesian_gp.py:354-371, restated in
    oracle/fbgp.py) is an opaque 2-D-only callable: tiles of it are evaluated on the device and reduced by
    sober_group_accumulate_gram; same selection as the CPU oracle with the broadcast form of the formula."""
    import sober_b200
    from _cases import fbgp_case
    X, Z, mu, models, w_qd, ofb = fbgp_case()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        idx_o, w_o = oracle.recombination(X, Z, 8, ofb.FullyBayesianGP(models, w_qd, strict=False).marginal_predictive_covariance,
                                          None, None, init_weights=mu.clone())
    Xg, Zg, mug, models_g, wq_g, _ = fbgp_case(cuda_device)
    kern = ofb.FullyBayesianGP(models_g, wq_g, strict=True).marginal_predictive_covariance
    with warnings.catch_warnings(), sober_b200.configure(mode="parity"):
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        idx, w = sober_b200.recombination(Xg, Zg, 8, kern, None, None, init_weights=mug)
    assert len(idx) <= 8 and bool((w > 0).all()) and abs(float(w.sum()) - 1.0) < 1e-12
    # parity mode on another device than the oracle: the range finder's normals differ (CPU vs CUDA generator), so the
    # comparison is on the invariants and on the quadrature error of the selected batch
    kern_cpu = ofb.FullyBayesianGP(models, w_qd, strict=False).marginal_predictive_covariance
    mmd_o = oracle.mmd_squared(kern_cpu, X, mu, idx_o, w_o)
    mmd_g = oracle.mmd_squared(kern_cpu, X, mu, idx.cpu(), w.cpu())
    assert float(mmd_g) <= 10 * float(mmd_o) + 1e-12
