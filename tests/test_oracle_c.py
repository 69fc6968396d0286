"""The plain-C restatement (oracle/tdq_oracle.c) agrees BITWISE with the torch-CPU oracle that is pinned to the
reference's golden vectors: the order of roundings the CUDA kernels reproduce is written down twice."""
import ctypes as C

import pytest
import torch

from oracle import build_c
from oracle import ode_oracle as O


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(build_c.build())


def _rand(n, dtype, seed, scale=1.0):
    return (torch.randn(n, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale).to(dtype)


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("dtype,sfx,ct", [(torch.float32, "f32", C.c_float), (torch.float64, "f64", C.c_double)])
@pytest.mark.parametrize("method", ["dopri5", "dopri8", "tsit5"])
def test_combine_error_fit_eval(lib, dtype, sfx, ct, method):
    n, dt = 1003, 0.0371
    tab = O.tableau(method)
    cast = O._cast_tableau(tab, dtype)
    S = tab["n_stages"]
    y0 = _rand(n, dtype, 1)
    ks = [_rand(n, dtype, 10 + j, 1e-2) for j in range(S + 1)]
    dtT = torch.tensor(dt, dtype=torch.float64).to(dtype)
    out = torch.empty(n, dtype=dtype)

    def c_weighted(vec_ks, coefs, base):
        idx = [j for j, c in enumerate(coefs.tolist()) if c != 0.0]
        kp = (C.c_void_p * len(idx))(*[vec_ks[j].data_ptr() for j in idx])
        cc = torch.stack([coefs[j] for j in idx]).contiguous()
        getattr(lib, "orc_combine_" + sfx)(_p(out), _p(base) if base is not None else None, kp, _p(cc), len(idx), C.c_size_t(n))
        return out.clone()
    for i in range(S):                                                       # rk_common.py:79
        coefs = cast["beta"][i] * dtT
        assert torch.equal(c_weighted(ks, coefs, y0), y0 + O._weighted(ks, coefs)), (method, i)
    err = O._weighted(ks, dtT * cast["c_err"])                               # rk_common.py:89
    assert torch.equal(c_weighted(ks, dtT * cast["c_err"], None), err)
    y1 = y0 + O._weighted(ks[:S], cast["beta"][S - 1] * dtT)
    fn = getattr(lib, "orc_error_sumsq_" + sfx)
    fn.restype = C.c_double
    got = fn(_p(err), _p(y0), _p(y1), C.c_double(1e-5), C.c_double(1e-7), C.c_size_t(n))
    tol = torch.tensor(1e-7, dtype=torch.float64) + torch.tensor(1e-5, dtype=torch.float64) * torch.max(y0.abs(), y1.abs())
    q = err / tol
    want = float((q * q).double().sum())
    assert abs(got - want) <= 1e-12 * want
    # interpolant (interp.py:1-48)
    ymid = y0 + O._weighted(ks, dtT * cast["c_mid"])
    coeffs = O.interp_fit(y0, y1, ks, torch.tensor(dt, dtype=torch.float64), cast)
    bufs = [torch.empty(n, dtype=dtype) for _ in range(5)]
    getattr(lib, "orc_interp_fit_" + sfx)(*[_p(b) for b in bufs], _p(y0), _p(y1), _p(ymid), _p(ks[0]), _p(ks[-1]),
                                          ct(float(dtT)), C.c_size_t(n))
    for got_c, want_c in zip(bufs, coeffs):
        assert torch.equal(got_c, want_c)
    t0, t1, tq = 0.5, 0.5 + dt, 0.5 + 0.37 * dt
    getattr(lib, "orc_interp_eval_" + sfx)(_p(out), *[_p(b) for b in bufs], C.c_double(t0), C.c_double(t1), C.c_double(tq),
                                           C.c_size_t(n))
    f64 = lambda v: torch.tensor(v, dtype=torch.float64)
    assert torch.equal(out, O.interp_eval(coeffs, f64(t0), f64(t1), f64(tq)))


@pytest.mark.parametrize("dtype,sfx,ct", [(torch.float32, "f32", C.c_float), (torch.float64, "f64", C.c_double)])
def test_rk4_and_controller(lib, dtype, sfx, ct):
    n = 517
    y0, k1, k2, k3, k4 = [_rand(n, dtype, s) for s in range(5)]
    h = torch.tensor(0.037, dtype=dtype)
    wants = [y0 + h * k1 * (1 / 3), y0 + h * (k2 - k1 * (1 / 3)), y0 + h * (k1 - k2 + k3),
             y0 + (k1 + 3 * (k2 + k3) + k4) * h * 0.125]
    out = torch.empty(n, dtype=dtype)
    for which, want in enumerate(wants, 1):
        getattr(lib, "orc_rk4_stage_" + sfx)(which, _p(out), _p(y0), _p(k1), _p(k2), _p(k3), _p(k4), ct(float(h)),
                                             C.c_size_t(n))
        assert torch.equal(out, want), which
    lib.orc_optimal_step.restype = C.c_double
    f64 = lambda v: torch.tensor(v, dtype=torch.float64)
    for ratio in (0.0, 1e-9, 0.3, 0.999, 1.0, 1.7, 250.0):
        for order in (5, 8):
            want = O.optimal_step(f64(0.1), f64(ratio), f64(0.9), f64(10.0), f64(0.2), order).clamp(f64(0.0), f64(float("inf")))
            got = lib.orc_optimal_step(C.c_double(0.1), C.c_double(ratio), C.c_double(0.9), C.c_double(10.0),
                                       C.c_double(0.2), order, C.c_double(0.0), C.c_double(float("inf")))
            assert abs(got - float(want)) <= 4e-16 * float(want), (ratio, order)
