"""GPU tests of the integer-tensor-core metric route (gamcsampler.jl_b200/csrc/metric_i8.cuh; gamc_set_metric_route): the digit planes
and the exact digit products against the NumPy statement of the scheme (tests/i8_ref.py), the metric against the FP64 oracle,
trajectories against the oracle on a shared random tape, determinism, the FP64 route beside it."""
import ctypes as C

import numpy as np
import pytest

import i8_ref
from util import relerr, example_config

pytestmark = pytest.mark.gpu


def _oracle():
    import oracle
    return oracle


@pytest.fixture()
def i8_engine(engine):
    engine.set_metric_route("i8")
    yield engine
    engine.set_metric_route("auto")


def _debug_read(engine):
    lib = engine.lib
    lib.gamc_debug_i8_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    lib.gamc_debug_i8_read.restype = C.c_int
    info = (C.c_longlong * 4)()
    assert lib.gamc_debug_i8_read(engine.h, 2, info, 32) == 0, "the integer route is not active"
    k, smax, ldn, kc = (int(v) for v in info)
    p = engine.p
    buf = np.empty(k * p * ldn, dtype=np.int8)
    assert lib.gamc_debug_i8_read(engine.h, 0, buf.ctypes.data_as(C.c_void_p), buf.nbytes) == 0
    planes = buf.reshape(ldn // kc, k, p, kc).transpose(1, 2, 0, 3).reshape(k, p, ldn)
    cg = np.empty(p)
    assert lib.gamc_debug_i8_read(engine.h, 1, cg.ctypes.data_as(C.c_void_p), cg.nbytes) == 0
    return k, smax, planes, cg


@pytest.mark.parametrize("N,p", [(200, 4), (33, 1), (1000, 37), (4096, 256), (4097, 300), (777, 512), (36 * 50 + 1, 64), (70000, 130)])
def test_i8_eval_parity(i8_engine, N, p):
    """Log-target, gradient and metric of the integer route against the FP64 oracle (1e-10 norm-wise, as for the FP64 route; the
    scheme itself is good for 1e-12), the digit products against NumPy on the device's own digit planes (exact integers), the digit
    planes against the NumPy statement (identical up to the last digit where the device's and NumPy's sigmoid differ by an ulp)."""
    eng = i8_engine
    o = _oracle()
    X, y, tt = o.synth_logistic(20260101, N, p)
    theta = tt + 0.05 * np.random.default_rng(7).standard_normal(p)
    eng.target_logistic(X, y, 100.0)
    lt, g, G = eng.eval_upto_tensor(theta)
    lt0, g0, G0 = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    assert abs(lt - lt0) <= 1e-10 * abs(lt0) and relerr(g, g0) <= 1e-10
    assert relerr(G, G0) <= 1e-10
    assert np.array_equal(G, G.T)
    ll, gl, Gl = eng.eval_partials(theta)
    ll0, gl0, Gl0 = o.logistic_partials(theta, X, y)
    assert relerr(Gl, Gl0) <= 2e-12, "the 6-plane / a + b <= 7 scheme is good for 1e-13 of the largest entry"
    k, smax, planes, cg = _debug_read(eng)
    assert (k, smax) == (6, 7)
    assert not planes[:, :, N:].any()
    planes = planes[:, :, :N]
    # exactness of the tensor-core path: the same integers summed by NumPy
    e = np.log2(cg).astype(np.int64)
    assert np.array_equal(np.ldexp(1.0, e), cg) and np.array_equal(e, i8_ref.column_exponents(X))
    Gn = i8_ref.metric_from_planes(planes, e, smax)
    assert relerr(Gl, Gn) <= 4e-16
    # digit planes: NumPy statement with the host's sigmoid
    sig = 1.0 / (1.0 + np.exp(-(X @ theta)))
    ref, _ = i8_ref.digit_planes(X, sig * (1.0 - sig), k)
    diff = planes.astype(np.int32) - ref.astype(np.int32)
    assert np.count_nonzero(diff[:k - 1]) == 0 or np.count_nonzero(diff) <= 1e-3 * diff.size
    assert np.abs(diff[k - 1][np.abs(diff[k - 1]) < 128]).max(initial=0) <= 2      # last digit: off by the rounding of w (wrap-arounds carry into the next plane)


def test_i8_rows_past_n_are_zero(i8_engine):
    eng = i8_engine
    o = _oracle()
    X, y, tt = o.synth_logistic(3, 130, 9)
    eng.target_logistic(X, y, 100.0)
    eng.eval_upto_tensor(tt)
    _, _, planes, _ = _debug_read(eng)
    assert planes.shape[2] == 256 and not planes[:, :, 130:].any()


def test_i8_deterministic_and_routes_agree(engine):
    """Two evaluations give identical bits (the schedule of the GEMM kernel is dynamic, the integer sums are not); the FP64 route on
    the same data agrees to 1e-12; fewer digit planes give the advertised, larger error."""
    o = _oracle()
    X, y, tt = o.synth_logistic(77, 20000, 96)
    theta = tt + 0.03 * np.random.default_rng(1).standard_normal(96)
    G0 = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)[2]
    try:
        engine.set_metric_route("i8")
        engine.target_logistic(X, y, 100.0)
        Ga = engine.eval_upto_tensor(theta)[2]
        Gb = engine.eval_upto_tensor(theta)[2]
        assert np.array_equal(Ga, Gb)
        engine.set_metric_route("f64")
        engine.target_logistic(X, y, 100.0)
        Gf = engine.eval_upto_tensor(theta)[2]
        assert relerr(Ga, Gf) <= 2e-12 and relerr(Gf, G0) <= 1e-12
        engine.set_metric_route("i8", slices=5, smax=6)
        engine.target_logistic(X, y, 100.0)
        G5 = engine.eval_upto_tensor(theta)[2]
        assert 1e-14 < relerr(G5, G0) <= 2e-11
        engine.set_metric_route("i8", slices=4, smax=5)
        engine.target_logistic(X, y, 100.0)
        G4 = engine.eval_upto_tensor(theta)[2]
        assert relerr(G4, G0) <= 5e-8
        with pytest.raises(Exception):
            engine.set_metric_route("i8", slices=9)
    finally:
        engine.set_metric_route("auto")


@pytest.mark.parametrize("am_mode", [0, 2, 3])
@pytest.mark.parametrize("N,p,nsteps", [(200, 4, 2000), (2000, 64, 1000), (6000, 256, 1000), (4000, 512, 400)])
def test_i8_trajectory_parity(gamc, i8_engine, N, p, nsteps, am_mode):
    """GAMC chains whose SMMALA steps use the integer-route metric: step for step against the oracle on a shared tape (identical
    accept decisions, values to 1e-8), including the BASELINE dimensions."""
    from test_gpu_parity import _run_both
    job, chain, osam, v, a, k = _run_both(gamc, i8_engine, N, p, nsteps, nsteps // 5, seed=42, am_mode=am_mode)
    assert np.array_equal(chain.diagnosticvalues.ravel(), a), "accept/reject decisions diverged"
    assert relerr(chain.value, v) <= 1e-8
    st = job.engine.state()
    assert st.updatetensorcount == osam.updatetensorcount and st.count == nsteps


def test_i8_reuse_tensor_chain_identical(gamc, i8_engine):
    """gamc_config.reuse_tensor skips the re-evaluation kernels of the integer route as well: identical chain."""
    o = _oracle()
    X, y, tt = o.synth_logistic(5, 3000, 40)
    chains = []
    for reuse in (False, True):
        sampler, tuner, rng_ = example_config(gamc, 300, 0, am_mode=2)
        sampler.reuse_tensor = reuse
        i8_engine.target_logistic(X, y, 100.0)
        job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": tt.copy()}, tuner=tuner, tape=gamc.RandomTape(300, 40, seed=8),
                              engine=i8_engine)
        job.run()
        chains.append(job.output().value.copy())
    assert np.array_equal(chains[0], chains[1])
