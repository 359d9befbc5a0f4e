"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the front-end
builds the right config, product-side statistics agree with the oracle, and the multi-process (gloo, world_size 2)
sharding path reduces to the single-process result."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest

import oracle as o
from util import relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "gamc_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gamc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(gamc):
    lib = gamc._lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libgamc_b200.so does not export {s}"
        assert s in gamc._lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(gamc._lib.SIGNATURES) == set(syms)
    assert lib.gamc_version().decode().endswith("sm_100a")


def test_struct_layouts_match_the_header(gamc, tmp_path):
    """Compile the header with gcc (plain C: the ABI must not need C++) and compare struct sizes / offsets with ctypes."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gamc_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(gamc_config), sizeof(gamc_state_scalars), '
                   'offsetof(gamc_config, nsteps), offsetof(gamc_config, am_mode), offsetof(gamc_state_scalars, chain_saved));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    C, S = gamc._lib.Config, gamc._lib.StateScalars
    assert [int(x) for x in out] == [ctypes.sizeof(C), ctypes.sizeof(S), C.nsteps.offset, C.am_mode.offset, S.chain_saved.offset]


def test_no_cpu_fallback(gamc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gamc._lib.CudaError):
        gamc.Engine(0)


def test_frontend_config_mirrors_reference_defaults(gamc):
    s = gamc.GAMC()
    assert (s.driftstep, s.minorscale, s.c, s.t0, s.update.name) == (1.0, 1.0, 0.05, 3, "rand_update!")
    for bad in (dict(driftstep=0.0), dict(minorscale=0.0), dict(c=-0.1), dict(t0=0)):
        with pytest.raises(AssertionError):
            gamc.GAMC(**bad)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(verbose=True), gamc.AcceptanceRateMCTuner(0.35))
    assert tuner.verbose and repr(tuner) == "GAMCTuner"
    cfg = gamc.frontend.make_config(gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01), tuner,
                                    gamc.BasicMCRange(110000, 10000))
    assert (cfg.schedule, cfg.sched_params[0], cfg.total_tuner_kind, cfg.targetrate, cfg.score_k, cfg.period) == (5, 10.0, 1, 0.35, 7.0, 100)
    assert (cfg.nsteps, cfg.burnin, cfg.verbose_am, cfg.verbose_smmala) == (110000, 10000, 1, 0)
    assert "drift step = 0.02" in repr(gamc.GAMC(driftstep=0.02))
    # decay functions of the front-end are the reference's (src/stats/schedule.jl:1-7)
    for f in ("exp_decay", "lin_decay", "pow_decay", "quad_decay"):
        assert getattr(gamc, f)(7, 90) == getattr(o, f)(7, 90)


def test_product_stats_match_oracle(gamc):
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.standard_normal((3, 5000)), axis=1) * 0.01 + rng.standard_normal((3, 5000))
    assert relerr(gamc.ess(x), o.ess(x, maxlag=4999)) < 1e-8
    s = gamc.summary(x, rng.random(5000) < 0.3, 2.0)
    assert s["efficiency"] == pytest.approx(s["ess"].min() / 2.0)


def test_shard_rows_partition():
    from gamc_b200.distributed import shard_rows
    for N, w in ((10, 3), (1 << 20, 8), (7, 8)):
        edges = [shard_rows(N, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == N
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    import oracle as oo
    from gamc_b200.distributed import shard_rows, broadcast_unique_id
    N, p = 2001, 6
    tt = oo.logistic.synth_theta_true(9, p)
    r0, r1 = shard_rows(N, world, rank)
    X, y, _ = oo.synth_logistic(9, N, p, row0=r0, nrows=r1 - r0, theta_true=tt)      # this rank's shard only
    ll, g, G = oo.logistic_partials(tt, X, y)
    buf = torch.from_numpy(np.concatenate([[ll], g, G.ravel()]))
    dist.all_reduce(buf)                                                            # the data-path collective (sum)

    class FakeEngine:                                                               # unique-id plumbing without NCCL
        @staticmethod
        def comm_unique_id():
            return np.arange(128, dtype=np.uint8)
    uid = broadcast_unique_id(FakeEngine, rank, world)
    q.put((rank, buf.numpy().copy(), uid.copy()))
    dist.destroy_process_group()


def test_gloo_world2_sharded_reduce_matches_whole():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    N, p = 2001, 6
    X, y, tt = o.synth_logistic(9, N, p)
    ll, g, G = o.logistic_partials(tt, X, y)
    whole = np.concatenate([[ll], g, G.ravel()])
    for rank, buf, uid in res:
        assert relerr(buf, whole) < 1e-12
        assert np.array_equal(uid, np.arange(128, dtype=np.uint8))
    assert np.array_equal(res[0][1], res[1][1])      # every rank holds the bit-identical reduced payload


def test_csv_pipeline_roundtrip(gamc, tmp_path):
    """chainNN.csv / diagnosticsNN.csv / times.csv / summary.* in the reference's formats
    (examples/logistic_regression/src/GAMC/simulations.jl:81-95, GAMC/table.jl:21-53)."""
    rng = np.random.default_rng(1)
    d = str(tmp_path)
    chains = [rng.standard_normal((4, 800)) for _ in range(3)]
    accs = [rng.random(800) < 0.3 for _ in range(3)]
    for i in range(3):
        gamc.write_chain_csv(d, i + 1, chains[i], accs[i])
    gamc.write_run_csvs(d, [1.0, 2.0, 3.0], [0.1, 0.2, 0.3], [10, 11, 12])
    res = gamc.table(d, 3)
    assert res["rate"] == pytest.approx(np.mean([a.mean() for a in accs]))
    assert relerr(res["ess"], np.mean([gamc.ess(c) for c in chains], axis=0)) < 1e-12
    assert res["time"] == 2.0 and res["efficiency"] == pytest.approx(res["ess"].min() / 2.0)
    assert np.array_equal(np.loadtxt(os.path.join(d, "chain02.csv"), delimiter=","), chains[1])
    assert open(os.path.join(d, "diagnostics01.csv")).readline().strip() in ("true", "false")
    assert len(open(os.path.join(d, "summary.txt")).read().split(" & ")) == 1 + 4 + 2


def test_rv_data_generation_and_csv(tmp_path):
    """the front-end's host-side Keplerian forward model (data generation only), the reference's truth vectors
    (utils_ex.jl:48-68) and its `time,obs,sigma` CSV format, against the oracle and the reference's shipped data file."""
    import gamc_b200 as g
    from oracle import targets as T
    for mine, ref in ((g.make_param_true_ex1(), T.make_param_true_ex1()), (g.make_param_true_ex2(), T.make_param_true_ex2())):
        assert np.array_equal(mine, ref)
        m = g.RVModel.simulate(mine, num_obs=300, seed=5)
        assert m.nplanets == len(mine) // 5 and np.all(np.diff(m.times) >= 0) and m.times.max() <= 730.5
        v = g.rv_model_velocity(mine, m.times)
        assert np.max(np.abs(v - T.RVModel(m.times, m.obs, m.sigma_obs).model_rv(ref))) <= 1e-12
        resid = (m.obs - v) / m.sigma_obs
        assert 0.85 < resid.std() < 1.15 and abs(resid.mean()) < 0.2          # obs = model + sigma * N(0, 1)
        path = tmp_path / "rv.csv"
        m.to_csv(path)
        back = g.RVModel.from_csv(path, m.nplanets)
        assert np.array_equal(back.times, m.times) and np.array_equal(back.obs, m.obs) and np.array_equal(back.sigma_obs, m.sigma_obs)
    with pytest.raises(ValueError):
        g.rv_model_velocity(np.array([3.0, 3.0, 0.8, 0.8, 1.0, 0.0]), np.array([1.0]))


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every entry point of include/gamc_b200.h to the reference interface it replaces."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "gamc_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(gamc_[a-z_0-9]+)\s*\(", header)))
    assert len(names) >= 37
    assert [n for n in names if n not in doc] == []


def test_metric_planner_invariants(tmp_path):
    """Host planner of the metric kernel (metric_plan.h) for every p in 1..512: each lower-triangle block pair is exactly one task,
    each column block has exactly one diagonal task, slot / task / CTA tables are consistent (tests/cpp/plan_check.cu; only host
    code runs, so no GPU is needed)."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "plan_check")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(root, "gamcsampler.jl_b200", "csrc"),
                    os.path.join(root, "tests", "cpp", "plan_check.cu"), "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]


def test_i8_planner_invariants(tmp_path):
    """Host planner of the integer-tensor-core metric (metric_i8_plan.h): every digit product of every output tile is issued exactly
    once into the right group slot, the int32 head room bounds the item length, every kind covers every row chunk exactly once in
    row order (tests/cpp/i8_plan_check.cu; only host code runs)."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "i8_plan_check")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(root, "gamcsampler.jl_b200", "csrc"),
                    os.path.join(root, "tests", "cpp", "i8_plan_check.cu"), "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]


def test_integer_metric_kernel_uses_tcgen05_without_issue_loops():
    """SASS of the built library (cuobjdump, no GPU needed): k_metric_i8 issues tcgen05.mma kind::i8 (UTCIMMA), commits to mbarriers
    (UTCBAR), reads TMEM (LDTM), is fed by TMA (UTMALDG) and adds its results with 64-bit integer reductions (REDG ... .64); and the MMA /
    TMA issue compiles to straight-line code -- no per-lane `BRA.U.ANY` loop around the uniform-register moves (what `if (lane == 0)`
    instead of `elect.sync` produced: three times slower issue than the tensor pipe needs)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "gamcsampler.jl_b200", "libgamc_b200.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library not available")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    body, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = "k_metric_i8" in line
        elif on:
            body.append(line)
    sass = "\n".join(body)
    assert sass.count("UTCIMMA") >= 16 and sass.count("UTCBAR") >= 2 and sass.count("LDTM") >= 1 and sass.count("UTMALDG") >= 1
    assert "REDG.E.ADD.64" in sass
    assert sass.count("BRA.U.ANY") == 0


def test_i8_scheme_golden():
    """The committed fixture of the integer metric route (tests/golden/i8_scheme.json, made by make_golden.py): the NumPy statement
    still produces the same column exponents, digit planes (bit for bit) and metric, and that metric is the FP64 oracle's to 1e-12."""
    import json
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import i8_ref
    import oracle as o
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "i8_scheme.json")))
    X, y, tt = o.synth_logistic(d["seed"], d["N"], d["p"])
    planes, e = i8_ref.digit_planes(X, np.array(d["w"]), d["slices"])
    assert np.array_equal(e, np.array(d["column_exponents"]))
    assert np.array_equal(planes, np.array(d["planes"], dtype=np.int8))
    G = i8_ref.metric_from_planes(planes, e, d["smax"])
    assert np.array_equal(G, np.array(d["G_lik"]))
    Gl = o.logistic_partials(np.array(d["theta"]), X, y)[2]
    assert np.max(np.abs(G - Gl)) <= 1e-12 * np.max(np.abs(Gl))


def test_i8_scheme_accuracy_numpy():
    """NumPy statement of the integer route (tests/i8_ref.py) against extended precision: 6 digit planes with the digit products
    a + b <= 7 reproduce the logistic metric to 1e-13 of its largest entry; the balanced digits reconstruct the fixed-point value
    exactly; the error grows as advertised with fewer planes."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import i8_ref
    import oracle as o
    X, y, tt = o.synth_logistic(20260101, 3000, 48)
    sig = 1.0 / (1.0 + np.exp(-(X @ tt)))
    w = sig * (1.0 - sig)
    Xl = X.astype(np.longdouble)
    G0 = np.array((Xl.T * w.astype(np.longdouble)) @ Xl, dtype=np.float64)
    rel = lambda A: float(np.max(np.abs(A - G0)) / np.max(np.abs(G0)))
    assert rel(i8_ref.metric_i8(X, w, 6, 7)) <= 3e-13
    assert rel(i8_ref.metric_i8(X, w, 6, 8)) <= 1e-14
    assert 1e-13 < rel(i8_ref.metric_i8(X, w, 5, 6)) <= 1e-11
    assert rel(i8_ref.metric_i8(X, w, 4, 5)) <= 2e-8
    planes, e = i8_ref.digit_planes(X, w, 6)
    q = sum(planes[s].astype(np.int64) << (8 * (5 - s)) for s in range(6))          # sum_a d_a 256^(k - a)
    S = np.sqrt(w)[:, None] * X
    d = q.T - np.rint(S * np.ldexp(1.0, (46 - e).astype(np.int64))[None, :]).astype(np.int64)
    assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= 1e-3 * d.size       # (the bias is added before the rounding: a tie can go the other way)
    assert np.abs(planes[0]).max() <= 65 and planes.min() >= -128 and planes.max() <= 127


def test_julia_binding_covers_every_entry_point():
    """The Julia `ccall` front-end (not executable in this image) binds every symbol include/gamc_b200.h declares, and its C
    structs list the same fields, in the same order, as the header."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "gamc_b200.h")).read()
    jl = open(os.path.join(root, "gamcsampler.jl_b200", "julia", "GAMCSamplerB200.jl")).read()
    names = sorted(set(re.findall(r"\b(gamc_[a-z_0-9]+)\s*\(", header)))
    bound = set(re.findall(r"\(:(gamc_[a-z_0-9]+), libgamc\)", jl))
    assert [n for n in names if n not in bound] == []
    for cname, jname in (("gamc_config", "CConfig"), ("gamc_state_scalars", "CState")):
        body = [c for c in header.split("typedef struct {") if re.search(r"\} " + cname + ";", c)][-1].split("} " + cname + ";")[0]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        cfields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(",")[0:1] + decl.split(",")[1:]:
                cfields.append(re.sub(r"\[.*?\]", "", part.strip().split()[-1]))
        jbody = re.search(r"struct " + jname + r"\n(.*?)\nend", jl, re.S).group(1)
        jfields = re.findall(r"([A-Za-z_0-9]+)::", jbody)
        assert jfields == cfields, (jname, jfields, cfields)


def _build_chain50(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "gamcsampler.jl_b200")
    if not os.path.exists(os.path.join(libdir, "libgamc_b200.so")):
        pytest.skip("library not built")
    exe = str(tmp_path / "chain50")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "chain50.c"),
                    "-L", libdir, "-lgamc_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe], check=True, capture_output=True)
    return exe


def test_c_program_links_against_the_library(tmp_path):
    """tests/cpp/chain50.c (plain C99, no CUDA headers, no Python) compiles against include/gamc_b200.h and links against the
    shared library: the header really is a C header and every symbol it uses is exported."""
    _build_chain50(tmp_path)


@pytest.mark.gpu
def test_c_program_runs_a_chain(tmp_path):
    """The same program executed: target, evaluation, sampler, 50 transitions through `gamc_iterate`, monitors, a user schedule
    function installed from C."""
    import subprocess
    exe = _build_chain50(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("CHAIN50 OK"), (out.stdout, out.stderr)
