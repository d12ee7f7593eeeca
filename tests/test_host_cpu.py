"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header declares, JSON
comment stripping, lowering of the reference's objects to the flat problem descriptor, dispatcher errors."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "pybitup_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pybitup_b200 import _capi
    names = _header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libpybitup_b200.so does not export %s" % n
    # the ctypes binding declares a signature for each of them
    assert set(names) == set(_capi.EXPORTED), set(names) ^ set(_capi.EXPORTED)
    assert _capi.lib.pbu_abi_version() == 1
    n, listing = _capi.list_kernels()
    assert n > 100 and "model=1 P=4 d=4" in listing


def test_ctypes_structs_match_header_layout(tmp_path):
    """The ctypes mirrors against the C compiler's view of include/pybitup_b200.h: sizes and every field offset."""
    import subprocess
    from pybitup_b200 import _capi
    structs = {"pbu_problem": _capi.pbu_problem, "pbu_sampler_cfg": _capi.pbu_sampler_cfg, "pbu_run_outputs": _capi.pbu_run_outputs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pybitup_b200.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in st._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    seen = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, st in structs.items():
        assert int(seen[name]) == ctypes.sizeof(st), name
        for fname, _ in st._fields_:
            assert int(seen["%s.%s" % (name, fname)]) == getattr(st, fname).offset, (name, fname)


def test_strip_json_comments():
    from pybitup_b200.solve_problem import strip_json_comments
    txt = '/* head */ {"a": 1, // one\n "b": "x // not a comment", "c": "q\\"/*no*/", /* gone */ "d": [1e2]}'
    assert json.loads(strip_json_comments(txt)) == {"a": 1, "b": "x // not a comment", "c": 'q"/*no*/', "d": [100.0]}


class _Heat:
    pass


def _posterior(model_cls_attrs=None):
    from pybitup_b200 import bayesian_inference as bi, distributions
    attrs = {"device_model": "heat"}
    attrs.update(model_cls_attrs or {})
    M = type("HeatConduction", (bi.Model,), attrs)
    m = M()
    m.param = m.param_nom = np.array([0.95, 0.95, 2.37, 21.29, -18.41, 0.00191])
    m.param_names = ["a", "b", "k", "T_amb", "phi", "h"]
    m.unpar_name = ["phi", "h"]
    m.x = np.arange(10.0, 70.0, 4.0)
    data = {"heat": bi.Data("T", m.x, {0: np.linspace(90, 25, 15)}, np.full(15, 0.26))}
    prior = distributions.set_probability_dist(["Mixture"], [["Uniform", "Uniform"], [[-100, 0.0], [1e-5, 1.0]]], 2)
    like = bi.Likelihood(data, {"heat": m}, 1.0)
    return bi.BayesianPosterior(prior, like, m, np.array([-18.41, 0.00191]))


def test_lowering_of_reference_objects():
    post = _posterior()
    spec = post.lower()
    assert spec.model == "heat" and spec.d == 2 and spec.n_points == 15
    # substring name matching of Model.run_model: "h" matches inside "phi h", "a" inside... no
    assert list(spec.unc_idx) == [4, 5]
    assert np.array_equal(spec.lb, [-100, 1e-5]) and np.array_equal(spec.ub, [0.0, 1.0])
    assert post.prior.compute_log_value([-18.0, 0.5]) == pytest.approx(np.log(1 / 100) + np.log(1 / (1 - 1e-5)))
    assert post.prior.compute_log_value([1.0, 0.5]) == -np.inf
    s, keep = spec._c_struct()
    assert s.n_param == 6 and s.n_unc == 2 and s.n_design_cols == 1


def test_substring_quirk_orders_by_model_parameter_order():
    from pybitup_b200 import bayesian_inference as bi
    m = bi.Model()
    m.param_nom = np.zeros(4)
    m.param_names = ["k", "alpha", "b", "phi"]
    m.unpar_name = ["phi", "alpha"]                  # prior order differs from model order
    # n-th uncertain VALUE goes to the n-th matching NAME in param_names order (bayesian_inference.py:356-366)
    assert list(m.uncertain_index()) == [1, 3]
    m.unpar_name = ["phi h"]                         # "h" is a substring of the joined names
    m.param_names = ["h", "phi", "z"]
    assert list(m.uncertain_index()) == [0]


def test_unsupported_problems_fail_loudly():
    from pybitup_b200 import _capi
    post = _posterior({"device_model": None})
    with pytest.raises(_capi.UnsupportedProblem):
        post.lower()
    # log sampling with the det_jac hook left at 1 is a registered change of variables (kind 1, constant log det_jac) ...
    spec = _posterior({"parametrization_backward": lambda self, Y=1, P=1: np.exp(Y)}).lower()
    assert list(spec.param["kind"]) == [1] * spec.d and np.all(spec.param["c"] == 1.0) and np.all(spec.param["ldj_s"] == 0.0)
    assert _posterior({}).lower().param is None
    # ... a map the device does not know is refused
    post = _posterior({"parametrization_backward": lambda self, Y=1, P=1: np.asarray(Y) ** 3})
    with pytest.raises(_capi.UnsupportedProblem):
        post.lower()
    from pybitup_b200 import distributions
    with pytest.raises(ValueError):
        distributions.set_probability_dist(["Gamma"], [1, 2], 1)


def test_sampler_dispatch_errors():
    from pybitup_b200.inference_problem import Sampler
    with pytest.raises(ValueError, match='Algorithm "NUTS" unknown.'):
        Sampler({}, None, {"name": "NUTS", "n_iterations": 10})
    s = Sampler({}, None, {"name": "RWMH", "n_iterations": 1e3, "n_chains": 8, "seed": 3,
                           "proposal": {"covariance": {"type": "diag", "value": [1.0, 4.0]}}})
    assert s.n_iterations == 1000 and s.opt_arg["save_eval_freq"] == 10 and s.opt_arg["n_chains"] == 8
    assert np.array_equal(s.proposal_cov, np.diag([1.0, 4.0]))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pybitup_b200.engine import ChainEngine, SamplerConfig
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ChainEngine(_posterior().lower(), SamplerConfig(), 1)


def test_percentile_lerp_matches_numpy():
    from pybitup_b200.propagation import lerp_percentile, percentile_ranks
    rng = np.random.default_rng(0)
    for n in (2, 5, 40, 1001):
        v = np.sort(rng.standard_normal(n))
        for q in (2.5, 50.0, 97.5, 100.0, 0.0):
            lo, hi = percentile_ranks(n, q)
            assert lerp_percentile(n, q, v[lo], v[hi]) == np.percentile(v, q)


def test_hessian_fd_matches_reference_golden():
    """compute_hessian_fd vs the reference's computeHessianFD (recorded by tests/golden/gen_golden.py --hessian)."""
    from pybitup_b200.metropolis_hastings_algorithms import compute_hessian_fd
    g = np.load(os.path.join(ROOT, "tests", "golden", "_hessian_fd.npz"))
    A, x = g["A"], g["x"]
    f = lambda X: np.array([float(-0.5 * z @ A @ z + 0.1 * np.sin(z).sum() + 0.05 * z[0] * z[1] * z[2]) for z in X])   # noqa: E731
    for full in (True, False):
        for pos in (True, False):
            H = compute_hessian_fd(f, x, eps=0.001, compute_all_element=full, positive_diag=pos)
            assert np.array_equal(H, g["H_%d_%d" % (full, pos)])


def test_near_pd_matches_reference_golden():
    """nearPD vs the reference's (recorded by tests/golden/gen_golden.py --nearpd), bit for bit."""
    from pybitup_b200.metropolis_hastings_algorithms import nearPD
    g = np.load(os.path.join(ROOT, "tests", "golden", "_near_pd.npz"))
    for k in ("indefinite", "pd", "scaled"):
        Y = nearPD(g["A_" + k])
        assert np.array_equal(Y, g["Y_" + k]), k
        assert np.allclose(np.diag(Y), np.abs(np.diag(g["A_" + k])))          # unit diagonal scaled back by sigma^2


def test_gradient_sampler_preconditioner_options(tmp_path):
    """covariance.value: Matrix / from_file are read on the host, PSD is refused loudly, anything else is the
    reference's ValueError (mha.py:708)."""
    from pybitup_b200 import _capi
    from pybitup_b200.metropolis_hastings_algorithms import ito_SDE
    C = np.array([[2.0, 0.3], [0.3, 1.0]])
    np.savetxt(tmp_path / "cov.csv", C, delimiter=",")
    io = {"fileID": {}, "path": {"cwd": tmp_path, "out_folder": tmp_path}}
    base = {"starting_it": 10, "update_it": 5}

    def mk(value, **extra):
        s = ito_SDE.__new__(ito_SDE)
        s.IO_util, s.nIterations = io, 100
        s._setup_gradient(dict(base, value=value, **extra), "Analytical")
        return s
    assert np.array_equal(mk("from_file").V, C)
    assert np.array_equal(mk("Matrix", matrix_value=C.tolist()).V, C)
    assert mk("nearPD").end_it == 101
    with pytest.raises(_capi.UnsupportedProblem):
        mk("PSD")
    with pytest.raises(ValueError, match="Unknown matrix type"):
        mk("Cholesky")


def test_device_math_against_libm(tmp_path):
    """pbu_math.cuh (the exp / log / table-driven exp2 / log2 the kernels use, host-compilable): tools/test_math.cu
    holds them to <= 2 ulp of libm over 2e7 random arguments, a whole Arrhenius stage to 3e-14, and pbu_exp_full (the
    acceptance ratio's exp) to its edge cases (+-inf, NaN, the overflow / underflow cut-offs)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "test_math")
    subprocess.check_call([nvcc, "-O2", "-o", exe, os.path.join(ROOT, "tools", "test_math.cu")], stderr=subprocess.DEVNULL)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
