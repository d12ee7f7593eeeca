"""Shared test helpers: load golden fixtures, build oracle problems."""
import glob
import json
import os

import numpy as np

from oracle import mcmc_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix=""):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))
    return [n for n in names if not n.startswith("_")]          # "_*.npz": fixtures that are not sampler runs


def load_golden(case):
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz"), allow_pickle=False))
    g["algo"] = json.loads(str(g["algo_json"]))
    g["model"] = str(g["model"])
    if "param_json" in g:
        g["param"] = json.loads(str(g["param_json"]))
    return g


def device_param(param):
    """The engine's form of a fixture's re-parametrisation (ProblemSpec.param): -log(det_jac) = ldj_c0 + ldj_s . Y with
    det_jac = prod_i 1/(dX_i/dY_i) -- or the constant 1 when the recorded model left the hook alone."""
    if param is None:
        return None
    kind = np.asarray(param["kind"], dtype=np.int32)
    a, b, c = (np.asarray(param[k], dtype=np.float64) for k in ("a", "b", "c"))
    true_det = bool(param.get("det_jac", True))
    s = np.where(kind == 1, c, 0.0) if true_det else np.zeros(len(kind))
    c0 = float(np.sum(np.log(np.abs(np.where(kind == 1, b * c, b))))) if true_det else 0.0
    return dict(kind=kind, a=a, b=b, c=c, ldj_c0=c0, ldj_s=s, grad_ok=True)


def install_parametrization(model, param):
    """Element-wise change of variables as a user of the reference would code it in a bi.Model subclass (hooks of
    bayesian_inference.py:314-336 plus grad_det_jac, :631).  Works on the reference's Model and on pybitup_b200's mirror."""
    kind = np.asarray(param["kind"])
    a, b, c = (np.asarray(param[k], dtype=np.float64) for k in ("a", "b", "c"))
    true_det = bool(param.get("det_jac", True))

    def backward(Y=1, P=1):
        Y = np.asarray(Y, dtype=np.float64)
        return np.where(kind == 0, a + b * Y, a + b * np.exp(c * Y))

    def forward(X=1, P=1):
        X = np.asarray(X, dtype=np.float64)
        with np.errstate(all="ignore"):
            return np.where(kind == 0, (X - a) / b, np.log((X - a) / b) / c)

    def dXdY(X):
        X = np.asarray(X, dtype=np.float64)
        return np.where(kind == 0, b, c * (X - a))

    def det_jac(X=1):
        return float(np.prod(1.0 / dXdY(X))) if true_det else 1

    def inv_jac(X=1):
        return np.diag(dXdY(X))

    def grad_det_jac(X=1):
        # d/dX_k prod_i 1/(dX_i/dY_i): kind 0 contributes nothing, kind 1 has d/dX_k 1/(c (X_k - a)) = -1/(c (X_k - a)^2)
        if not true_det:
            return np.zeros(len(X))
        X = np.asarray(X, dtype=np.float64)
        return np.where(kind == 0, 0.0, -det_jac(X) / (X - a))

    model.parametrization_backward = backward
    model.parametrization_forward = forward
    model.parametrization_det_jac = det_jac
    model.parametrization_inv_jac = inv_jac
    model.grad_det_jac = grad_det_jac


def _design_of(g, prefix):
    return {k[len(prefix):]: (v.item() if getattr(v, "ndim", 1) == 0 else v) for k, v in g.items() if k.startswith(prefix)}


def model_members(g):
    """The models of a golden fixture's likelihood, one dict each (a single-model fixture has one)."""
    members = [dict(theta_nom=g["theta_nom"], unc_idx=g["unc_idx"], y=g["y"], std_y=g["std_y"], design=_design_of(g, "design_"))]
    for m in range(1, int(g["n_models"]) if "n_models" in g else 1):
        pre = "m%d_" % m
        members.append(dict(theta_nom=g[pre + "theta_nom"], unc_idx=g[pre + "unc_idx"], y=g[pre + "y"], std_y=g[pre + "std_y"],
                            design=_design_of(g, pre + "design_")))
    return members


def problem_dict(g):
    """The dict layout of pybitup_b200.synthetic builders, from a golden fixture or a builder dict."""
    if "design" in g:
        return g
    mem = model_members(g)
    out = dict(model=str(g["model"]), theta_nom=g["theta_nom"], unc_idx=g["unc_idx"], y=g["y"],
               std_y=g["std_y"], lb=g["lb"], ub=g["ub"], gamma=float(g["gamma"]), design=mem[0]["design"],
               x0=g["x0"], prop_cov=g["prop_cov"])
    if "param" in g:
        out["param"] = device_param(g["param"])
    if len(mem) > 1:          # several models: points concatenated, one block per model
        out["y"] = np.concatenate([np.atleast_2d(m["y"]) for m in mem], axis=1)
        out["std_y"] = np.concatenate([m["std_y"] for m in mem])
        out["blocks"] = [dict(theta_nom=m["theta_nom"], unc_idx=m["unc_idx"], design=m["design"],
                              n_points=np.atleast_2d(m["y"]).shape[1]) for m in mem]
    return out


def _oracle_member(model, m, lb, ub, gamma):
    design = dict(m["design"])
    if "n_steps" in design:
        design["n_steps"] = int(design["n_steps"])
    return orc.Problem(model=model, theta_nom=np.array(m["theta_nom"], dtype=np.float64),
                       unc_idx=np.array(m["unc_idx"], dtype=np.int64), y=np.array(np.atleast_2d(m["y"]), dtype=np.float64),
                       std_y=np.array(m["std_y"], dtype=np.float64), lb=np.array(lb, dtype=np.float64),
                       ub=np.array(ub, dtype=np.float64), gamma=float(gamma), design=design)


def oracle_problem(g):
    if "design" in g and "blocks" not in g:
        mem = [dict(theta_nom=g["theta_nom"], unc_idx=g["unc_idx"], y=g["y"], std_y=g["std_y"], design=g["design"])]
    elif "blocks" in g:
        mem, at = [], 0
        for b in g["blocks"]:
            n = int(b["n_points"])
            mem.append(dict(theta_nom=b["theta_nom"], unc_idx=b["unc_idx"], y=np.atleast_2d(g["y"])[:, at:at + n],
                            std_y=g["std_y"][at:at + n], design=b["design"]))
            at += n
    else:
        mem = model_members(g)
    probs = [_oracle_member(str(g["model"]), m, g["lb"], g["ub"], g["gamma"]) for m in mem]
    probs[0].extra = probs[1:]
    if g.get("param") is not None and "det_jac" in g["param"]:       # (a fixture's recorded hooks; builder dicts carry the engine's form)
        probs[0].param = g["param"]
    return probs[0]


def start_point(g):
    """Starting point of a recorded run in the SAMPLING space: parametrization_forward(initial_val)
    (solve_problem.py:309), which is the first chain row the reference saved."""
    return np.array(g["chain"][0], dtype=np.float64) if "param" in g else np.array(g["x0"], dtype=np.float64)


def wants_map(g):
    return g["algo"].get("estimate_max_distr") == "yes"


def mh_kwargs(algo):
    name = algo["name"]
    kw = dict(adaptive=name in ("AMH", "DRAM"), delayed=name in ("DR", "DRAM"))
    if name in ("AMH", "DRAM"):
        kw.update(updating_it=int(algo[name]["updating_it"]), eps_v=float(algo[name]["eps_v"]))
    if name in ("DR", "DRAM"):
        kw.update(gamma_dr=float(algo[name]["gamma"]))
    return kw


def golden_chain(g):
    """Reference chain rows 0..T (DRAM records the initial row twice: double __init__, Q8)."""
    chain = g["chain"]
    if g["algo"]["name"] == "DRAM":
        chain = chain[1:]
    return chain


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    den = np.maximum(np.abs(b), 1e-300)
    e = np.abs(a - b) / den
    e = np.where(both_inf, 0.0, e)
    e = np.where(np.isnan(a) & np.isnan(b), 0.0, e)
    return float(np.max(e)) if e.size else 0.0
