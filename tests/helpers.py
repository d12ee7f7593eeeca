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
    return g


def problem_dict(g):
    """The dict layout of pybitup_b200.synthetic builders, from a golden fixture or a builder dict."""
    if "design" in g:
        return g
    design = {k[len("design_"):]: (v.item() if getattr(v, "ndim", 1) == 0 else v)
              for k, v in g.items() if k.startswith("design_")}
    return dict(model=str(g["model"]), theta_nom=g["theta_nom"], unc_idx=g["unc_idx"], y=g["y"],
                std_y=g["std_y"], lb=g["lb"], ub=g["ub"], gamma=float(g["gamma"]), design=design,
                x0=g["x0"], prop_cov=g["prop_cov"])


def oracle_problem(g):
    p = problem_dict(g)
    design = dict(p["design"])
    if "n_steps" in design:
        design["n_steps"] = int(design["n_steps"])
    return orc.Problem(model=p["model"], theta_nom=np.array(p["theta_nom"], dtype=np.float64),
                       unc_idx=np.array(p["unc_idx"], dtype=np.int64), y=np.array(p["y"], dtype=np.float64),
                       std_y=np.array(p["std_y"], dtype=np.float64), lb=np.array(p["lb"], dtype=np.float64),
                       ub=np.array(p["ub"], dtype=np.float64), gamma=float(p["gamma"]), design=design)


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
