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
    return probs[0]


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
