"""Host-side handle on the CUDA chain engine (thin wrapper over the C ABI).

``ProblemSpec`` is the flat problem descriptor the reference's objects are lowered to
(SURVEY.md §8b "what the engine must extract"); ``ChainEngine`` owns one ``pbu_engine``.
PyTorch is used only to allocate the device output buffers and to hand over the CUDA stream.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi as capi


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


@dataclass
class ProblemSpec:
    """One Bayesian inverse problem with a registered device forward model."""
    model: str                       # 'linear' | 'poly' | 'arrhenius' | 'heat'
    theta_nom: np.ndarray            # f64[P]   Model.param_nom
    unc_idx: np.ndarray              # int[d]   position of each uncertain parameter in theta
    y: np.ndarray                    # f64[n_runs, N]
    std_y: np.ndarray                # f64[N]
    lb: np.ndarray                   # f64[d]
    ub: np.ndarray                   # f64[d]
    gamma: float = 1.0
    design: dict = field(default_factory=dict)
    log_const: float = None          # analytic densities: constant that replaces the log-prior constant
    # several models in one likelihood (Likelihood.models, bayesian_inference.py:500-524): one dict per model with
    # theta_nom [P], unc_idx [d] (-1 = not a parameter of that model), design {...}, n_points; y / std_y hold the
    # models' points concatenated in this order, and theta_nom / unc_idx / design above are those of the first model
    blocks: list = None
    # element-wise change of variables Y (sampling space) -> X (model space), the lowered Model.parametrization_* hooks
    # (bayesian_inference.py:314-336): {"kind": int[d] (0: X = a + b Y, 1: X = a + b exp(c Y)), "a", "b", "c": f64[d],
    # "ldj_c0": float, "ldj_s": f64[d]} with -log(det_jac(X)) = ldj_c0 + ldj_s . Y; None = identity
    param: dict = None

    @classmethod
    def from_dict(cls, p):
        return cls(model=str(p["model"]), theta_nom=_f64(p["theta_nom"]),
                   unc_idx=np.ascontiguousarray(np.asarray(p["unc_idx"], dtype=np.int32)),
                   y=_f64(np.atleast_2d(p["y"])), std_y=_f64(p["std_y"]), lb=_f64(p["lb"]), ub=_f64(p["ub"]),
                   gamma=float(p.get("gamma", 1.0)), design=dict(p.get("design", {})), log_const=p.get("log_const"),
                   blocks=p.get("blocks"), param=p.get("param"))

    @property
    def d(self):
        return int(len(self.unc_idx))

    @property
    def n_points(self):
        return int(self.y.shape[1])

    @staticmethod
    def _design_and_consts(model, design_dict, N, P):
        """(design [N, ncols] or None, ncols, consts[8]) of one model's N points."""
        consts = [0.0] * 8
        design, ncols = None, 0
        if model == "linear":
            design = _f64(design_dict["Phi"])
            if design.shape != (N, P):
                raise ValueError("linear model: Phi must be [N, P] = [%d, %d], got %s" % (N, P, design.shape))
            ncols = design.shape[1]
        elif model == "soize":
            design = np.array([[0.0], [1.0]])
            ncols = 1
        elif model in ("poly", "heat"):
            design = _f64(design_dict["x"]).reshape(N, 1)
            ncols = 1
            if model == "heat":
                consts[0] = float(design_dict.get("L", 70.0))
        elif model == "arrhenius":
            dg = design_dict
            if int(dg["n_steps"]) != N:
                raise ValueError("arrhenius model: n_steps (%d) must equal the number of data points (%d)"
                                 % (int(dg["n_steps"]), N))
            consts[0], consts[1], consts[2] = float(dg["T0"]), float(dg["dT"]), float(dg["beta"])
        return design, ncols, consts

    def _c_struct(self):
        """Build the pbu_problem struct; returns (struct, keepalive list)."""
        if self.model not in capi.MODEL_IDS:
            raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED,
                                          "model %r is not a registered device model %s"
                                          % (self.model, sorted(capi.MODEL_IDS)))
        keep = []
        theta = _f64(self.theta_nom)
        unc = np.ascontiguousarray(np.asarray(self.unc_idx, dtype=np.int32))
        y = _f64(np.atleast_2d(self.y))
        std = _f64(self.std_y)
        lb, ub = _f64(self.lb), _f64(self.ub)
        N = y.shape[1]
        s = capi.pbu_problem()
        if not self.blocks:
            design, ncols, consts = self._design_and_consts(self.model, self.design, N, len(theta))
            s.n_models = 0
        else:
            # several models in one likelihood: points concatenated, per-model parameters / name map / constants
            P, d = len(theta), len(unc)
            first, count, th, ui, cs, designs = [], [], [], [], [], []
            at = 0
            for blk in self.blocks:
                n = int(blk["n_points"])
                dsg, ncols, c = self._design_and_consts(self.model, blk["design"], n, P)
                bt = _f64(blk["theta_nom"])
                bu = np.asarray(blk["unc_idx"], dtype=np.int32)
                if bt.shape != (P,) or bu.shape != (d,):
                    raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED,
                                                  "the models of one posterior must have the same number of parameters")
                first.append(at)
                count.append(n)
                th.append(bt)
                ui.append(bu)
                cs.append(c)
                designs.append(dsg)
                at += n
            if at != N:
                raise ValueError("the models' points add up to %d, the data has %d" % (at, N))
            design = None if designs[0] is None else np.ascontiguousarray(np.concatenate(designs, axis=0))
            consts = cs[0]
            mf = np.ascontiguousarray(np.array(first, dtype=np.int32))
            mc = np.ascontiguousarray(np.array(count, dtype=np.int32))
            mt = np.ascontiguousarray(np.stack(th))
            mu = np.ascontiguousarray(np.stack(ui).astype(np.int32))
            mk = np.ascontiguousarray(np.array(cs, dtype=np.float64))
            keep += [mf, mc, mt, mu, mk]
            s.n_models = len(self.blocks)
            s.model_first = mf.ctypes.data_as(C.POINTER(C.c_int32))
            s.model_count = mc.ctypes.data_as(C.POINTER(C.c_int32))
            s.model_theta_nom = capi.dptr(mt)
            s.model_unc_idx = mu.ctypes.data_as(C.POINTER(C.c_int32))
            s.model_consts = capi.dptr(mk)
        if std.shape != (N,) or lb.shape != unc.shape or ub.shape != unc.shape:
            raise ValueError("inconsistent problem array shapes")
        keep += [theta, unc, y, std, lb, ub, design]
        s.model = capi.MODEL_IDS[self.model]
        s.n_param, s.n_unc, s.n_points, s.n_runs, s.n_design_cols = len(theta), len(unc), N, y.shape[0], ncols
        s.theta_nom = capi.dptr(theta)
        s.unc_idx = unc.ctypes.data_as(C.POINTER(C.c_int32))
        s.design = capi.dptr(design) if design is not None else None
        s.y, s.std_y, s.lb, s.ub = capi.dptr(y), capi.dptr(std), capi.dptr(lb), capi.dptr(ub)
        s.gamma = float(self.gamma)
        s.consts = (C.c_double * 8)(*consts)
        s.has_log_const = 0 if self.log_const is None else 1
        s.log_const = 0.0 if self.log_const is None else float(self.log_const)
        if self.param is not None:
            d = len(unc)
            kind = np.ascontiguousarray(np.asarray(self.param["kind"], dtype=np.int32))
            pa, pb, pc, ps = (_f64(self.param[k]) for k in ("a", "b", "c", "ldj_s"))
            if not all(v.shape == (d,) for v in (kind, pa, pb, pc, ps)):
                raise ValueError("parametrisation arrays must have one entry per uncertain parameter")
            keep += [kind, pa, pb, pc, ps]
            s.par_kind = kind.ctypes.data_as(C.POINTER(C.c_int32))
            s.par_a, s.par_b, s.par_c, s.par_ldj_s = capi.dptr(pa), capi.dptr(pb), capi.dptr(pc), capi.dptr(ps)
            s.par_ldj_c0 = float(self.param["ldj_c0"])
        return s, keep

    # -- the change of variables on the host (start points, MAP / chain files) -------------------------------------
    def backward(self, Y):
        """X = parametrization_backward(Y) of the lowered change of variables; Y [..., d]."""
        Y = np.asarray(Y, dtype=np.float64)
        if self.param is None:
            return Y
        k, a, b, c = (np.asarray(self.param[n]) for n in ("kind", "a", "b", "c"))
        return np.where(k == 1, a + b * np.exp(c * Y), a + b * Y)


@dataclass
class SamplerConfig:
    """Algorithm block (inference_problem.py:37-107) plus engine-only knobs."""
    algo: str = "RWMH"               # RWMH | AMH | DR | DRAM | ISDE
    updating_it: int = 1
    eps_v: float = 0.0
    gamma_dr: float = 1.0
    isde_h: float = 0.0
    isde_f0: float = 0.0
    isde_gradient: str = "Analytical"
    adapt_start: int = 0
    adapt_update: int = 1
    adapt_end: int = 0
    am_welford: bool = False         # False = the reference's literal recursion (mha.py:455,:463)
    nan_rejects: bool = False        # False = reference semantics, min(1, nan) == 1 accepts
    lanes_per_chain: int = 0
    block_threads: int = 0
    chains_per_thread: int = 0
    tile_points: int = 0
    seed: int = 0
    chain_offset: int = 0
    hmc_step_size: float = 0.0
    hmc_num_steps: int = 0
    cooperative: int = 0             # 0 auto, 1 cooperative teams, -1 replicated layout only
    track_map: bool = False          # per-chain MAP estimate (opt_arg estimate_max_distr, mha.py:226-237)

    def _c_struct(self):
        s = capi.pbu_sampler_cfg()
        if self.algo not in capi.ALGO_IDS:
            raise ValueError('Algorithm "{}" unknown.'.format(self.algo))     # inference_problem.py:32-33
        s.algo = capi.ALGO_IDS[self.algo]
        s.updating_it, s.eps_v, s.gamma_dr = int(self.updating_it), float(self.eps_v), float(self.gamma_dr)
        s.isde_h, s.isde_f0 = float(self.isde_h), float(self.isde_f0)
        s.isde_gradient = 0 if self.isde_gradient == "Analytical" else 1
        s.adapt_start, s.adapt_update, s.adapt_end = int(self.adapt_start), int(self.adapt_update), int(self.adapt_end)
        s.am_welford, s.nan_rejects = int(self.am_welford), int(self.nan_rejects)
        s.lanes_per_chain, s.block_threads = int(self.lanes_per_chain), int(self.block_threads)
        s.chains_per_thread = int(self.chains_per_thread)
        s.tile_points = int(self.tile_points)
        s.seed, s.chain_offset = int(self.seed) & (2 ** 64 - 1), int(self.chain_offset)
        s.hmc_step_size, s.hmc_num_steps = float(self.hmc_step_size), int(self.hmc_num_steps)
        s.cooperative = int(self.cooperative)
        s.track_map = int(self.track_map)
        return s


@dataclass
class RunResult:
    """Device-resident outputs of one ``ChainEngine.run`` (torch tensors, chain-minor layout)."""
    n_steps: int
    thin: int
    samples: object = None      # [n_keep, d, Cs]  (Cs = sample_chains leading chains)
    aux: object = None          # [n_keep, d, Cs]  ISDE momentum
    lp_kept: object = None      # [n_keep, C]
    trace_lp1: object = None    # [n_steps, C]
    trace_lp2: object = None    # [n_steps, C]
    trace_acc: object = None    # [n_steps, C] uint8


class ChainEngine:
    """C chains of one problem on one GPU."""

    def __init__(self, problem, config, n_chains, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pybitup_b200 needs a CUDA device; there is no CPU fallback")
        self._torch = torch
        self.problem = problem if isinstance(problem, ProblemSpec) else ProblemSpec.from_dict(problem)
        self.config = config
        self.n_chains = int(n_chains)
        self.device = int(device)
        self.d = self.problem.d
        ps, keep = self.problem._c_struct()
        cs = config._c_struct()
        h = C.c_void_p()
        capi.check(capi.lib.pbu_engine_create(C.byref(ps), C.byref(cs), self.n_chains, self.device, C.byref(h)))
        self._h = h
        del keep
        self.use_torch_stream()

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            capi.lib.pbu_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def use_torch_stream(self):
        s = self._torch.cuda.current_stream(self.device)
        capi.check(capi.lib.pbu_engine_set_stream(self._h, C.c_void_p(s.cuda_stream)))

    # -- state --------------------------------------------------------------------------------
    def init_chains(self, x0, V0=None):
        x0 = _f64(x0)
        if x0.ndim == 1:
            x0 = np.ascontiguousarray(np.broadcast_to(x0, (self.n_chains, self.d)))
        if x0.shape != (self.n_chains, self.d):
            raise ValueError("x0 must be [n_chains, d] = [%d, %d], got %s" % (self.n_chains, self.d, x0.shape))
        V = None
        if V0 is not None:
            V = _f64(V0)
            if V.shape != (self.d, self.d):
                raise ValueError("V0 must be [d, d]")
        capi.check(capi.lib.pbu_engine_init_chains(self._h, capi.dptr(x0), capi.dptr(V) if V is not None else None))

    def set_streams(self, normals, uniforms):
        """Parity mode: per-chain pre-generated draws, normals [C, n], uniforms [C, m]."""
        if normals is None:
            capi.check(capi.lib.pbu_engine_set_streams(self._h, None, 0, None, 0))
            return
        nz = _f64(normals)
        if nz.ndim == 1:
            nz = nz[None, :]
        if nz.shape[0] != self.n_chains:
            raise ValueError("streams need one row per chain")
        nz = np.ascontiguousarray(nz)
        if uniforms is None:          # the Ito-SDE sampler draws normals only
            capi.check(capi.lib.pbu_engine_set_streams(self._h, capi.dptr(nz), nz.shape[1], None, 0))
            return
        un = _f64(uniforms)
        if un.ndim == 1:
            un = un[None, :]
        if un.shape[0] != self.n_chains:
            raise ValueError("streams need one row per chain")
        un = np.ascontiguousarray(un)
        capi.check(capi.lib.pbu_engine_set_streams(self._h, capi.dptr(nz), nz.shape[1], capi.dptr(un), un.shape[1]))

    @property
    def iteration(self):
        return int(capi.lib.pbu_engine_iteration(self._h))

    def launch_geometry(self):
        """(lanes per chain, threads per block, blocks, dynamic shared memory bytes) of the step kernel."""
        a, q, b, g, m = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        capi.check(capi.lib.pbu_engine_get_launch(self._h, C.byref(a), C.byref(q), C.byref(b), C.byref(g), C.byref(m)))
        mx, co, rp, un = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        capi.check(capi.lib.pbu_engine_get_variant(self._h, C.byref(mx), C.byref(co), C.byref(rp), C.byref(un)))
        return dict(lanes=a.value, chains_per_thread=q.value, block=b.value, grid=g.value, smem_bytes=m.value,
                    mixed=mx.value, cooperative=co.value, replicated_tables=rp.value, uniform_weight=un.value)

    # -- the hot loop -------------------------------------------------------------------------
    def run(self, n_steps, thin=1, keep_samples=True, keep_lp=False, trace=False, out=None, sample_chains=None,
            keep_aux=False):
        """Advance all chains by n_steps iterations in one kernel launch (asynchronous)."""
        torch = self._torch
        n_steps, thin = int(n_steps), int(thin)
        dev = torch.device("cuda", self.device)
        res = out if out is not None else RunResult(n_steps=n_steps, thin=thin)
        n_keep = int(capi.lib.pbu_engine_n_keep(self._h, n_steps, thin))
        Cn, d = self.n_chains, self.d
        Cs = Cn if sample_chains is None else max(1, min(int(sample_chains), Cn))
        if out is None:
            if keep_samples:
                res.samples = torch.empty((n_keep, d, Cs), dtype=torch.float64, device=dev)
            if keep_aux:
                res.aux = torch.empty((n_keep, d, Cs), dtype=torch.float64, device=dev)
            if keep_lp:
                res.lp_kept = torch.empty((n_keep, Cn), dtype=torch.float64, device=dev)
            if trace:
                res.trace_lp1 = torch.empty((n_steps, Cn), dtype=torch.float64, device=dev)
                res.trace_lp2 = torch.empty((n_steps, Cn), dtype=torch.float64, device=dev)
                res.trace_acc = torch.empty((n_steps, Cn), dtype=torch.uint8, device=dev)
        o = capi.pbu_run_outputs()
        o.n_sample_chains = res.samples.shape[2] if res.samples is not None else (res.aux.shape[2] if res.aux is not None else Cn)
        for name in ("samples", "lp_kept", "trace_lp1", "trace_lp2", "trace_acc", "aux"):
            t = getattr(res, name)
            setattr(o, name, C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None)
        capi.check(capi.lib.pbu_engine_run(self._h, n_steps, thin, C.byref(o)))
        return res

    def synchronize(self):
        capi.check(capi.lib.pbu_engine_synchronize(self._h))

    # -- read back ----------------------------------------------------------------------------
    def state(self):
        x = np.empty((self.n_chains, self.d))
        lp = np.empty(self.n_chains)
        capi.check(capi.lib.pbu_engine_get_state(self._h, capi.dptr(x), capi.dptr(lp)))
        return x, lp

    def map_estimate(self):
        """(arg_MAP [C, d], MAP_val [C] in log) of every chain (mha.py:138-139, :226-237); needs track_map."""
        x = np.empty((self.n_chains, self.d))
        lp = np.empty(self.n_chains)
        capi.check(capi.lib.pbu_engine_get_map(self._h, capi.dptr(x), capi.dptr(lp)))
        return x, lp

    def counters(self):
        n_rej = np.empty(self.n_chains, dtype=np.int64)
        mean_r = np.empty(self.n_chains)
        n_eval = np.empty(self.n_chains, dtype=np.int64)
        status = np.empty(self.n_chains, dtype=np.int32)
        capi.check(capi.lib.pbu_engine_get_counters(
            self._h, n_rej.ctypes.data_as(C.POINTER(C.c_int64)), capi.dptr(mean_r),
            n_eval.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(n_rejected=n_rej, mean_r=mean_r, n_eval=n_eval, status=status)

    def read_state(self):
        """State and counters of every chain through PINNED staging buffers (allocated once): one asynchronous copy per
        array, one synchronisation.  Returns numpy views of the staging buffers (valid until the next call): x [C, d] (a
        transposed view of the device layout [d, C]), lp, n_rejected, mean_r, n_eval, status."""
        torch = self._torch
        if not hasattr(self, "_stage"):
            Cn, d = self.n_chains, self.d
            mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()      # noqa: E731
            self._stage = dict(x=mk((d, Cn), torch.float64), lp=mk((Cn,), torch.float64), n_rejected=mk((Cn,), torch.int64),
                               mean_r=mk((Cn,), torch.float64), n_eval=mk((Cn,), torch.int64), status=mk((Cn,), torch.int32))
        st = self._stage
        capi.check(capi.lib.pbu_engine_read_state(self._h, *[C.c_void_p(st[k].data_ptr()) for k in
                                                             ("x", "lp", "n_rejected", "mean_r", "n_eval", "status")]))
        out = {k: v.numpy() for k, v in st.items()}
        out["x"] = out["x"].T
        return out

    def proposal(self):
        R = np.empty((self.n_chains, self.d, self.d))
        V = np.empty((self.n_chains, self.d, self.d))
        xav = np.empty((self.n_chains, self.d))
        capi.check(capi.lib.pbu_engine_get_proposal(self._h, capi.dptr(R), capi.dptr(V), capi.dptr(xav)))
        return dict(R=R, V=V, X_av=xav)

    def chain_sums(self, mu=None):
        """Additive sums of the kept-state moments over this engine's chains (K4), about the point mu."""
        n = int(capi.lib.pbu_engine_chain_sums_len(self._h))
        out = np.empty(n)
        m = None if mu is None else _f64(mu)
        capi.check(capi.lib.pbu_engine_chain_sums(self._h, capi.dptr(m) if m is not None else None, capi.dptr(out)))
        return out

    def reset_moments(self):
        capi.check(capi.lib.pbu_engine_reset_moments(self._h))

    def set_proposal(self, V):
        """Pooled-covariance mode: give every chain the proposal covariance / preconditioner V [d, d]."""
        V = _f64(V)
        if V.shape != (self.d, self.d):
            raise ValueError("V must be [d, d]")
        capi.check(capi.lib.pbu_engine_set_proposal(self._h, capi.dptr(V)))

    def pooled_statistics(self, reducer=None, apply=False, scale=None, eps=None):
        """Pooled mean / covariance / Gelman-Rubin R-hat of the kept states of ALL chains (of all ranks when `reducer` is
        an active NCCL ``distributed.Reducer``), computed on the device: two passes over the per-chain running moments,
        each followed by one in-place all-reduce (``2 + d``, then ``2 + 4d + d(d+1)/2`` doubles), and a finalisation kernel.  Nothing is
        allocated per call after the first, nothing is copied to the host and the stream is not synchronised.

        apply=True also turns the pooled covariance into every chain's proposal (MH family: ``R = chol(scale (cov + eps
        I))``, mha.py:463, :486-487) or preconditioner (ISDE / HMC: ``C_approx``, ``inv_L_c``, mha.py:745-746, :773-777)
        -- the pooled-covariance mode that replaces one chain's recursion by the covariance of all chains.
        Returns the device statistics vector (torch tensor, layout in include/pybitup_b200.h); ``statistics_dict`` turns
        it into the host dictionary."""
        torch = self._torch
        if not hasattr(self, "_mom"):
            dev = torch.device("cuda", self.device)
            L = int(capi.lib.pbu_engine_chain_sums_len(self._h))
            n = int(capi.lib.pbu_engine_pooled_stats_len(self._h))
            self._mom = dict(s1=torch.zeros(2 + self.d, dtype=torch.float64, device=dev), s2=torch.zeros(L, dtype=torch.float64, device=dev),
                             mu=torch.zeros(self.d, dtype=torch.float64, device=dev),
                             stats=torch.zeros(n, dtype=torch.float64, device=dev))
        m = self._mom
        dsum = reducer.device_sum if reducer is not None else None
        if reducer is not None and reducer.active and reducer.world_size > 1 and dsum is None:
            raise RuntimeError("pooled_statistics needs an NCCL process group (the all-reduce runs on the device)")
        if dsum is not None and reducer.device.index != self.device:
            raise RuntimeError("the reducer is bound to cuda:%s, the engine to cuda:%d" % (reducer.device.index, self.device))
        vp = lambda t: C.c_void_p(t.data_ptr())          # noqa: E731
        capi.check(capi.lib.pbu_engine_chain_sums_dev(self._h, None, vp(m["s1"]), 2 + self.d))     # pass 1: count and first moments
        if dsum is not None:
            dsum(m["s1"])
        capi.check(capi.lib.pbu_engine_pooled_mean_dev(self._h, vp(m["s1"]), None, vp(m["mu"])))
        capi.check(capi.lib.pbu_engine_chain_sums_dev(self._h, vp(m["mu"]), vp(m["s2"]), 0))
        if dsum is not None:
            dsum(m["s2"])
        isde = self.config.algo in ("ISDE", "HMC")
        if scale is None:
            scale = 1.0 if isde else 2.38 ** 2 / self.d              # mha.py:630, :402
        if eps is None:
            eps = 1e-10 if isde else float(self.config.eps_v)        # mha.py:631, :403
        capi.check(capi.lib.pbu_engine_pooled_finalize_dev(self._h, vp(m["s2"]), vp(m["mu"]), float(scale), float(eps),
                                                           1 if apply else 0, vp(m["stats"])))
        return m["stats"]

    def statistics_dict(self, stats):
        """Host dictionary of a statistics vector returned by ``pooled_statistics`` (one small device -> host copy)."""
        v = stats.cpu().numpy()
        d, nt = self.d, self.d * (self.d + 1) // 2
        o = 4
        out = dict(n_chains=int(round(v[0])), n=int(round(v[1])), proposal_ok=bool(v[2] != 0.0))
        for name, n, shape in (("mean", d, (d,)), ("cov", d * d, (d, d)), ("W", d, (d,)), ("B_over_n", d, (d,)), ("rhat", d, (d,))):
            out[name] = v[o:o + n].reshape(shape).copy()
            o += n
        iu = np.triu_indices(d)
        for name in ("V", "R", "inv_R"):
            M = np.zeros((d, d))
            M[iu] = v[o:o + nt]
            out[name] = M + np.triu(M, 1).T if name == "V" else M
            o += nt
        return out

    def statistics(self, reducer=None):
        """Pooled mean / covariance and R-hat of the kept states of all chains (of all ranks with an NCCL reducer)."""
        return self.statistics_dict(self.pooled_statistics(reducer))

    def isde_state(self):
        P = np.empty((self.n_chains, self.d))
        Cm = np.empty((self.n_chains, self.d, self.d))
        capi.check(capi.lib.pbu_engine_get_isde_state(self._h, capi.dptr(P), capi.dptr(Cm)))
        return dict(P=P, C=Cm)

    # -- evaluation only ----------------------------------------------------------------------
    def _check_X(self, X):
        X = _f64(np.atleast_2d(X))
        if X.ndim != 2 or X.shape[1] != self.d:
            raise ValueError("X must be [S, d] = [S, %d], got %s" % (self.d, X.shape))
        return X

    def log_posterior(self, X):
        X = self._check_X(X)
        out = np.empty(X.shape[0])
        capi.check(capi.lib.pbu_eval_logpost(self._h, capi.dptr(X), X.shape[0], capi.dptr(out)))
        return out

    def model_eval(self, X):
        X = self._check_X(X)
        out = np.empty((X.shape[0], self.problem.n_points))
        capi.check(capi.lib.pbu_eval_model(self._h, capi.dptr(X), X.shape[0], capi.dptr(out)))
        return out
