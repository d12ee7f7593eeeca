"""Sampler classes: host-side mirror of ``pybitup/metropolis_hastings_algorithms.py``.

Same class names, constructor signatures and ``run_algorithm()`` entry point as the reference
(``MetropolisHastings`` :26-368, ``AdaptiveMetropolisHastings`` :370-487, ``DelayedRejectionMetropolisHastings``
:490-567, ``DelayedRejectionAdaptiveMetropolisHastings`` :570-595, ``ito_SDE`` :886-1040), but the per-iteration
Python loop is one fused CUDA kernel over ``n_chains`` independent chains (``ChainEngine``).  Engine-only options
ride in ``opt_arg`` (the ``Sampler`` copies them from the ``Algorithm`` JSON block): ``n_chains`` (default 1),
``seed``, ``init_jitter``, ``device``, ``launch_iterations``, ``save_all_chains``, ``am_welford``,
``pooled_covariance`` / ``pooled_update_it``.

One process per GPU: when ``torch.distributed`` is initialised (``torchrun``), ``n_chains`` is the GLOBAL number of chains;
every rank owns a contiguous range of the global chain index space (``distributed.shard_range``; the reference splits
its propagation samples over MPI ranks the same way, ``solve_problem.py:497-529``), runs on its ``LOCAL_RANK`` GPU with
``chain_offset`` = the first global index (which keys the Philox streams, so a chain does not depend on how the chains
are sharded), and rank 0 -- the owner of global chain 0 -- writes the reference's files.  The data path has no
collective; the only exchange is the small moment all-reduce of ``ChainEngine.pooled_statistics``: once at the end
(pooled mean / covariance, Gelman-Rubin R-hat) and, with ``"pooled_covariance": "yes"``, every ``updating_it`` /
``update_it`` iterations, when the covariance of ALL chains replaces each chain's own recursion in the proposal
(``adapt_covariance``, mha.py:432-487, :736-777).

Side effects reproduced: ``output/mcmc_chain.csv`` and ``mcmc_chain_reparam.csv`` (T+1 rows of
chain 0, ``%f``), the ``output.txt`` header parsed by ``post_process`` (lines 1 and 3) and footer,
``output/model_eval/<id>_fun_eval-<it>.npy`` every ``save_eval_freq`` iterations, ``./cov.csv``, the in-place update
of the caller's ``param_init`` to the final state.  Random numbers come from Philox on the device (the reference
uses the unseeded stdlib generator), so chains agree with the reference statistically; arithmetic parity with
injected streams is what ``tests/test_gpu_parity.py`` checks through the C ABI.
"""
import os
import time

import numpy as np

from .engine import ChainEngine, SamplerConfig


def time_it(f):
    def wrapped(*args, **kw):
        t0 = time.perf_counter()
        result = f(*args, **kw)
        print("{}: {:.3f} s".format(f.__name__, time.perf_counter() - t0))
        return result
    return wrapped


def compute_hessian_fd(fun_batch, x, eps=0.01, compute_all_element=True, positive_diag=True):
    """Finite-difference Hessian with the reference's stencil (computeHessianFD, mha.py:1115-1212), all stencil points
    evaluated in ONE batched device call: fun_batch(X[S, d]) -> f[S].

    Diagonal: central second difference with step x_i*eps, forced to -|.| when positive_diag (:1162-1164).
    Off-diagonal (i < j): (f_ij+ - f_i+ - f_j+ + 2 f_0 - f_i- - f_j- + f_ij-) / (2 d_i d_j).  In the reference the
    doubly perturbed points are built IN PLACE on the i-perturbed vectors (:1186-1189 alias them), so for a given i the
    perturbations of successive j accumulate; that is reproduced here because it decides the preconditioner."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    delta = x * eps
    pts = [x.copy()]
    for i in range(n):
        for sgn in (1.0, -1.0):
            q = x.copy()
            q[i] = x[i] + sgn * delta[i]
            pts.append(q)
    pair_slots = {}
    if compute_all_element:
        for i in range(n):
            acc_p, acc_m = x.copy(), x.copy()
            acc_p[i] = x[i] + delta[i]
            acc_m[i] = x[i] - delta[i]
            for j in range(i + 1, n):
                acc_p[j] = acc_p[j] + delta[j]          # cumulative over j: the reference's aliasing
                acc_m[j] = acc_m[j] - delta[j]
                pair_slots[(i, j)] = len(pts)
                pts.append(acc_p.copy())
                pts.append(acc_m.copy())
    f = np.asarray(fun_batch(np.array(pts)), dtype=np.float64)
    f0 = f[0]
    H = np.zeros((n, n))
    for i in range(n):
        num = f[1 + 2 * i] - 2 * f0 + f[2 + 2 * i]
        H[i, i] = -np.abs(num / delta[i] ** 2) if positive_diag else num / delta[i] ** 2
    for (i, j), k in pair_slots.items():
        H[i, j] = (f[k] - f[1 + 2 * i] - f[1 + 2 * j] + 2 * f0 - f[2 + 2 * i] - f[2 + 2 * j] + f[k + 1]) / (2 * delta[j] * delta[i])
        H[j, i] = H[i, j]
    return H


def nearPD(A, nit=10):
    """The reference's nearPD (mha.py:1253-1284, helpers :1214-1250), restated with plain arrays.

    As written upstream it is Higham's alternating projections run on A ITSELF (the correlation matrix computed at
    :1266 is overwritten by ``corr = A`` at :1267) with the identity weight: project on the positive semi-definite
    cone by clipping eigenvalues (_getAplus; eigenvectors from the general ``eig``), then put ones on the diagonal
    (_getPu with W = I), ``nit`` times; the result is finally scaled by sigma_i sigma_j, sigma = sqrt|diag A|."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    sigma = np.sqrt(np.abs(np.diag(A)))
    W = np.identity(n)
    delta_s = 0
    Yk = A.copy()
    for _ in range(nit):
        Rk = Yk - delta_s
        eigval, Q = np.linalg.eig(Rk)                          # _getPs with W = I is _getAplus
        Xk = (Q @ np.diag(np.maximum(eigval, 0)) @ Q.T)
        delta_s = Xk - Rk
        Yk = np.array(Xk, copy=True)
        Yk[W > 0] = W[W > 0]                                   # _getPu
    return Yk * np.outer(sigma, sigma)                         # computeCovMat


class _DeviceSampler:
    """Common driver: build the engine, run the iterations in a few launches, write the reference's files."""

    algo_name = "RWMH"

    def _setup(self, IO_util, nIterations, param_init, V, prob_distr, opt_arg):
        self.nIterations = int(nIterations)
        self.opt_arg = dict(opt_arg or {})
        self.save_eval_freq = int(self.opt_arg.get("save_eval_freq", 10))
        self.estimate_max_distr = bool(self.opt_arg.get("estimate_max_distr", False))
        self.bool_estimate_sigma = bool(self.opt_arg.get("estimate_sigma", False))
        self.V = None if V is None else np.array(V, dtype=np.float64)
        self.V_0 = self.V
        self.prob_distr = prob_distr
        self.param_init = param_init                      # aliased and updated in place, as the reference does (:95-99)
        self.n_param = int(np.size(param_init))
        self.current_val = self.param_init
        self.IO_util = IO_util
        self.IO_fileID = IO_util["fileID"]
        from .distributed import Reducer, shard_range
        self.reducer = Reducer()                          # active under torchrun: one process per GPU
        self.rank, self.world = self.reducer.rank, self.reducer.world_size
        self.n_chains_total = int(self.opt_arg.get("n_chains", 1))
        if self.n_chains_total < self.world:
            raise ValueError("n_chains ({}) must be at least the number of ranks ({})".format(self.n_chains_total, self.world))
        self.chain_first, self.n_chains = shard_range(self.n_chains_total, self.world, self.rank)
        self.writer = self.rank == 0                      # owner of global chain 0: writes the reference's files
        self.pooled = str(self.opt_arg.get("pooled_covariance", "no")) == "yes"
        self.n_rejected = 0
        self.mean_r = 0.0
        self.it = 0
        self.problem = prob_distr.lower()                 # raises UnsupportedProblem: no CPU fallback
        self.engine = None

    # -- configuration of the engine (overridden per algorithm) -----------------------------------------
    def _config(self):
        return SamplerConfig(algo=self.algo_name)

    def _proposal_matrix(self, eng):
        return self.V

    def _pooled_interval(self):
        """Iterations between two pooled-covariance refreshes (pooled mode)."""
        return int(self.opt_arg.get("pooled_update_it", 0))

    def _pooled_due(self, it):
        k = self._pooled_interval()
        return self.pooled and k > 0 and it % k == 0

    def _start_points(self):
        """Starting points of THIS rank's chains.  They are drawn for the global chain index space and sliced, so that a
        chain starts at the same point however the chains are sharded."""
        C = self.n_chains_total
        x0 = np.tile(np.asarray(self.param_init, dtype=np.float64), (C, 1))
        jitter = float(self.opt_arg.get("init_jitter", 0.0))
        if C > 1 and jitter > 0.0:
            rng = np.random.default_rng(int(self.opt_arg.get("seed", 0)) + 7919)
            scale = np.where(x0[0] == 0.0, 1.0, np.abs(x0[0]))
            x0[1:] += jitter * scale * rng.standard_normal((C - 1, self.n_param))
            if getattr(self.problem, "param", None) is None:          # (the box lives in the model space X)
                lb, ub = self.problem.lb, self.problem.ub
                x0 = np.minimum(np.maximum(x0, lb), ub)
        return np.ascontiguousarray(x0[self.chain_first:self.chain_first + self.n_chains])

    def _write_header(self):
        f = self.IO_fileID['out_data']                    # mha.py:131-135
        f.write("$RandomVarName$\n")
        for name in self.prob_distr.name_random_var:
            f.write(name + ' ')
        f.write("\n$IterationNumber$\n{}\n".format(self.nIterations))

    def _write_rows(self, rows):
        """rows [n, d] of chain 0 -> both chain files, %f, comma (bayesian_inference.py:131-138)."""
        if len(rows) == 0:
            return
        np.savetxt(self.IO_fileID['MChains_reparam'], rows, fmt="%f", delimiter=",")
        model = getattr(self.prob_distr, "model", None)       # analytic densities have no model / parametrisation
        back = rows if model is None else np.array([model.parametrization_backward(r) for r in rows])
        np.savetxt(self.IO_fileID['MChains'], back, fmt="%f", delimiter=",")

    def _launch_len(self, done=0):
        n = int(self.opt_arg.get("launch_iterations", 0))
        if n <= 0:
            n = max(1, min(self.nIterations, 2000))
        k = self._pooled_interval() if self.pooled else 0
        if k > 0:                                         # a launch ends where the pooled covariance is refreshed
            n = min(n, k - done % k)
        return n

    @time_it
    def run_algorithm(self):
        from .distributed import local_device
        cfg = self._config()
        cfg.seed = int(self.opt_arg.get("seed", 0))
        cfg.chain_offset = int(self.opt_arg.get("chain_offset", 0)) + self.chain_first
        cfg.track_map = self.estimate_max_distr         # estimate_max on the device (mha.py:226-237, :1011-1019)
        device = int(self.opt_arg["device"]) if self.opt_arg.get("device") is not None else local_device()
        if self.reducer.device_sum is not None and self.reducer.device.index != device:
            raise RuntimeError("rank {} runs its chains on cuda:{} but its process group is bound to cuda:{}".format(
                self.rank, device, self.reducer.device.index))
        self.engine = eng = ChainEngine(self.problem, cfg, self.n_chains, device=device)
        self.t1 = time.perf_counter()
        x0 = self._start_points()
        eng.init_chains(x0, self._proposal_matrix(eng))
        save_all = str(self.opt_arg.get("save_all_chains", "no")) == "yes"
        isde = self.algo_name == "ISDE"                 # aux_variables.csv rows (the HMC file is opened but left empty)
        if self.writer:
            for key in ('MChains', 'MChains_reparam'):
                self.IO_fileID[key].seek(0)
            self._write_header()
            if self.algo_name == "DRAM":                  # double __init__ writes the header twice
                self._write_header()
            self._write_rows(x0[:1])
        kept0, kept_all, aux_rows = [x0[0].copy()], [], []
        done = 0
        while done < self.nIterations:
            n = min(self._launch_len(done), self.nIterations - done)
            # rank 0 records global chain 0 (the reference's chain files); the other ranks record nothing unless asked
            keep = self.writer or save_all
            res = eng.run(n, thin=1, keep_samples=keep, sample_chains=None if save_all else 1, keep_aux=isde and self.writer)
            if self._pooled_due(done + n):                # covariance of ALL chains -> every chain's proposal (device only)
                eng.pooled_statistics(self.reducer, apply=True)
            if keep:
                eng.synchronize()
                s = res.samples.cpu().numpy()             # [n, d, Cs]
                if save_all:
                    kept_all.append(s)
            if self.writer:
                rows = np.ascontiguousarray(s[:, :, 0])
                self._write_rows(rows)
                if isde:
                    aux_rows.append(np.ascontiguousarray(res.aux.cpu().numpy()[:, :, 0]))
                kept0.append(rows)
                # model evaluations of the current (last accepted) state every save_eval_freq iterations (:366-368)
                its = np.arange(done + 1, done + n + 1)
                sel = its[its % self.save_eval_freq == 0]
                if sel.size and self.save_eval_freq <= self.nIterations and hasattr(self.prob_distr, "likelihood"):
                    ev = eng.model_eval(rows[sel - done - 1])
                    for k, it in enumerate(sel):
                        self.prob_distr.save_value(self.IO_util, int(it), ev[k])
            done += n
            self.it = done
        self.chain = np.vstack([kept0[0][None, :]] + kept0[1:]) if len(kept0) > 1 else kept0[0][None, :]
        if save_all and kept_all:
            self.chains_all = np.concatenate(kept_all, axis=0)
            name = "mcmc_chains_all.npy" if self.world == 1 else "mcmc_chains_all_rank{}.npy".format(self.rank)
            np.save(str(self.IO_util['path']['out_folder'] / name), self.chains_all)
        if isde and self.writer:
            self._write_aux(x0[0], aux_rows)
        self._finish()

    def _write_aux(self, x0, aux_rows):
        """aux_variables.csv: T + 1 blocks of d rows [P, xi] -- the initial (0, xi_0) written by ito_SDE.__init__ and then
        (P_n, xi_n) AFTER the update of every iteration (mha.py:1032 runs after the assignment at :1003-1004)."""
        if 'aux_variables' not in self.IO_fileID or not aux_rows:
            return
        P = np.vstack([np.zeros((1, self.n_param))] + aux_rows)
        for it in range(self.nIterations + 1):
            np.savetxt(self.IO_fileID['aux_variables'], np.transpose(np.array([P[it], self.chain[it]])), fmt="%f", delimiter=",")

    def _finish(self):
        eng = self.engine
        # many-chain statistics (new): pooled mean / covariance, Gelman-Rubin R-hat over the chains of ALL ranks -- one
        # small all-reduce on the device; every rank takes part
        self.statistics = eng.statistics(self.reducer) if self.n_chains_total > 1 else None
        x, lp = eng.state()
        ctr = eng.counters()
        # in-place: the caller's array ends at the final state of global chain 0 (:215), on every rank
        self.param_init[...] = self.reducer.broadcast0(x[0])
        self.current_val = self.param_init
        self.distr_fun_current_val = float(lp[0])
        self.n_rejected = int(ctr["n_rejected"][0])
        self.mean_r = float(ctr["mean_r"][0])
        self.status = ctr["status"]
        n_bad = int(self.reducer(np.array([np.sum((ctr["status"] & 1) > 0)], dtype=np.float64), "sum")[0])
        if n_bad:
            raise np.linalg.LinAlgError("adapted covariance not positive definite in {} chain(s) "
                                        "(scipy.linalg.cholesky raises here in the reference, mha.py:487)".format(n_bad))
        prop = eng.proposal()
        self.R = prop["R"][0]
        self.V_i = prop["V"][0]
        self.X_av_i = prop["X_av"][0]
        # compute_covariance (:295-307): mean / covariance of chain 0 (from the exact states, not the 6-decimal CSV)
        self.mean_c = np.mean(self.chain, axis=0)
        self.cov_c = np.atleast_2d(np.cov(self.chain, rowvar=False)) if self.chain.shape[0] > 1 else np.zeros((self.n_param, self.n_param))
        rejection_rate = 0 if self.nIterations == 0 else self.n_rejected / self.nIterations * 100
        if self.estimate_max_distr:
            # MAP estimate (mha.py:138-139, :226-237): per chain on the device; the legacy attributes / files are global
            # chain 0's (what a single-chain reference run reports), all chains in arg_MAP_all / MAP_val_all
            self.arg_MAP_all, self.MAP_val_all = eng.map_estimate()
            self.arg_MAP = self.reducer.broadcast0(self.arg_MAP_all[0])
            self.MAP_val = float(self.reducer.broadcast0(self.MAP_val_all[:1])[0])
        if self.writer:
            if self.bool_estimate_sigma and 'estimated_sigma' in self.IO_fileID and hasattr(self.prob_distr, "likelihood"):
                import pandas as pd                           # terminate_loop :332-339: the unchanged std_y
                for model_id in self.prob_distr.likelihood.data.keys():
                    est_sigma = self.prob_distr.likelihood.data[model_id].std_y
                    pd.DataFrame({'model_id': est_sigma}, columns=['model_id']).to_csv(self.IO_fileID['estimated_sigma'])
            if self.estimate_max_distr and 'estimate_arg_max_val_distr' in self.IO_fileID:          # :352-355
                model = getattr(self.prob_distr, "model", None)
                arg = self.arg_MAP if model is None else model.parametrization_backward(self.arg_MAP)
                np.savetxt(self.IO_fileID['estimate_arg_max_val_distr'], np.array([arg]), fmt="%f", delimiter=",")
                np.savetxt(self.IO_fileID['estimate_max_val_distr'], np.array([np.exp(self.MAP_val)]), fmt="%f", delimiter=",")
            np.savetxt("cov.csv", self.cov_c, delimiter=",")
            print("\nRejection rate is {} %".format(rejection_rate))
            f = self.IO_fileID['out_data']                    # terminate_loop (:322-357)
            f.write("\nRejection rate is {} % \n".format(rejection_rate))
            f.write("\nElapsed time: {} \n".format(time.strftime("%H:%M:%S", time.gmtime(time.perf_counter() - self.t1))))
            f.write("\nMaximum Likelihood Estimator (MLE) \n")
            f.write("Log-likelihood value \n")
            f.write("\nCovariance Matrix is \n{}".format(self.cov_c))
            if self.statistics is not None:
                f.write("\nChains: {}\nPooled mean \n{}\nPooled covariance \n{}\nGelman-Rubin R-hat \n{}".format(
                    self.n_chains_total, self.statistics["mean"], self.statistics["cov"], self.statistics["rhat"]))
            for key in ('MChains', 'MChains_reparam', 'out_data'):
                self.IO_fileID[key].flush()
        eng.close()


class MetropolisHastings(_DeviceSampler):
    """Random-walk Metropolis-Hastings (mha.py:26-368)."""
    algo_name = "RWMH"

    def __init__(self, IO_util, nIterations, param_init, V, prob_distr, opt_arg):
        self._setup(IO_util, nIterations, param_init, V, prob_distr, opt_arg)


class AdaptiveMetropolisHastings(_DeviceSampler):
    """Adaptive Metropolis (mha.py:370-487); starting_it is parsed and unused, as in the reference."""
    algo_name = "AMH"

    def __init__(self, IO_util, nIterations, param_init, V, prob_distr, starting_it, updating_it, eps_v, opt_arg):
        self._setup(IO_util, nIterations, param_init, V, prob_distr, opt_arg)
        self.starting_it, self.updating_it, self.eps_v = int(starting_it), int(updating_it), float(eps_v)

    def _config(self):
        if self.pooled:       # the covariance of all chains replaces the per-chain recursion: the RWMH kernel, refreshed factor
            return SamplerConfig(algo="RWMH", eps_v=self.eps_v)
        return SamplerConfig(algo=self.algo_name, updating_it=self.updating_it, eps_v=self.eps_v,
                             am_welford=str(self.opt_arg.get("am_welford", "no")) == "yes")

    def _pooled_interval(self):
        return int(self.opt_arg.get("pooled_update_it", self.updating_it))


class DelayedRejectionMetropolisHastings(_DeviceSampler):
    """Delayed rejection (mha.py:490-567)."""
    algo_name = "DR"

    def __init__(self, IO_util, nIterations, param_init, V, prob_distr, gamma, opt_arg):
        self._setup(IO_util, nIterations, param_init, V, prob_distr, opt_arg)
        self.gamma = float(gamma)

    def _config(self):
        return SamplerConfig(algo=self.algo_name, gamma_dr=self.gamma)


class DelayedRejectionAdaptiveMetropolisHastings(_DeviceSampler):
    """DRAM (mha.py:570-595)."""
    algo_name = "DRAM"

    def __init__(self, IO_util, nIterations, param_init, V, prob_distr, starting_it, updating_it, eps_v, gamma, opt_arg):
        self._setup(IO_util, nIterations, param_init, V, prob_distr, opt_arg)
        self.starting_it, self.updating_it, self.eps_v, self.gamma = int(starting_it), int(updating_it), float(eps_v), float(gamma)

    def _config(self):
        if self.pooled:
            return SamplerConfig(algo="DR", eps_v=self.eps_v, gamma_dr=self.gamma)
        return SamplerConfig(algo=self.algo_name, updating_it=self.updating_it, eps_v=self.eps_v, gamma_dr=self.gamma,
                             am_welford=str(self.opt_arg.get("am_welford", "no")) == "yes")

    def _pooled_interval(self):
        return int(self.opt_arg.get("pooled_update_it", self.updating_it))


class _GradientSampler(_DeviceSampler):
    """Shared by ito_SDE and HamiltonianMonteCarlo: the GradientBasedMCMC preconditioner options (mha.py:621-731)."""

    def _setup_gradient(self, C_matrix, gradient):
        self.C_matrix, self.gradient = dict(C_matrix), gradient
        if gradient not in ("Numerical", "Analytical"):
            raise ValueError('Unknown gradient type "{}" for estimating gradients.'.format(gradient))     # :730-731
        value = self.C_matrix.get("value", "Identity")
        if value == "Identity":                                              # :661-664
            self.V = None
        elif value == "Matrix":                                              # :689-693
            self.V = np.array(self.C_matrix["matrix_value"], dtype=np.float64)
        elif value == "from_file":                                           # :694-697 (the generalised eigenproblem it
            path = os.path.join(str(self.IO_util['path']['cwd']), "cov.csv")      # then prints, :699-706, is diagnostics only)
            self.V = np.atleast_2d(np.loadtxt(path, delimiter=",", dtype=np.float64))
        elif value == "PSD":                                                 # :674-678: estimate_hessian_model hard-codes
            from ._capi import UnsupportedProblem, PBU_ERR_UNSUPPORTED        # a 14-species chemistry model
            raise UnsupportedProblem(PBU_ERR_UNSUPPORTED, 'preconditioner "PSD" is not available on the device path '
                                     '(Identity, Matrix, from_file, Hessian, Hessian_diag, nearPD are)')
        elif value not in ("Hessian", "Hessian_diag", "nearPD"):             # :653-660, :679-684, computed in _proposal_matrix
            raise ValueError('Unknown matrix type "{}" for estimating the conditioning matrix G .'.format(value))     # :708
        par = getattr(getattr(self, "problem", None), "param", None)
        if gradient == "Analytical" and par is not None:
            # the analytical gradient goes through the user's inv_jac / grad_det_jac hooks (bayesian_inference.py:604-636);
            # the device applies diag(dX/dY) and the gradient of -log(det_jac), so the hooks must be exactly that
            from ._capi import UnsupportedProblem, PBU_ERR_UNSUPPORTED
            if not par.get("grad_ok", False):
                raise UnsupportedProblem(PBU_ERR_UNSUPPORTED, "analytical gradient of a re-parametrised model: " + par.get("grad_why", ""))
            if self.problem.blocks:
                raise UnsupportedProblem(PBU_ERR_UNSUPPORTED, "analytical gradient of a re-parametrised posterior with several "
                                                              "models (each model's own inv_jac enters upstream) is not lowered")
        self.starting_it = int(self.C_matrix['starting_it'])
        self.update_it = int(self.C_matrix['update_it'])
        self.end_it = int(self.C_matrix['end_it']) if self.C_matrix.get("end_it") is not None else self.nIterations + 1

    def _pooled_interval(self):
        return int(self.opt_arg.get("pooled_update_it", self.update_it))

    def _pooled_due(self, it):
        # the adaptation window of GradientBasedMCMC.adapt_covariance (:740, :767): refreshes at multiples of update_it
        # inside [starting_it, end_it + 1]
        return _DeviceSampler._pooled_due(self, it) and self.starting_it <= it <= self.end_it + 1

    def _adapt_window(self):
        """(adapt_start, adapt_update, adapt_end) of the engine: the per-chain window, switched off in pooled mode."""
        return (0, self.update_it, 0) if self.pooled else (self.starting_it, self.update_it, self.end_it)

    def _proposal_matrix(self, eng):
        """C_approx: identity, a given matrix, or the inverse of minus the finite-difference Hessian of the
        log-posterior at the initial point (mha.py:653-660), its stencil evaluated in one device batch."""
        value = self.C_matrix.get("value", "Identity")
        if value in ("Hessian", "Hessian_diag", "nearPD"):
            self.hess_mat = -compute_hessian_fd(eng.log_posterior, np.array(self.param_init, dtype=np.float64), eps=0.001,
                                                compute_all_element=(value != "Hessian_diag"))
            self.V = np.linalg.inv(self.hess_mat)
            if value == "nearPD":                                            # :679-684
                self.V = nearPD(self.V)
        if self.V is not None:
            # :713-722: a preconditioner scipy cannot factorise is replaced by the identity (the FD Hessian of
            # computeHessianFD is easily indefinite: its aliased stencil)
            import scipy.linalg
            try:
                self.L_c = scipy.linalg.cholesky(np.asarray(self.V, dtype=np.float64))
            except (scipy.linalg.LinAlgError, ValueError):
                print('Non positive defintie Hessian matrix. Set to identity matrix.')
                self.V = None
        self.C_approx = np.identity(self.n_param) if self.V is None else np.asarray(self.V, dtype=np.float64)
        return self.V


class HamiltonianMonteCarlo(_GradientSampler):
    """Hamiltonian Monte Carlo (mha.py:779-884): leapfrog trajectories of num_steps steps of size step_size."""
    algo_name = "HMC"

    def __init__(self, IO_fileID, nIterations, param_init, prob_distr, C_matrix, gradient, step_size, num_steps, opt_arg):
        self._setup(IO_fileID, nIterations, param_init, None, prob_distr, opt_arg)
        self._setup_gradient(C_matrix, gradient)
        self.epsilon, self.L = float(step_size), int(num_steps)

    def _config(self):
        a0, au, a1 = self._adapt_window()
        return SamplerConfig(algo="HMC", hmc_step_size=self.epsilon, hmc_num_steps=self.L, isde_gradient=self.gradient,
                             adapt_start=a0, adapt_update=au, adapt_end=a1)

    def _finish(self):
        st = self.engine.isde_state()
        self.distr_grad_log_like_current_val, self.C_approx = st["P"][0], st["C"][0]
        _DeviceSampler._finish(self)
        self.inv_L_c = self.R


class ito_SDE(_GradientSampler):
    """Ito-SDE sampler (mha.py:886-1040) with the GradientBasedMCMC preconditioner options the device supports:
    covariance.value in {"Identity", "Matrix", "from_file", "Hessian", "Hessian_diag", "nearPD"}."""
    algo_name = "ISDE"

    def __init__(self, IO_fileID, nIterations, param_init, prob_distr, C_matrix, gradient, h, f0, opt_arg):
        self._setup(IO_fileID, nIterations, param_init, None, prob_distr, opt_arg)
        self._setup_gradient(C_matrix, gradient)
        self.h, self.f0 = float(h), float(f0)

    def _config(self):
        a0, au, a1 = self._adapt_window()
        return SamplerConfig(algo="ISDE", isde_h=self.h, isde_f0=self.f0, isde_gradient=self.gradient,
                             adapt_start=a0, adapt_update=au, adapt_end=a1)

    def _finish(self):
        st = self.engine.isde_state()
        self.P_nm, self.C_approx = st["P"][0], st["C"][0]
        _DeviceSampler._finish(self)
        self.xi_nm = self.current_val
        self.inv_L_c = self.R
