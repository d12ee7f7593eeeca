"""``Sampler``: the drop-in boundary of the MCMC path (host-side mirror of ``pybitup/inference_problem.py:6-109``).

``Sampler(IO_fileID, prob_dist, algo).sample(sample_init)`` parses the ``Algorithm`` block exactly as the reference
does (proposal covariance :37-49, n_iterations :53, save_eval_freq :57-60, estimate_* :64-73, per-algorithm keys
:79-107) and instantiates the sampler classes of ``metropolis_hastings_algorithms`` -- which here drive the CUDA
engine.  Engine-only keys of the block (n_chains, seed, init_jitter, device, launch_iterations, save_all_chains,
am_welford, chain_offset, pooled_covariance, pooled_update_it) are forwarded through ``opt_arg``.
"""
import numpy as np
import pandas as pd

from . import metropolis_hastings_algorithms as mha

ENGINE_KEYS = ("n_chains", "seed", "init_jitter", "device", "launch_iterations", "save_all_chains", "am_welford",
               "chain_offset", "pooled_covariance", "pooled_update_it")


class Sampler:
    def __init__(self, IO_fileID, prob_dist, algo):
        self.prob_dist = prob_dist
        self.algo = algo
        self.IO_fileID = IO_fileID
        self.opt_arg = {}
        av_algo = {'RWMH': "random-walk Metropolis-Hastings", 'AMH': "adaptive Metropolis-Hastings",
                   'DR': "delayed-rejection Metropolis-Hastings", 'DRAM': "delayed-rejection adative Metropolis-Hastings",
                   'HMC': "Hamiltonian Monte Carlo", 'ISDE': "Ito-SDE"}
        self.algo_name = self.algo['name']
        if self.algo_name not in av_algo:
            raise ValueError('Algorithm "{}" unknown.'.format(self.algo_name))
        print("Running {} algorithm on the CUDA engine.".format(av_algo[self.algo_name]))
        if self.algo_name in ("RWMH", "AMH", "DR", "DRAM"):
            cov = self.algo['proposal']['covariance']
            if cov['type'] == "diag":
                if isinstance(cov['value'], str):
                    A = np.transpose(pd.read_csv(cov['value'], header=None).values)
                    self.proposal_cov = np.diag(A[0])
                else:
                    self.proposal_cov = np.diag(cov['value'])
            elif cov['type'] == "full":
                self.proposal_cov = np.array(cov['value'])
            else:
                raise ValueError("Invalid InferenceAlgorithmProposalConvarianceType name {}".format(cov['type']))
        self.n_iterations = int(self.algo['n_iterations'])
        self.opt_arg['save_eval_freq'] = int(self.algo['save_eval_freq']) if self.algo.get("save_eval_freq") is not None else 10
        self.opt_arg["estimate_max_distr"] = self.algo.get("estimate_max_distr") == "yes"
        self.opt_arg["estimate_sigma"] = self.algo.get("estimate_sigma") == "yes"
        for k in ENGINE_KEYS:
            if self.algo.get(k) is not None:
                self.opt_arg[k] = self.algo[k]

    def sample(self, sample_init):
        a, n = self.algo, self.algo_name
        if n == "RWMH":
            run = mha.MetropolisHastings(self.IO_fileID, self.n_iterations, sample_init, self.proposal_cov, self.prob_dist, self.opt_arg)
        elif n == "AMH":
            run = mha.AdaptiveMetropolisHastings(self.IO_fileID, self.n_iterations, sample_init, self.proposal_cov, self.prob_dist,
                                                 int(a['AMH']['starting_it']), int(a['AMH']['updating_it']), a['AMH']['eps_v'], self.opt_arg)
        elif n == "DR":
            run = mha.DelayedRejectionMetropolisHastings(self.IO_fileID, self.n_iterations, sample_init, self.proposal_cov, self.prob_dist,
                                                         a['DR']['gamma'], self.opt_arg)
        elif n == "DRAM":
            run = mha.DelayedRejectionAdaptiveMetropolisHastings(self.IO_fileID, self.n_iterations, sample_init, self.proposal_cov,
                                                                 self.prob_dist, int(a['DRAM']['starting_it']), int(a['DRAM']['updating_it']),
                                                                 a['DRAM']['eps_v'], a['DRAM']['gamma'], self.opt_arg)
        elif n == "HMC":
            run = mha.HamiltonianMonteCarlo(self.IO_fileID, self.n_iterations, sample_init, self.prob_dist, a['covariance'], a['gradient'],
                                            a['HMC']['step_size'], a['HMC']['num_steps'], self.opt_arg)
        else:
            run = mha.ito_SDE(self.IO_fileID, self.n_iterations, sample_init, self.prob_dist, a['covariance'], a['gradient'],
                              a['ISDE']['h'], a['ISDE']['f0'], self.opt_arg)
        self.run_MCMCM = run
        run.run_algorithm()
        return run
