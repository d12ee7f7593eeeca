"""JSON-driven entry points: host-side mirror of ``pybitup/solve_problem.py`` for the sampling and Monte-Carlo
propagation paths.

``Sampling(json).sample({id: model})`` (reference :84-358) and ``Propagation(json).propagate({id: model})``
(:361-647, ``emulator == "None"``) keep the reference's input schema, data loading and output files; the sampler loop
and the per-sample model loop run on the CUDA engine.  ``Sampling.Distribution`` (sampling an analytic density, :134-164)
is lowered to a device pseudo-model; branches of the reference outside the hot path (``ComputeAnalyticalDistribution``,
the PCE emulator) raise ``NotImplementedError``.

One process per GPU (``torchrun``): every rank parses the input and builds the problem, rank 0 alone owns the output
files (the other ranks' handles point at ``os.devnull``), chains / samples are sharded by global index.
"""
import json
import os
import pathlib
import pickle
import shutil
import time

import numpy as np
import pandas as pd

from . import bayesian_inference, distributions, inference_problem
from .distributed import Reducer, shard_range


def strip_json_comments(text):
    """Remove // and /* */ comments outside strings (the reference uses the jsmin package, :35-37)."""
    out, i, n, in_str = [], 0, len(text), False
    while i < n:
        c = text[i]
        if in_str:
            out.append(c)
            if c == "\\" and i + 1 < n:
                out.append(text[i + 1])
                i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
            out.append(c)
        elif text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
            continue
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            i = n if j < 0 else j + 2
            continue
        else:
            out.append(c)
        i += 1
    return "".join(out)


class SolveProblem:
    def __init__(self, input_file_name):
        with open("{}".format(input_file_name)) as js_file:
            self.user_inputs = json.loads(strip_json_comments(js_file.read()))
        self.IO_util = {'path': {}, 'fileID': {}}
        p_cwd = pathlib.Path.cwd()
        self.IO_util['path']['cwd'] = p_cwd
        self.IO_util['path']['out_folder'] = pathlib.Path(p_cwd, "output")
        self.IO_util['path']['out_folder'].mkdir(exist_ok=True)
        self.IO_util['path']['fun_eval_folder'] = pathlib.Path(self.IO_util['path']['out_folder'], "model_eval")
        self.IO_util['path']['out_data'] = pathlib.Path(self.IO_util['path']['out_folder'], "output.txt")
        from .distributed import ensure_process_group
        ensure_process_group()                              # torchrun: one process per GPU, NCCL
        self.rank = Reducer().rank
        self.IO_util['fileID']['out_data'] = self._open(self.IO_util['path']['out_data'], 'a+')
        if self.rank == 0:
            try:
                shutil.copy(input_file_name, self.IO_util['path']['out_folder'])
            except shutil.SameFileError:
                pass

    def _open(self, path, mode):
        """Output files belong to rank 0; the other ranks of a torchrun job get a sink (they would truncate its files)."""
        return open(path, mode) if self.rank == 0 else open(os.devnull, "w+")

    def close(self):
        for f in self.IO_util['fileID'].values():
            if not f.closed:
                f.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sampling(SolveProblem):
    def __init__(self, input_file_name):
        SolveProblem.__init__(self, input_file_name)
        new_file_list = {'MChains': "mcmc_chain.csv", 'MChains_reparam': "mcmc_chain_reparam.csv"}
        if self.user_inputs["Sampling"].get('Algorithm') is not None:
            self.algo_inputs = self.user_inputs["Sampling"].get('Algorithm')
            if self.algo_inputs.get("estimate_sigma") == "yes":
                new_file_list['estimated_sigma'] = "estimated_sigma.csv"
            if self.algo_inputs.get("ISDE") is not None or self.algo_inputs.get("HMC") is not None:
                new_file_list['aux_variables'] = "aux_variables.csv"
            if self.algo_inputs.get("estimate_max_distr") == "yes":
                new_file_list['estimate_arg_max_val_distr'] = "arg_MAP_estimation.csv"
                new_file_list['estimate_max_val_distr'] = "MAP_estimation.csv"
        for key, name in new_file_list.items():
            self.IO_util['path'][str(key)] = pathlib.Path(self.IO_util['path']['out_folder'], name)
            self.IO_util['fileID'][str(key)] = self._open(self.IO_util['path'][str(key)], "w+")

    @staticmethod
    def _read_data(c_data, model):
        """Data block 'ReadFromFile' (:206-251), including the file-name sequence _0, _0, _1, ... of SURVEY Q14."""
        if c_data['Type'] != "ReadFromFile":
            raise ValueError("Invalid DataType {} (GenerateSynthetic is broken in the reference, :262-271; "
                             "write the data to a CSV file)".format(c_data['Type']))
        if c_data.get("n_runs") is not None:
            n_runs, app_name = c_data['n_runs'], "_0.csv"
        else:
            n_runs, app_name = 1, ".csv"
        data_name = c_data['FileName']
        reader = pd.read_csv(data_name + app_name)
        x = reader[c_data['xField'][0]].values
        model.x = x
        y_tot, std_y, dataName = {}, np.array([]), ""
        for c_run in range(n_runs):
            y, std_y = [], []
            reader = pd.read_csv(data_name + app_name)
            for nfield, yfield in enumerate(c_data['yField']):
                y = np.concatenate((y, reader[yfield].values))
                std_y = np.concatenate((std_y, reader[c_data['sigmaField'][nfield]].values))
                dataName = c_data['yField'][0] + "_" + data_name
            y_tot[c_run] = y
            app_name = "_" + str(c_run) + '.csv'
        return bayesian_inference.Data(dataName, x, y_tot, std_y)

    def sample(self, models={}):
        if self.user_inputs.get("Sampling") is None:
            raise ValueError('Ask for sampling distribution but no Sampling inputs were provided')
        S = self.user_inputs["Sampling"]
        if S.get("Distribution") is not None:
            return self._sample_distribution(S)
        if S.get("BayesianPosterior") is None:
            raise ValueError('No samling distribution provided or invalid name.')
        self.IO_util['path']['fun_eval_folder'].mkdir(exist_ok=True)
        BP = S["BayesianPosterior"]
        data, model_id = {}, None
        for data_set, model_id in enumerate(models.keys()):
            c_model = BP['Model'][data_set]
            if c_model.get("input_file") is not None:
                models[model_id].input_file_name = c_model['input_file']
            models[model_id].param = np.array(c_model['param_values'])
            models[model_id].param_nom = np.array(c_model['param_values'])
            models[model_id].param_names = c_model['param_names']
            data[model_id] = self._read_data(BP['Data'][data_set], models[model_id])
        if self.rank == 0:
            with open(pathlib.Path(self.IO_util['path']['out_folder'], "data.bin"), 'wb') as f:
                pickle.dump(data, f)
        unpar_name = list(BP['Prior']['Param'].keys())
        for mid in models.keys():
            models[mid].unpar_name = unpar_name
            models[mid].unpar_name_dict = {name: i for i, name in enumerate(unpar_name)}
        vals = list(BP['Prior']['Param'].values())
        unpar_init_val = np.array(models[model_id].parametrization_forward([v['initial_val'] for v in vals]), dtype=np.float64)
        init_file = pathlib.Path(self.IO_util['path']['cwd'], "mcmc_chain_init.csv")          # manual restart (:312-318)
        if init_file.exists():
            raw = pd.read_csv(init_file, header=None).values
            unpar_init_val = np.array(models[model_id].parametrization_forward(np.float64(raw[0, :])))
        prior = distributions.set_probability_dist([BP['Prior']['Distribution']],
                                                   [[v['prior_name'] for v in vals], [v['prior_param'] for v in vals]],
                                                   len(unpar_init_val))
        gamma = BP['Likelihood'].get('gamma', 1.0) if BP.get('Likelihood') else 1.0
        likelihood = bayesian_inference.Likelihood(data, models, gamma)
        self.sample_dist = bayesian_inference.BayesianPosterior(prior, likelihood, models[model_id], unpar_init_val)
        if S.get('Algorithm') is not None:
            sampler = inference_problem.Sampler(self.IO_util, self.sample_dist, S["Algorithm"])
            self.sampler = sampler.sample(unpar_init_val)
        if S.get('ComputeAnalyticalDistribution') is not None:
            raise NotImplementedError("ComputeAnalyticalDistribution (grid evaluation for plots) is outside the device hot path")
        return getattr(self, "sampler", None)


    def _sample_distribution(self, S):
        """Sampling.Distribution: sample a known density (solve_problem.py:134-164).  Unlike the reference, which
        crashes at the end of such a run (its covariance step reads a chain file that this branch never writes,
        SURVEY.md 8f-2), both chain files are written."""
        D = S['Distribution']
        distr_param = []
        for h in D['hyperparameters']:
            distr_param.append(pd.read_csv(h, header=None).values if isinstance(h, str) else h)
        if isinstance(D["init_val"], str):
            unpar_init_val = np.transpose(pd.read_csv(D["init_val"], header=None).values)[0].astype(np.float64)
        else:
            unpar_init_val = np.array(D["init_val"], dtype=np.float64)
        self.sample_dist = distributions.set_probability_dist([D['name']], distr_param, D['n_rand_var'])
        if S.get('Algorithm') is not None:
            sampler = inference_problem.Sampler(self.IO_util, self.sample_dist, S["Algorithm"])
            self.sampler = sampler.sample(unpar_init_val)
        return getattr(self, "sampler", None)


class Propagation(SolveProblem):
    """Monte-Carlo propagation of parameter samples through registered device models (:361-647)."""

    def __init__(self, input_file_name):
        SolveProblem.__init__(self, input_file_name)
        self.IO_util['path']['out_folder_prop'] = pathlib.Path(self.IO_util['path']['out_folder'], "propagation")
        self.IO_util['path']['out_folder_prop'].mkdir(exist_ok=True)

    def propagate(self, models=[], device=None):
        """device: CUDA device of this process; default LOCAL_RANK under torchrun, else torch's current device."""
        if self.user_inputs.get("Propagation") is None:
            raise ValueError('Ask for uncertainty propagation but no Propagation inputs were provided')
        from .engine import ChainEngine, ProblemSpec, SamplerConfig
        from .propagation import DevicePropagation
        P = self.user_inputs["Propagation"]
        unpar_inputs = P["Uncertain_param"]
        unpar = {}
        for name, c in unpar_inputs.items():
            if c["filename"] == "None":
                raise NotImplementedError("labelled input distributions belong to the PCE emulator path")
            # pd.read_csv(filename) takes the first chain row as header: reproduced
            unpar[name] = pd.read_csv(c["filename"]).values[:, c["field"]].astype(np.float64)
        if P["Model"][0]["emulator"] != "None":
            raise NotImplementedError("only emulator == \"None\" (plain Monte Carlo) runs on the device path")
        names = list(unpar_inputs.keys())
        n_sample = P["Model_evaluation"].get("Sample_number") if P.get("Model_evaluation") else None
        n_sample = len(unpar[names[-1]]) if n_sample is None else int(n_sample)
        X_all = np.stack([unpar[name][:n_sample] for name in names], axis=1)
        from .distributed import local_device
        device = local_device() if device is None else int(device)
        red = Reducer()                                   # one process per GPU: samples sharded, statistics all-reduced
        if red.device_sum is not None and red.device.index != device:
            raise RuntimeError("rank {} propagates on cuda:{} but its process group is bound to cuda:{}".format(
                red.rank, device, red.device.index))
        if n_sample < red.world_size:
            raise ValueError("Sample_number ({}) must be at least the number of ranks ({})".format(n_sample, red.world_size))
        first, count = shard_range(n_sample, red.world_size, red.rank)
        self.IO_util['path']['fun_eval_folder'].mkdir(exist_ok=True)
        results = {}
        t1 = time.perf_counter()
        for model_num, model_id in enumerate(models.keys()):
            c_model = P["Model"][model_num]
            m = models[model_id]
            dp_in = c_model["design_points"]
            m.x = pd.read_csv(dp_in["filename"])[dp_in["field"]].values
            m.param = np.array(c_model['param_values'])
            m.param_nom = np.array(c_model['param_values'])
            m.param_names = c_model['param_names']
            m.unpar_name = names
            N = m.size_x()
            design = m.device_design()
            spec = ProblemSpec(model=m.device_model, theta_nom=np.array(m.param_nom, dtype=np.float64),
                               unc_idx=m.uncertain_index(), y=np.zeros((1, N)), std_y=np.ones(N),
                               lb=np.full(len(names), -np.inf), ub=np.full(len(names), np.inf), gamma=1.0, design=design)
            with ChainEngine(spec, SamplerConfig(algo="RWMH"), 1, device=device) as eng:
                with DevicePropagation(eng, count) as dp:
                    dp.evaluate(X_all[first:first + count])
                    st = dp.statistics((2.5, 97.5), reduce=red if red.active else None, n_total=n_sample)
                    if red.rank == 0:
                        fe = dp.evals(0, min(count, int(P.get("Model_evaluation", {}).get("save_evaluations", 1000))))
                        with open(pathlib.Path(self.IO_util['path']['fun_eval_folder'], f"{model_id}_fun_eval-0.npy"), 'ab') as f:
                            np.savetxt(f, fe, fmt="%f", delimiter=",")       # text rows, as the reference writes (:611)
            results[model_id] = st
            if red.rank == 0:
                out = self.IO_util['path']['out_folder']
                pd.DataFrame({"x": m.x, "mean": st["mean"], "lower_bound": st["min"], "upper_bound": st["max"]}).to_csv(
                    pathlib.Path(out, f"{model_id}_interval.csv"), index=None)
                pd.DataFrame({"x": m.x, "CI_lb": st["ci_lb"], "CI_ub": st["ci_ub"]}).to_csv(
                    pathlib.Path(out, f"{model_id}_CI.csv"), index=None)
        print("Elapsed time: {} sec".format(time.strftime("%H:%M:%S", time.gmtime(time.perf_counter() - t1))))
        self.results = results
        return results
