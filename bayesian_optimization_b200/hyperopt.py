"""Host side of hyper-parameter estimation: the reference's L-BFGS-B restart loop around the DEVICE
likelihood + gradient.  Follows bayes_optim/surrogate/gaussian_process/gpr.py:1058-1197 statement by
statement in behaviour (parameter list per estimation mode, log10-space bounds, warm start from a previous
``theta_``, restarts drawn from the GLOBAL numpy RNG, stagnation counter, shrinking evaluation budget, final
evaluation that fills ``env``), including quirk g4 of SURVEY.md App. A: the objective hands L-BFGS-B the
gradient w.r.t. the raw parameters although it optimises their log10 (gpr.py:1113-1121).
The "CMA" optimiser hangs forever upstream on Python 3 and is rejected here."""
from __future__ import annotations

import numpy as np
from numpy import log10
from scipy.optimize import fmin_l_bfgs_b


def hyperparameter_bounds(gp, par_list):
    """gpr.py:1042-1056."""
    bounds = []
    for name in par_list:
        if name == "theta":
            bounds.append(np.c_[gp.thetaL, gp.thetaU])
        elif name == "sigma2":
            bounds.append(np.atleast_2d([1e-5, max(1e-3, gp.y.std() ** 2)]))
        elif name in ("alpha", "noise_var"):
            bounds.append(np.atleast_2d([1e-10, 1.0 - 1e-10]))
    return np.concatenate(bounds, axis=0).astype(np.float64)


def optimize_hyperparameter(gp):
    restricted = gp.likelihood == "restricted"
    if gp.optimizer != "BFGS":
        raise NotImplementedError('optimizer="CMA" hangs in the reference on Python 3 and is not provided')

    par_list, par_len = ["theta"], [len(gp.thetaL)]
    if restricted or gp.estimation_mode == "noisy":  # gpr.py:1073-1076
        par_list += ["sigma2"]
        par_len.append(1)
    if gp.estimation_mode == "noise_estim":          # gpr.py:1078-1084
        par_list += ["noise_var" if restricted else "alpha"]
        par_len.append(1)

    bounds = hyperparameter_bounds(gp, par_list)
    log10bounds = log10(bounds)
    n_theta = len(gp.thetaL)
    if hasattr(gp, "theta_"):  # warm start, gpr.py:1095-1096
        log10theta0 = log10(gp.theta_)
    else:
        log10theta0 = (
            log10(gp.theta0) if gp.theta0 is not None else np.random.uniform(log10(gp.thetaL), log10(gp.thetaU))
        )
    if gp.estimation_mode == "noiseless" and not restricted:  # gpr.py:1103-1106
        log10param = log10theta0
    else:
        log10param = np.r_[log10theta0, np.random.uniform(log10bounds[n_theta:, 0], log10bounds[n_theta:, 1])]

    n_par = len(log10param)
    eval_budget = 200 * n_par if gp.eval_budget is None else gp.eval_budget
    llf_opt = np.inf

    def obj_func(log10param):
        gp.eval_count += 1
        param = 10.0 ** np.array(log10param)
        llf, grad = (gp.log_likelihood_restricted if restricted else gp.log_likelihood_concentrated)(param, eval_grad=True)
        return -1.0 * llf, -1.0 * np.asarray(grad, dtype=np.float64).ravel()

    gp.eval_count = 0
    wait_count = 0
    for iteration in range(gp.random_start):
        if iteration != 0:
            log10param = np.random.uniform(log10bounds[:, 0], log10bounds[:, 1])
        param_opt_, llf_opt_, info = fmin_l_bfgs_b(obj_func, log10param, bounds=log10bounds, maxfun=eval_budget)
        if iteration == 0:
            param_opt, llf_opt = param_opt_, llf_opt_
        elif llf_opt_ <= llf_opt:
            param_opt, llf_opt = param_opt_, llf_opt_
            wait_count = 0
        else:
            wait_count += 1
        if gp.verbose:
            print("restart {} takes {} evals".format(iteration + 1, info["funcalls"]))
            print("best log likekihood value: {}".format(-llf_opt))
        eval_budget -= info["funcalls"]
        if eval_budget <= 0 or wait_count >= gp.wait_iter:
            break

    optimal_param = 10.0 ** param_opt
    env = {}
    # leaves the device state AT the optimum
    optimal_llf_value = (gp.log_likelihood_restricted if restricted else gp.log_likelihood_concentrated)(optimal_param, env)
    param, i = {}, 0
    for k, name in enumerate(par_list):
        param[name] = optimal_param[i : i + par_len[k]]
        i += par_len[k]
    return param, optimal_llf_value, env
