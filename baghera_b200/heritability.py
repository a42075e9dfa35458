"""``gw_normal`` — drop-in for ref: baghera_tool/heritability.py:26-93 (genome-wide model, D = 2)."""
from __future__ import annotations

import logging

import numpy as np

from . import stats as st
from .gene_regression import OPTIONS


def gw_normal(snps_object, output_summary_filename, output_logger, SWEEPS, TUNE, CHAINS, CORES, N_1kG,
              fix_intercept=False):
    import torch

    from . import synth
    from .engine import Engine

    logging.info("Bayesian analysis started")
    snps_object.table = snps_object.table.reset_index(drop=True)
    nPATIENTS = snps_object.table["sample_size"][0]                       # (:42)
    c = float(nPATIENTS) / float(N_1kG)
    tab = snps_object.table
    from .replicas import sample_on_devices
    chi2 = tab["stats"].values.astype(np.float32)
    ld = tab["l"].values.astype(np.float32)

    def make_engine(device, n_chains):
        return Engine(chi2, ld, None, 1, c, int(n_chains), model="gw", fix_intercept=bool(fix_intercept), device=device)

    start = synth.jittered_start(2, int(CHAINS), seed=OPTIONS["seed"] + 1, model="gw")
    if fix_intercept:
        start[:, 0] -= 1.0        # interval-transformed Uniform: test point = midpoint -> 0 in the unconstrained space
    out, n_launches = sample_on_devices(make_engine, start, int(SWEEPS), int(TUNE), OPTIONS["devices"] or [OPTIONS["device"]],
                                        OPTIONS["seed"], progress=OPTIONS["progress"], target_accept=0.90)
    trace = out["trace"]
    x0 = trace[:, 0, :].double()
    if fix_intercept:
        te = (0.9999 + (1.0000001 - 0.9999) * torch.sigmoid(x0)).cpu().numpy()   # e ~ Uniform(0.9999, 1.0000001) (:46)
    else:
        te = x0.cpu().numpy()
    tmi = torch.sigmoid(trace[:, 1, :].double()).cpu().numpy()

    if CHAINS > 1:
        output_logger.info("evaluating Gelman-Rubin")
        GR = {"mi": st.gelman_rubin(tmi)}
        output_logger.info("DIAGNOSTIC (gelman-rubin) " + str(GR) + "\n"
                           + "(If this number is >> 1 the method has some convergence problem, \n try increasing the number of s and b)")
    su = st.summary({"e": te, "mi": tmi})                                # pm.summary(trace, extend=True, ...) (:74)
    logging.info("Writing output")
    su.to_csv(output_summary_filename, sep=",", mode="w")

    mi_mean = np.mean(tmi)
    mi_median = np.mean(tmi)                                             # sic (Q10)
    output_logger.info(" heritability mean: " + str(mi_mean) + "\n")
    output_logger.info(" heritability median: " + str(mi_median) + "\n")
    output_logger.info(" heritability std: " + str(np.std(tmi)) + "\n")
    output_logger.info(" heritability 95perc: " + str(np.percentile(tmi, 95)) + "\n")
    output_logger.info(" heritability 5perc: " + str(np.percentile(tmi, 5)) + "\n")
    intercept = np.mean(te)
    gw_normal.last_run = dict(seconds_tune=out["seconds_tune"], seconds_draws=out["seconds_draws"],
                              leapfrogs=out["leapfrogs"], stats=out["stats"], kernel_launches=n_launches,
                              devices=out["devices"])
    return [intercept, mi_mean]
