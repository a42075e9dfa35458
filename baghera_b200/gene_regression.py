"""``analyse_normal`` — drop-in for ref: baghera_tool/gene_regression.py:30-180 on the B200 path.

Same positional signature, same side effects (summary CSV, results-log lines) and the same returned
DataFrame columns; the ``with pm.Model()`` block and ``pm.sample`` are replaced by the CUDA engine
(engine.Engine) and the batched NUTS driver (sampler.NutsSampler).  The posterior tables are
computed on the device from the stored trace (K6), not from an in-RAM PyMC3 trace.
"""
from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from . import stats as st

# run-time options that have no slot in the reference signature (set by command / cli)
# ``devices``: list of CUDA ordinals -> the chains run as independent replicas over these GPUs (replicas.py);
# None = the single ``device``
# ``stream_stats``: None = automatic (on when the fp64 trace would exceed STREAM_ABOVE_BYTES: the kernel then accumulates the
# per-gene mean / variance / P(beta > mi) in fp64 while sampling (K6), keeps the shared parameters of every draw in fp64, and the
# stored positions — only needed for the per-gene quantiles — shrink to fp32); True / False force it
OPTIONS = dict(device=0, seed=20211, trace_dtype="float64", progress=None, devices=None, stream_stats=None)
STREAM_ABOVE_BYTES = 8 << 30


def _use_stream_stats(draws, chains, D):
    if OPTIONS["stream_stats"] is not None:
        return bool(OPTIONS["stream_stats"])
    return 8 * int(draws) * int(chains) * int(D) > STREAM_ABOVE_BYTES


def build_gene_index(snp_dataset):
    """ref: gene_regression.py:56-71 — cat[j] = rank of SNP j's gene label in sorted label order,
    Mg = SNPs per gene, genes = labels (groupby iterates in sorted label order, Q6)."""
    codes, genes = pd.factorize(snp_dataset["gene"], sort=True)
    cat = codes.astype(np.int64)
    Mg = np.bincount(cat, minlength=len(genes)).astype(np.int64)
    return cat, Mg, [str(g) for g in genes]


def device_quantiles(x, qs):
    """numpy.percentile(..., interpolation='linear') along dim 0 of a device tensor x[N, K]."""
    import torch
    xs, _ = torch.sort(x, dim=0)
    n = xs.shape[0]
    out = []
    for q in qs:
        pos = (n - 1) * (q / 100.0)
        lo = int(np.floor(pos))
        hi = min(lo + 1, n - 1)
        w = pos - lo
        out.append(xs[lo] * (1.0 - w) + xs[hi] * w)
    return out


def gene_posterior_table(trace, genes, chunk=512, k6=None, head=3):
    """Per-gene posterior statistics (ref: gene_regression.py:165-177) from the device trace
    [draws, D, C]: P = mean(beta - mi > 0), bg_mean/median/var/5perc/95perc over all draws and chains.

    k6 [3, D, C] (streaming sums of the sampler kernel, fp64): P, mean and variance come from it — exact whatever the
    trace's dtype — and the trace only serves the quantiles."""
    import torch
    draws, D, C = trace.shape
    G = D - head
    mi = torch.sigmoid(trace[:, head - 1, :].double())              # [draws, C]
    cols = {k: np.empty(G) for k in ("P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc")}
    if k6 is not None:
        n = float(draws * C)
        tot = k6[:, head:, :].double().sum(dim=2)                   # [3, G]
        mean = tot[0] / n
        cols["bg_mean"][:] = mean.cpu().numpy()
        cols["bg_var"][:] = (tot[1] / n - mean * mean).cpu().numpy()
        cols["P"][:] = (tot[2] / n).cpu().numpy()
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        b = trace[:, head + g0:head + g1, :].double()               # [draws, g, C]
        flat = b.permute(0, 2, 1).reshape(draws * C, g1 - g0)
        q5, q50, q95 = device_quantiles(flat, (5.0, 50.0, 95.0))
        if k6 is None:
            P = ((b - mi[:, None, :]) > 0).double().mean(dim=(0, 2))
            cols["P"][g0:g1] = P.cpu().numpy()
            cols["bg_mean"][g0:g1] = flat.mean(dim=0).cpu().numpy()
            cols["bg_var"][g0:g1] = flat.var(dim=0, unbiased=False).cpu().numpy()
        cols["bg_median"][g0:g1] = q50.cpu().numpy()
        cols["bg_5perc"][g0:g1] = q5.cpu().numpy()
        cols["bg_95perc"][g0:g1] = q95.cpu().numpy()
    df = pd.DataFrame({"name": genes})
    for k in ("P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc"):
        df[k] = cols[k]
    return df


def analyse_normal(snps_object, output_summary_filename, output_logger, SWEEPS, TUNE, CHAINS, CORES, N_1kG,
                   fix_intercept=False):
    """Bayesian hierarchical regression with the normal model (ref: gene_regression.py:30-180).

    CORES is accepted for signature compatibility; chains run as one batch on the GPU."""
    import torch

    from . import synth
    from .engine import Engine

    snp_dataset = snps_object.table.copy().reset_index(drop=True)
    n_patients = snps_object.n_patients
    cat, Mg, genes = build_gene_index(snp_dataset)
    G = len(genes)

    logging.info("Model Evaluation Started")
    logging.info("Average stats: %f" % np.mean(snps_object.table["stats"].values))

    c = float(n_patients) / float(N_1kG)
    from .replicas import sample_on_devices
    chi2 = snp_dataset["stats"].values.astype(np.float32)
    ld = snp_dataset["l"].values.astype(np.float32)
    cat32 = cat.astype(np.int32)

    def make_engine(device, n_chains):
        return Engine(chi2, ld, cat32, G, c, int(n_chains), model="gene", fix_intercept=bool(fix_intercept), device=device)

    devices = OPTIONS["devices"] or [OPTIONS["device"]]
    stream = _use_stream_stats(SWEEPS, CHAINS, G + 3)
    dt = torch.float32 if (OPTIONS["trace_dtype"] == "float32" or stream) else torch.float64
    start = synth.jittered_start(G + 3, int(CHAINS), seed=OPTIONS["seed"] + 1, model="gene")
    # nuts_kwargs=dict(target_accept=0.90)
    torch.cuda.reset_peak_memory_stats()
    out, n_launches = sample_on_devices(make_engine, start, int(SWEEPS), int(TUNE), devices, OPTIONS["seed"],
                                        trace_dtype=dt, progress=OPTIONS["progress"], target_accept=0.90, stream_stats=stream)
    trace = out["trace"]                                           # [draws, D, C], unconstrained space
    n_div = int(out["stats"][int(TUNE):, 3].sum())
    if n_div:
        logging.warning("There were %d divergences after tuning. Increase `target_accept` or reparameterize." % n_div)

    heads = out["head_trace"] if stream else trace[:, :3, :].double()   # shared parameters of every draw, fp64
    tW = torch.exp(heads[:, 0, :]).cpu().numpy()                    # W  = exp(W_log__)
    te = heads[:, 1, :].cpu().numpy()
    tmi = torch.sigmoid(heads[:, 2, :]).cpu().numpy()               # mi = sigmoid(mi_logodds__)
    w = torch.as_tensor(Mg / float(N_1kG), dtype=torch.float64, device=trace.device)
    acc = torch.zeros((trace.shape[0], trace.shape[2]), dtype=torch.float64, device=trace.device)
    for g0 in range(0, G, 512):                                     # herTOT = sum(beta * Mg / N_1kG), gene chunks (:83)
        g1 = min(G, g0 + 512)
        acc += torch.einsum("dgc,g->dc", trace[:, 3 + g0:3 + g1, :].double(), w[g0:g1])
    therTOT = acc.cpu().numpy()

    if CHAINS > 1:
        logging.info("evaluating Gelman-Rubin")
        GR = {"mi": st.gelman_rubin(tmi)}
        output_logger.info("DIAGNOSTIC (gelman-rubin) " + str(GR) + "\n\t"
                           + "(If this number is >> 1 the method has some convergence problem, \n\t try increasing the number of s and b)")

    logging.info("Writing output")
    su = st.summary({"mi": tmi, "herTOT": therTOT, "e": te, "W": tW})         # (:120-125)
    su.to_csv(output_summary_filename, sep=",", mode="w")

    output_logger.info(" Intercept: " + str(np.mean(te)) + " (sd= " + str(np.std(te)) + ")\n")
    output_logger.info(" W: " + str(np.mean(tW)) + " (sd= " + str(np.std(tW)) + ")\n")
    output_logger.info(" heritability from genes: " + str(np.median(therTOT)) + " (sd= " + str(np.std(therTOT)) + ")\n")
    mi_mean, mi_median, mi_std = np.mean(tmi), np.median(tmi), np.std(tmi)
    mi_5perc, mi_95perc = np.percentile(tmi, 5), np.percentile(tmi, 95)
    output_logger.info(" Heritability: " + str(mi_mean) + " (std= " + str(mi_std) + ")\n" + "[ 5th perc= "
                       + str(mi_5perc) + "," + " 95 perc= " + str(mi_95perc) + "]\n")

    df = gene_posterior_table(trace, genes, k6=out["k6"] if stream else None)
    df["mi_mean"] = mi_mean
    df["mi_median"] = mi_median
    analyse_normal.last_run = dict(seconds_tune=out["seconds_tune"], seconds_draws=out["seconds_draws"],
                                   seconds_total=out.get("seconds_total"), ticks=out.get("ticks"),
                                   leapfrogs=out["leapfrogs"], stats=out["stats"], kernel_launches=n_launches,
                                   devices=out["devices"], stream_stats=stream, trace_dtype=str(dt),
                                   peak_hbm_bytes=int(torch.cuda.max_memory_allocated()))
    logging.info("analysis done")
    return df


def analyse_gamma(snps_object, output_summary_filename, output_logger, SWEEPS, TUNE, CHAINS, CORES, N_1kG,
                  fix_intercept=False):
    """Bayesian hierarchical regression with the gamma model (ref: gene_regression.py:183-325).

    ``e ~ Normal(1, 0.001)``, ``mi ~ Beta(1,1)``, ``beta ~ Gamma(alpha=mi, beta=N_1kG)``, likelihood mean
    ``n_patients * beta[cat] * l + e`` (:229-247).  Free vector q = [e, mi_logodds__, beta_log__...].  The
    likelihood kernel is the normal model's; only the prior / transform closing formulas differ.
    Outputs as the reference: summary rows mi, herTOT, e (:274-280); per-gene columns from N_1kG * beta (:283),
    P = mean(beta - mi / N_1kG > 0) (:232, :314)."""
    import torch

    from . import synth
    from .engine import Engine

    snp_dataset = snps_object.table.copy().reset_index(drop=True)
    n_patients = snps_object.n_patients
    cat, Mg, genes = build_gene_index(snp_dataset)
    G = len(genes)

    logging.info("Model Evaluation Started")
    logging.info("Average stats: %f" % np.mean(snps_object.table["stats"].values))

    from .replicas import sample_on_devices
    chi2 = snp_dataset["stats"].values.astype(np.float32)
    ld = snp_dataset["l"].values.astype(np.float32)
    cat32 = cat.astype(np.int32)

    def make_engine(device, n_chains):
        return Engine(chi2, ld, cat32, G, float(n_patients), int(n_chains), model="gamma", fix_intercept=bool(fix_intercept),
                      device=device, gamma_rate=float(N_1kG))

    start = synth.jittered_start(G + 2, int(CHAINS), seed=OPTIONS["seed"] + 1, model="gamma")
    start[:, 2:] += np.log(1290028.0 / float(N_1kG))      # test value of Gamma(mi, N_1kG) = 0.5 / N_1kG
    dt = torch.float32 if OPTIONS["trace_dtype"] == "float32" else torch.float64
    out, n_launches = sample_on_devices(make_engine, start, int(SWEEPS), int(TUNE), OPTIONS["devices"] or [OPTIONS["device"]],
                                        OPTIONS["seed"], trace_dtype=dt, progress=OPTIONS["progress"], target_accept=0.90)
    trace = out["trace"]                                           # [draws, D, C], unconstrained space
    n_div = int(out["stats"][int(TUNE):, 3].sum())
    if n_div:
        logging.warning("There were %d divergences after tuning. Increase `target_accept` or reparameterize." % n_div)

    te = trace[:, 0, :].double().cpu().numpy()
    tmi_dev = torch.sigmoid(trace[:, 1, :].double())               # [draws, C]
    tmi = tmi_dev.cpu().numpy()
    w = torch.as_tensor(Mg.astype(np.float64), dtype=torch.float64, device=trace.device)
    therTOT = np.zeros(te.shape)
    chunk = 512
    cols = {k: np.empty(G) for k in ("P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc")}
    draws, _, C = trace.shape
    acc = torch.zeros((draws, C), dtype=torch.float64, device=trace.device)
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        b = torch.exp(trace[:, 2 + g0:2 + g1, :].double())         # beta = exp(beta_log__)  [draws, g, C]
        acc += torch.einsum("dgc,g->dc", b, w[g0:g1])              # herTOT = sum(beta * Mg)            (:233)
        P = ((b - tmi_dev[:, None, :] / float(N_1kG)) > 0).double().mean(dim=(0, 2))     # diff (:232)
        flat = (float(N_1kG) * b).permute(0, 2, 1).reshape(draws * C, g1 - g0)           # d["beta"] (:283)
        q5, q50, q95 = device_quantiles(flat, (5.0, 50.0, 95.0))
        cols["P"][g0:g1] = P.cpu().numpy()
        cols["bg_mean"][g0:g1] = flat.mean(dim=0).cpu().numpy()
        cols["bg_var"][g0:g1] = flat.var(dim=0, unbiased=False).cpu().numpy()
        cols["bg_median"][g0:g1] = q50.cpu().numpy()
        cols["bg_5perc"][g0:g1] = q5.cpu().numpy()
        cols["bg_95perc"][g0:g1] = q95.cpu().numpy()
    therTOT = acc.cpu().numpy()

    if CHAINS > 1:
        logging.info("evaluating Gelman-Rubin")
        GR = {"mi": st.gelman_rubin(tmi)}
        output_logger.info("DIAGNOSTIC (gelman-rubin) " + str(GR) + "\n"
                           + "(If this number is >> 1 the method has some convergence problem, \n try increasing the number of s and b)")

    logging.info("Writing output")
    su = st.summary({"mi": tmi, "herTOT": therTOT, "e": te})                  # (:274-280)
    su.to_csv(output_summary_filename, sep=",", mode="w")

    output_logger.info(" Intercept: " + str(np.mean(te)) + " (sd= " + str(np.std(te)) + ")\n")
    output_logger.info(" heritability from genes: " + str(np.median(therTOT)) + " (sd= " + str(np.std(therTOT)) + ")\n")
    mi_mean, mi_median, mi_std = np.mean(tmi), np.median(tmi), np.std(tmi)
    mi_5perc, mi_95perc = np.percentile(tmi, 5), np.percentile(tmi, 95)
    output_logger.info(" Heritability: " + str(mi_mean) + " (std= " + str(mi_std) + ")\n" + "[ 5th perc= "
                       + str(mi_5perc) + "," + " 95 perc= " + str(mi_95perc) + "]\n")

    df = pd.DataFrame({"name": genes})
    for k in ("P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc"):
        df[k] = cols[k]
    df["mi_mean"] = mi_mean
    df["mi_median"] = mi_median
    analyse_gamma.last_run = dict(seconds_tune=out["seconds_tune"], seconds_draws=out["seconds_draws"],
                                  leapfrogs=out["leapfrogs"], stats=out["stats"], kernel_launches=n_launches,
                                  devices=out["devices"])
    logging.info("analysis done")
    return df
