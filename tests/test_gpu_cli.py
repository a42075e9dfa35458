"""GPU end-to-end: annotated SNP CSV -> `baghera-tool gene-heritability / gw-heritability` -> the three
output artefacts of the reference (results CSV, summary CSV, results log)."""
import numpy as np
import pandas as pd
import pytest

from baghera_b200 import synth

pytestmark = pytest.mark.gpu

RESULT_COLUMNS = ["name", "chrom", "n_snps", "ld_min", "ld_max", "ld_var", "stats_min", "stats_max", "stats_avg",
                  "P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc", "mi_mean", "mi_median", "h2g"]
SUMMARY_COLUMNS = ["mean", "sd", "mc_error", "hpd_2.5", "hpd_97.5", "median", "5", "50", "95", "n_eff", "Rhat"]


@pytest.fixture(scope="module")
def snp_csv(tmp_path_factory):
    df, truth = synth.make_snp_table(12000, 30, seed=8)
    d = tmp_path_factory.mktemp("cli")
    p = d / "snps.csv"
    df.to_csv(p, index=False)
    return d, str(p), df, truth


def test_gene_heritability_end_to_end(snp_csv):
    from baghera_b200.cli import main
    d, path, raw, truth = snp_csv
    out_g, out_s, out_l = str(d / "genes.csv"), str(d / "summary.csv"), str(d / "log.txt")
    rc = main(["gene-heritability", path, out_g, out_s, out_l, "--sweeps", "600", "-b", "600", "--n-chains", "4",
               "--n-cores", "4", "-m", "normal", "--seed", "3"])
    assert rc == 0
    res = pd.read_csv(out_g)
    assert list(res.columns) == RESULT_COLUMNS
    kept = raw[raw["maf"] > 0.01]
    counts = kept.groupby("gene").size()
    expect = set(counts[counts >= 10].index) | {"NonCoding"}
    assert set(res["name"]) == expect                                     # inner merge drops sub-threshold genes (Q2)
    assert res["P"].between(0, 1).all() and res["P"].is_monotonic_increasing          # sorted by P (numeric)
    assert np.allclose(res["h2g"], res["bg_mean"] * res["n_snps"] / 1290028.0)        # command.py:118-119
    assert (res["bg_5perc"] <= res["bg_median"]).all() and (res["bg_median"] <= res["bg_95perc"]).all()
    su = pd.read_csv(out_s, index_col=0)
    assert list(su.index) == ["mi", "herTOT", "e", "W"] and list(su.columns) == SUMMARY_COLUMNS
    assert su["Rhat"].max() < 1.05 and su["n_eff"].min() > 100
    assert abs(su.loc["e", "mean"] - truth["e"]) < 0.15                   # data were simulated with e = 1.02
    assert 0.05 < su.loc["mi", "mean"] < 0.45 and 0.05 < su.loc["W", "mean"] < 0.6
    # heritability-enriched genes (beta = mi + 1) must come out on top
    names = truth["names"]
    hot = [names[i] for i in np.flatnonzero(truth["beta"] > truth["mi"] + 0.9) if names[i] in expect]
    if hot:
        assert res.set_index("name").loc[hot, "P"].min() > 0.9
    log = open(out_l).read()
    for needle in ("[INFO]: BAGHERA resutls", "gene level regression, model: normal", "[INFO]:  Sample size 361194",
                   "After baghera init filter.\n\t Number of SNPs: %d" % len(kept), "Output gene table initialised:",
                   "DIAGNOSTIC (gelman-rubin) {'mi':", " Intercept: ", " W: ", " heritability from genes: ",
                   " Heritability: ", " Non coding heritability : ", " Coding heritability : "):
        assert needle in log, needle


def test_gw_heritability_end_to_end(snp_csv):
    from baghera_b200.cli import main
    d, path, raw, truth = snp_csv
    out_s, out_l = str(d / "gw_summary.csv"), str(d / "gw_log.txt")
    rc = main(["gw-heritability", path, out_s, out_l, "--sweeps", "500", "-b", "500", "--n-chains", "4", "--seed", "4"])
    assert rc == 0
    su = pd.read_csv(out_s, index_col=0)
    assert list(su.index) == ["e", "mi"] and list(su.columns) == SUMMARY_COLUMNS
    assert su["Rhat"].max() < 1.05
    assert abs(su.loc["e", "mean"] - truth["e"]) < 0.2 and 0.05 < su.loc["mi", "mean"] < 0.5
    log = open(out_l).read()
    for needle in ("genome-wide regression, model: normal", " heritability mean: ", " heritability 95perc: "):
        assert needle in log, needle


def test_gene_heritability_gamma_end_to_end(snp_csv):
    """-m gamma (ref: gene_regression.py:183-325): same artefacts; summary rows mi, herTOT, e; per-gene columns
    from N_1kG * beta."""
    from baghera_b200.cli import main
    d, path, raw, truth = snp_csv
    out_g, out_s, out_l = str(d / "g2.csv"), str(d / "s2.csv"), str(d / "l2.txt")
    rc = main(["gene-heritability", path, out_g, out_s, out_l, "-m", "gamma", "--sweeps", "400", "-b", "600",
               "--n-chains", "4", "--seed", "5"])
    assert rc == 0
    res = pd.read_csv(out_g)
    assert list(res.columns) == RESULT_COLUMNS
    assert np.all(np.isfinite(res[["P", "bg_mean", "bg_median", "bg_var", "bg_5perc", "bg_95perc", "h2g"]].values))
    assert (res["bg_mean"] > 0).all()                                      # beta ~ Gamma is positive
    assert res["P"].between(0, 1).all()
    su = pd.read_csv(out_s, index_col=0)
    assert list(su.index) == ["mi", "herTOT", "e"] and list(su.columns) == SUMMARY_COLUMNS
    assert abs(su.loc["e", "mean"] - 1.0) < 0.01                           # e ~ Normal(1, 0.001)
    assert 0.0 < su.loc["mi", "mean"] < 1.0
    # the enriched genes (largest simulated effects) still rank high
    names = truth["names"]
    kept = raw[raw["maf"] > 0.01]
    counts = kept.groupby("gene").size()
    expect = set(counts[counts >= 10].index) | {"NonCoding"}
    hot = [names[i] for i in np.flatnonzero(truth["beta"] > truth["mi"] + 0.9) if names[i] in expect]
    if hot:
        r = res.set_index("name")
        assert r.loc[hot, "bg_mean"].min() > r["bg_mean"].median()
    log = open(out_l).read()
    for needle in ("gene level regression, model: gamma", " Intercept: ", " heritability from genes: ", " Heritability: "):
        assert needle in log, needle


def test_streaming_statistics_equal_the_full_trace_table(snp_csv):
    """K6: the per-gene mean / variance / P accumulated inside the sampler kernel (fp64 sums per coordinate and chain) equal
    the table computed from the full fp64 trace of the same run; the shared parameters kept in fp64 equal the trace's; the
    quantiles from an fp32 copy of the trace agree to fp32 rounding.  BASELINE config-1 shape."""
    import torch
    from baghera_b200 import gene_regression as gr
    from baghera_b200.engine import Engine
    from baghera_b200.sampler import NutsSampler
    d = synth.make_arrays(20_000, 200)
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 4)
    s = NutsSampler(eng, seed=4)
    s.init(synth.jittered_start(d.G + 3, 4, seed=5))
    out = s.sample(400, 150, stream_stats=True)
    tr = out["trace"]                                                     # fp64 [draws, D, C]
    assert torch.equal(out["head_trace"], tr[:, :3, :])
    genes = ["g%d" % i for i in range(d.G)]
    full = gr.gene_posterior_table(tr, genes)
    stream = gr.gene_posterior_table(tr.float(), genes, k6=out["k6"])
    for col in ("bg_mean", "P"):
        assert np.allclose(stream[col], full[col], rtol=1e-12, atol=1e-15), col
    assert np.allclose(stream["bg_var"], full["bg_var"], rtol=1e-9, atol=1e-16)      # E[x^2] - E[x]^2 vs two-pass
    for col in ("bg_median", "bg_5perc", "bg_95perc"):
        assert np.allclose(stream[col], full[col], rtol=1e-6, atol=1e-7), col
    # without any stored positions the sums are still there
    s2 = NutsSampler(eng, seed=4)
    s2.init(synth.jittered_start(d.G + 3, 4, seed=5))
    out2 = s2.sample(400, 150, stream_stats=True, keep_trace=False)
    assert out2["trace"] is None and torch.equal(out2["k6"], out["k6"]) and torch.equal(out2["head_trace"], out["head_trace"])
    eng.close()


def test_devices_do_not_change_the_draws():
    """Philox keys by the global chain index: a block of chains sampled by its own context (as --devices does) draws what
    the same chains draw inside the full batch."""
    from baghera_b200.engine import Engine
    from baghera_b200.sampler import NutsSampler
    d = synth.make_arrays(5000, 20, seed=4)
    q0 = synth.jittered_start(d.G + 3, 4, seed=1)
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 4)
    s = NutsSampler(eng, seed=9)
    s.init(q0)
    a = s.sample(20, 30)["trace"].cpu().numpy()
    eng.close()
    eng2 = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 2)
    s2 = NutsSampler(eng2, seed=9, chain_offset=2)
    s2.init(q0[2:], start_mean=q0.mean(axis=0))
    b = s2.sample(20, 30)["trace"].cpu().numpy()
    eng2.close()
    assert np.allclose(a[:, :, 2:], b, rtol=1e-9, atol=1e-12)
