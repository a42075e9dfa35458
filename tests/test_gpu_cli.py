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
