"""CPU tests of the host-side mirror of the reference's table semantics and CLI surface."""
import os

import numpy as np
import pandas as pd
import pytest

from baghera_b200 import genes as g
from baghera_b200 import snps as s
from baghera_b200 import synth
from baghera_b200.cli import build_parser
from baghera_b200.gene_regression import build_gene_index


@pytest.fixture(scope="module")
def snp_csv(tmp_path_factory):
    df, truth = synth.make_snp_table(6000, 40, seed=3)
    p = tmp_path_factory.mktemp("data") / "snps.csv"
    df.to_csv(p, index=False)
    return str(p), df


def _load(path):
    sn = s.Snps()
    sn.read_table(path)
    sn.generate_stats()
    sn.update_summary()
    return sn


def test_read_stats_filter(snp_csv):
    path, raw = snp_csv
    sn = _load(path)
    assert sn.table["chr"].map(type).eq(str).all()                       # chr cast to str (snps.py:43)
    assert np.allclose(sn.table["stats"], raw["z"] ** 2)                 # stats = z**2 (snps.py:78)
    assert sn.n_patients == 361194
    n0 = len(sn.table)
    sn.apply_filter_table(s.baghera_filter)
    sn.update_summary()
    kept = raw[raw["maf"] > 0.01]
    assert len(sn.table) == len(kept) < n0                               # MAF filter; MHC filter is a no-op (Q1)
    assert sn.min_ld >= 1.0                                              # l = 1 + l*(l>0)  (Q7)
    neg = kept["l"] <= 0
    assert np.all(sn.table.loc[neg[neg].index, "l"] == 1.0)
    # opt-in fix of the MHC filter removes chr6 26-34Mb
    t2 = s.baghera_filter(_load(path).table, fix_mhc=True)
    mhc = (kept["chr"].astype(str) == "6") & (kept.position > 26000000) & (kept.position < 34000000)
    assert len(t2) == len(kept) - int(mhc.sum())


def test_gene_table_and_index(snp_csv):
    path, raw = snp_csv
    sn = _load(path)
    sn.apply_filter_table(s.baghera_filter)
    sn.update_summary()
    sn.rename_non_annotated("NonCoding")
    assert sn.table["gene"].isna().sum() == 0
    ge = g.Genes()
    ge.initialise_genes(sn.table.copy(), snps_thr=10)
    assert ge.cut_genes == []                                            # Q2
    tab = ge.table
    assert list(tab.columns) == ["name", "chrom", "n_snps", "ld_min", "ld_max", "ld_var", "stats_min", "stats_max", "stats_avg"]
    assert tab["name"].iloc[-1] == "NonCoding" and np.isnan(tab["chrom"].iloc[-1])
    counts = sn.table.groupby("gene").size()
    small = counts[(counts < 10) & (counts.index != "NonCoding")]
    # genes under the threshold are not in the table but their SNPs are counted in the NonCoding row (Q2)
    assert not set(small.index) & set(tab["name"])
    assert tab["n_snps"].iloc[-1] == counts["NonCoding"] + small.sum()
    one = tab[tab["name"] == tab["name"].iloc[0]].iloc[0]
    grp = sn.table[sn.table["gene"] == one["name"]]
    assert one["n_snps"] == len(grp) and np.isclose(one["ld_var"], np.var(grp["l"].values) / np.var(sn.table["l"]))
    assert np.isclose(one["stats_avg"], grp["stats"].mean()) and one["chrom"] == float(grp["chr"].iloc[0])
    # gene index = rank of the label in sorted order; small genes stay separate genes of the regression
    cat, Mg, names = build_gene_index(sn.table.reset_index(drop=True))
    assert names == sorted(names) and len(names) == sn.n_genes and Mg.sum() == len(sn.table)
    assert np.array_equal(Mg, counts.loc[names].values)
    sn.set_non_annotated(ge.cut_genes, "NonCoding")                      # no-op
    assert len(set(sn.table["gene"])) == sn.n_genes


def test_cut_single_chrom(snp_csv):
    path, _ = snp_csv
    sn = _load(path)
    sn.apply_filter_table(s.cut_single_chrom, chromosome=3)
    assert len(sn.table) > 0 and set(sn.table["chr"]) == {"3"}


def test_cli_surface():
    ap = build_parser()
    a = ap.parse_args(["gene-heritability", "in.csv", "genes.csv", "summary.csv", "log.txt"])
    assert (a.sweeps, a.burnin, a.n_chains, a.n_cores, a.N_1kG, a.chromosome, a.snp_thr, a.sep, a.model,
            a.fix_intercept) == (1000, 1000, 4, 4, 1290028, "all", 10, ",", "normal", False)   # command.py:28-37
    a = ap.parse_args(["gene-heritability", "i", "g", "s", "l", "--sweeps", "10000", "-b", "2500", "--n-chains", "4",
                       "--n-cores", "4", "-m", "normal", "-N", "100", "-c", "2", "-f"])
    assert a.sweeps == 10000 and a.burnin == 2500 and a.N_1kG == 100 and a.chromosome == "2" and a.fix_intercept
    a = ap.parse_args(["gw-heritability", "i", "s", "l", "--sweeps", "5"])
    assert a.command == "gw-heritability" and not hasattr(a, "snp_thr")


def test_results_log_format(tmp_path):
    from baghera_b200 import logging as blog
    p = tmp_path / "log.txt"
    lg = blog.setup_logger("test_logger_fmt", str(p))
    blog.initialise_log(lg, "gene level regression, model: normal", ["a.csv"], ["g.csv", "s.csv"], 10, 5,
                        other_params_diz={"chains": 4})
    blog.close_logger(lg)
    txt = p.read_text()
    assert txt.startswith("[INFO]: BAGHERA resutls\n --------------------------------\nCurrent date & time ")
    assert "[INFO]: Analysis: gene level regression, model: normal \n" in txt
    assert "[INFO]:  Output files: g.csv,s.csv \n" in txt and "[INFO]: \t- chains: 4\n" in txt
    assert "- sweeps: 10, \n- burnin: 5, \n- chr: all\nOther params:\n" in txt
