"""CPU tests of the native SNP-table reader (csrc/bgh_ingest.cpp) against the pandas path the reference uses
(ref: baghera_tool/snps.py:32-64, :117-134; gene index ref: gene_regression.py:56-71)."""
import numpy as np
import pandas as pd
import pytest

from baghera_b200 import ingest, snps as s, synth
from baghera_b200.gene_regression import build_gene_index


@pytest.fixture(scope="module")
def csv_path(tmp_path_factory, built_lib):
    p = tmp_path_factory.mktemp("ingest") / "snps.csv"
    df, _ = synth.make_snp_table(30_000, 120, seed=3)
    df.to_csv(p, index=False)
    return str(p)


def _pandas_table(path, sep=","):
    t = pd.read_csv(path, sep=sep, float_precision="round_trip")
    t["chr"] = t["chr"].astype(str)
    return t


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_matches_pandas(csv_path, threads):
    ref = _pandas_table(csv_path)
    got = ingest.read_snp_table(csv_path, n_threads=threads)
    assert list(got.columns) == [c for c in ref.columns if c != "rs_id"]
    assert len(got) == len(ref)
    for c in got.columns:
        if c in ("gene", "chr"):
            a, b = got[c].astype(object).values, ref[c].astype(object).values
            na, nb = pd.isna(a), pd.isna(b)
            assert np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb]), c
        else:
            assert got[c].dtype == ref[c].dtype, (c, got[c].dtype, ref[c].dtype)
            assert np.array_equal(got[c].values, ref[c].values, equal_nan=True), c


def test_model_inputs_match_reference_prelude(csv_path):
    """file -> (chi2, ld, cat) must equal read_table -> generate_stats -> baghera_filter -> rename -> gene index."""
    sn = s.Snps()
    sn.table = _pandas_table(csv_path)
    sn.generate_stats()
    sn.update_summary()
    sn.apply_filter_table(s.baghera_filter)
    sn.rename_non_annotated(name="NonCoding")
    cat, Mg, genes = build_gene_index(sn.table)
    got = ingest.model_inputs(csv_path)
    assert got["genes"] == genes and np.array_equal(got["Mg"], Mg)
    assert np.array_equal(got["gene_id"], cat.astype(np.int32))
    assert np.array_equal(got["chi2"], sn.table["stats"].values.astype(np.float32))
    assert np.array_equal(got["ld"], sn.table["l"].values.astype(np.float32))
    assert got["n_patients"] == sn.n_patients
    # single chromosome
    one = ingest.model_inputs(csv_path, chromosome="2")
    ref2 = s.cut_single_chrom(sn.table, chromosome="2")
    assert len(one["chi2"]) == len(ref2) and set(one["genes"]) == set(ref2["gene"])


def test_edge_cases(tmp_path, built_lib):
    p = tmp_path / "e.csv"
    p.write_text("chr,rs_id,position,cm,maf,l,gene,sample_size,z\r\n"
                 "1,rs1,100,0.5,0.2,-3.5,,1000,1.5\r\n"            # missing gene, negative ld, CRLF
                 "X,rs2,200,0.6,0.005,2.0,\"G B\",1000,-2\r\n"        # quoted label, maf below the filter
                 "\r\n"                                               # blank line
                 "6,rs3,30000000,0.7,0.3,NA,GA,1000,1e-3")            # NA number, no trailing newline
    t = ingest.read_snp_table(str(p))
    assert len(t) == 3 and t["chr"].tolist() == ["1", "X", "6"]
    assert pd.isna(t["gene"][0]) and t["gene"][1] == "G B" and np.isnan(t["l"][2]) and t["z"][2] == 1e-3
    assert t["position"].dtype == np.int64
    m = ingest.model_inputs(str(p))
    assert m["genes"] == ["GA", "NonCoding"] and m["gene_id"].tolist() == [1, 0] and m["ld"][0] == 1.0
    assert len(ingest.model_inputs(str(p), fix_mhc=True)["chi2"]) == 1
    extra = tmp_path / "x.csv"                                           # columns outside the schema: typed by their first value
    extra.write_text(",chr,maf,l,gene,sample_size,z,allele,score\n0,1,0.2,1.0,A,10,2,T,0.5\n1,2,0.3,2.0,B,10,3,C,NA\n")
    x = ingest.read_snp_table(str(extra))
    assert x["allele"].tolist() == ["T", "C"] and x["score"][0] == 0.5 and np.isnan(x["score"][1]) and x[""].tolist() == [0.0, 1.0]
    tab = tmp_path / "t.tsv"
    tab.write_text("chr\tmaf\tl\tgene\tsample_size\tz\n1\t0.2\t1.0\tA\t10\t2\n")
    assert ingest.read_snp_table(str(tab), separator="t")["z"].tolist() == [2.0]


def test_errors(tmp_path, built_lib):
    with pytest.raises(ingest.IngestError, match="cannot open"):
        ingest.read_snp_table(str(tmp_path / "missing.csv"))
    p = tmp_path / "bad.csv"
    p.write_text("chr,maf,l,gene,sample_size,z\n1,0.2,1.0,A,10\n")
    with pytest.raises(ingest.IngestError, match="fields"):
        ingest.read_snp_table(str(p))
    p.write_text("chr,maf,l,gene,sample_size,z\n1,0.2,abc,A,10,1\n")
    with pytest.raises(ingest.IngestError, match="cannot parse 'abc'"):
        ingest.read_snp_table(str(p))
    p.write_text("chr,maf,l,sample_size,z\n1,0.2,1,10,1\n")
    with pytest.raises(ingest.IngestError, match="'gene' not found"):
        ingest.model_inputs(str(p))


def test_error_reports_the_file_line(tmp_path, built_lib):
    """A bad row deep inside a file that several threads parse is reported with its FILE line number (header = line 1),
    not with a row index relative to a thread's byte range."""
    rows = ["chr,maf,l,gene,sample_size,z"] + ["1,0.2,%d.5,G%d,10,1.0" % (i % 7 + 1, i % 50) for i in range(20_000)]
    bad_line = 15_432                                  # 1-based file line
    rows[bad_line - 1] = "1,0.2,oops,G1,10,1.0"
    p = tmp_path / "deep.csv"
    p.write_text("\n".join(rows) + "\n")
    for threads in (1, 4):
        with pytest.raises(ingest.IngestError, match=r"line %d: column 'l': cannot parse 'oops'" % bad_line):
            ingest.read_snp_table(str(p), n_threads=threads)
    rows[bad_line - 1] = "1,0.2,1.5,G1,10"
    p.write_text("\n".join(rows) + "\n")
    with pytest.raises(ingest.IngestError, match=r"line %d: has 5 fields" % bad_line):
        ingest.read_snp_table(str(p), n_threads=4)
