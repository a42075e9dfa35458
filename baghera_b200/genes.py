"""Gene table (mirror of ref: baghera_tool/genes.py:13-78): per-gene descriptive statistics, the
``snp_thr`` filter, the NonCoding row and the results-CSV writer.  Quirk Q2 is kept: ``cut_genes`` is
computed after the table has already been cut to ``n_snps >= thr`` and is therefore always empty."""
import logging

import numpy as np
import pandas as pd

COLUMNS = ["gene", "chrom", "n_snps", "ld_min", "ld_max", "ld_var", "stats_min", "stats_max", "stats_avg"]


class Genes(object):
    def __init__(self):
        self.filename = None
        self.n_genes = None
        self.table = pd.DataFrame()
        self.cut_genes = []

    def update_summary(self):
        pass

    def initialise_genes(self, snps_table, snps_thr, non_annotated_name="NonCoding"):
        """ref: genes.py:39-75."""
        logging.info("Initialisation of genes table")
        vTot = np.var(snps_table["l"])
        coding = snps_table[snps_table["gene"] != non_annotated_name]
        g = coding.groupby("gene", sort=True)
        # vectorised equivalent of groupby(...).apply(get_stats): all-float columns like the reference's result
        tab = pd.DataFrame({
            "chrom": g["chr"].first().astype(int).astype(float),
            "n_snps": g.size().astype(float),
            "ld_min": g["l"].min(),
            "ld_max": g["l"].max(),
            "ld_var": g["l"].var(ddof=0).fillna(0.0),
            "stats_min": g["stats"].min(),
            "stats_max": g["stats"].max(),
            "stats_avg": g["stats"].mean(),
        }).reset_index()
        tab = tab[tab["n_snps"] >= snps_thr]
        self.cut_genes = tab[tab["n_snps"] < snps_thr].gene.values.tolist()      # always [] (Q2)
        self.gene_names = tab["gene"].values.tolist()
        non_coding = snps_table[~(snps_table["gene"].isin(self.gene_names))]
        row = {
            "gene": non_annotated_name,
            "chrom": np.nan,
            "n_snps": float(len(non_coding)),
            "ld_var": np.var(non_coding["l"].values),
            "ld_min": np.min(non_coding["l"].values),
            "ld_max": np.max(non_coding["l"].values),
            "stats_avg": np.average(non_coding["stats"]),
            "stats_min": np.min(non_coding["stats"]),
            "stats_max": np.max(non_coding["stats"]),
        }
        tab = pd.concat([tab, pd.DataFrame([row])[COLUMNS]], ignore_index=True)
        tab["ld_var"] = tab["ld_var"] / vTot
        tab = tab.rename(columns={"gene": "name"})
        self.table = tab
        self.n_genes = len(self.table)

    def save_table(self, output_filename, separator=","):
        self.table.to_csv(output_filename, sep=",", index=False)
