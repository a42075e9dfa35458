"""SNP table semantics that define the kernel inputs (mirror of ref: baghera_tool/snps.py:17-140).

Same class / method names and behaviour as the reference, re-implemented on current pandas; the
reference's quirks that change what the hot path sees are kept and documented:
Q1 the MHC filter compares a string column with the int 6 and is therefore a no-op; Q7 the LD
transform ``l = 1 + l*(l>0)`` runs inside the filter.  Q3 (``sample_size[0]`` is a label lookup
that raises once row 0 has been filtered out) is NOT reproduced: the first remaining row is used.
"""
import logging

import numpy as np
import pandas as pd


def square(a):
    return a ** 2


class Snps(object):
    def __init__(self):
        self.table = pd.DataFrame()
        self.n_patients = None
        self.n_snps = None
        self.n_genes = None
        self.min_maf = None
        self.max_maf = None
        self.min_ld = None
        self.max_ld = None
        self.avg_stats = None

    def read_table(self, input_snp_filename, separator=",", native=True):
        """ref: snps.py:32-64 — ',' / 't' (tab) / any other separator; ``chr`` is cast to str.
        ``native=True`` (default) parses with the library's multi-threaded reader (csrc/bgh_ingest.cpp via
        ``ingest.read_snp_table``: same table minus the unused ``rs_id`` column, ~3x faster than pandas at
        1.1M rows); multi-character separators and ``native=False`` use ``pd.read_csv`` like the reference."""
        sep = "\t" if separator == "t" else separator
        if native and len(sep) == 1:
            # the native reader covers the documented schema without quoting; anything else goes through pandas, like the
            # reference (quoted fields with doubled quotes, extra columns whose dtype pandas would infer differently)
            with open(input_snp_filename, "rb") as f:
                head = f.read(1 << 16)
            cols = head.split(b"\n", 1)[0].decode("utf-8", "replace").strip().split(sep)
            schema = {"chr", "rs_id", "position", "cm", "maf", "l", "gene", "sample_size", "z"}
            if b'"' in head or not set(c.strip() for c in cols) <= schema:
                native = False
        if native and len(sep) == 1:
            from . import ingest
            open(input_snp_filename).close()        # a missing file raises here, as in the reference
            try:
                self.table = ingest.read_snp_table(input_snp_filename, separator=sep)
            except ingest.IngestError:
                logging.exception("Wrong format of the input file")
            return
        with open(input_snp_filename) as f:
            try:
                self.table = pd.read_csv(f, sep=sep)
                self.table["chr"] = self.table["chr"].astype(str)
            except ValueError:
                logging.exception("Wrong format of the input file")

    def generate_stats(self, from_column="z", apply=square):
        """ref: snps.py:66-78 — stats = z**2."""
        logging.warning(self.table.columns)
        if "stats" in self.table.columns:
            logging.info("stats column already exists and is going to be replaced")
        self.table["stats"] = apply(self.table[from_column])

    def update_summary(self):
        """ref: snps.py:80-92."""
        self.n_patients = self.table["sample_size"].iloc[0]     # Q3: the reference uses the label 0
        self.n_snps = len(self.table)
        self.min_maf = np.min(self.table["maf"].values)
        self.max_maf = np.max(self.table["maf"].values)
        self.min_ld = np.min(self.table["l"].values)
        self.max_ld = np.max(self.table["l"].values)
        self.avg_stats = np.average(self.table["stats"].values)
        self.n_genes = len(set(self.table["gene"].values.tolist()))

    def apply_filter_table(self, fltr, **args):
        self.table = fltr(self.table, **args)

    def rename_non_annotated(self, name="NonCoding"):
        """ref: snps.py:98-102 — SNPs without a gene go to the dummy gene."""
        self.table["gene"] = self.table["gene"].astype(object).where(self.table["gene"].notna(), name)

    def set_non_annotated(self, names, non_annotated_name):
        """ref: snps.py:104-109."""
        if len(names):
            self.table["gene"] = self.table["gene"].replace(to_replace=list(names), value=non_annotated_name)


def baghera_filter(table, fix_mhc=False):
    """ref: snps.py:117-134.  ``fix_mhc=True`` applies the chr6 26-34 Mb exclusion the reference
    intended (opt-in; by default the comparison ``chr != 6`` on a str column keeps every SNP, Q1)."""
    table = table.copy()
    table["l"] = 1 + table["l"] * (table["l"] > 0)      # ld-score [1,+inf)
    logging.info("table length: %d" % len(table))
    table = table[table["maf"] > 0.01]
    logging.info("table length maf > 0.01: %d" % len(table))
    chr6 = (table["chr"] == "6") if fix_mhc else pd.Series(False, index=table.index)
    table = table[(~chr6) | ((table.position >= 34000000) | (table.position <= 26000000))]
    logging.info("table no HPC chrom 6: %d" % len(table))
    return table


def cut_single_chrom(table, chromosome=1):
    """ref: snps.py:137-140 (the reference calls it as a method and crashes, Q4; works here)."""
    return table[table["chr"] == str(chromosome)]
