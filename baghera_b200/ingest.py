"""Native reader of the annotated SNP table: ``read_snp_table`` is a drop-in for the
``pd.read_csv`` of ref: baghera_tool/snps.py:32-64 on the schema of ref: docs/createdata.rst:45-47.

The parse runs in libbaghera_b200.so (csrc/bgh_ingest.cpp: mmap + one thread per byte range, typed columns,
sorted label dictionaries); this module only wraps the columns into the DataFrame the host layer expects
(``chr`` as str, ``gene`` as object with NaN for missing labels, numeric columns as float64 / int64).
``rs_id`` is not materialised (nothing downstream reads it).  ``model_inputs`` goes straight from the file to
the kernel inputs (chi2, ld, gene index) without a DataFrame.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class IngestError(ValueError):
    pass


class _Table:
    def __init__(self, path, sep=",", n_threads=0):
        self.L = _lib.lib()
        sep = "\t" if sep == "t" else sep                       # the reference's CLI spells tab as "t"
        if len(sep) != 1:
            raise IngestError("separator must be one character")
        h = ctypes.c_void_p()
        rc = self.L.bgh_csv_read(str(path).encode(), ord(sep), int(n_threads), ctypes.byref(h))
        if rc != 0:
            raise IngestError(self.L.bgh_ingest_last_error().decode())
        self.h = h
        self.n_rows = int(self.L.bgh_table_rows(h))
        self.columns = [self.L.bgh_table_column_name(h, c).decode() for c in range(self.L.bgh_table_n_columns(h))]

    def close(self):
        if getattr(self, "h", None):
            self.L.bgh_table_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def index(self, name):
        try:
            return self.columns.index(name)
        except ValueError:
            raise IngestError("column %r not found (columns: %s)" % (name, ", ".join(self.columns)))

    def f64(self, name):
        c = self.index(name)
        p = self.L.bgh_table_f64(self.h, c)
        if not p:
            raise IngestError("column %r is not numeric" % name)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(self.n_rows,)).copy() \
            if self.n_rows else np.empty(0)

    def labels(self, name):
        """-> (codes int32[n_rows] with -1 = missing, labels list[str] sorted)."""
        c = self.index(name)
        p = self.L.bgh_table_codes(self.h, c)
        if not p:
            raise IngestError("column %r is not a label column" % name)
        codes = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int32)), shape=(self.n_rows,)).copy() \
            if self.n_rows else np.empty(0, np.int32)
        labels = [self.L.bgh_table_label(self.h, c, i).decode() for i in range(self.L.bgh_table_n_labels(self.h, c))]
        return codes, labels

    def kind(self, name):
        return int(self.L.bgh_table_column_kind(self.h, self.index(name)))


def read_snp_table(path, separator=",", n_threads=0):
    """-> pandas.DataFrame with the file's columns (minus ``rs_id``), dtypes as ``pd.read_csv`` + the
    reference's ``chr.astype(str)`` would give them."""
    import pandas as pd

    t = _Table(path, separator, n_threads)
    try:
        cols = {}
        for name in t.columns:
            k = t.kind(name)
            if k == 0:
                v = t.f64(name)
                # pandas infers int64 for columns without a decimal point / NaN (position, sample_size)
                if name in ("position", "sample_size") and v.size and np.all(np.isfinite(v)) and np.all(v == np.rint(v)):
                    v = v.astype(np.int64)
                cols[name] = v
            elif k == 1:
                codes, labels = t.labels(name)
                lab = np.array(labels + [np.nan], dtype=object)      # code -1 -> NaN
                cols[name] = lab[codes]
        df = pd.DataFrame(cols)
        if "chr" in df.columns:
            df["chr"] = df["chr"].astype(str)
        return df
    finally:
        t.close()


def model_inputs(path, separator=",", n_threads=0, non_annotated_name="NonCoding", fix_mhc=False, chromosome="all"):
    """File -> kernel inputs without a DataFrame: the reference's prelude (ref: command.py:64-86:
    stats = z^2, snps.py:117-134 filter with l = 1 + l*(l>0) and maf > 0.01, NaN gene -> NonCoding,
    optional single chromosome) and its gene index (ref: gene_regression.py:56-71).
    -> dict(chi2 float32[M], ld float32[M], gene_id int32[M], genes list[str], Mg int64[G], n_patients)."""
    t = _Table(path, separator, n_threads)
    try:
        z, l, maf = t.f64("z"), t.f64("l"), t.f64("maf")
        n_patients = t.f64("sample_size")[0] if t.n_rows else float("nan")
        codes, labels = t.labels("gene")
        keep = maf > 0.01
        if fix_mhc or chromosome != "all":
            ccodes, clabels = t.labels("chr")
            chrom = np.array(clabels + [""], dtype=object)[ccodes]
            if fix_mhc:
                pos = t.f64("position")
                keep &= ~((chrom == "6") & (pos > 26000000) & (pos < 34000000))
            if chromosome != "all":
                keep &= chrom == str(chromosome)
        l = 1.0 + l * (l > 0)
        stats = z * z
        # NaN -> NonCoding, then rank in the sorted label order of the labels that SURVIVE the filter
        names = list(labels)
        nc = len(names)
        codes = np.where(codes < 0, nc, codes)
        names.append(non_annotated_name)
        codes, stats, l = codes[keep], stats[keep], l[keep]
        present = np.unique(codes)
        order = sorted(set(names[i] for i in present))            # "NonCoding" may coincide with a real label
        rank = {n: i for i, n in enumerate(order)}
        lut = np.full(nc + 1, -1, np.int64)
        for i in present:
            lut[i] = rank[names[i]]
        gid = lut[codes].astype(np.int32)
        return dict(chi2=stats.astype(np.float32), ld=l.astype(np.float32), gene_id=gid, genes=order,
                    Mg=np.bincount(gid, minlength=len(order)).astype(np.int64), n_patients=n_patients,
                    n_rows_file=t.n_rows)
    finally:
        t.close()
