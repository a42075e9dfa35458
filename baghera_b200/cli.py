#!/usr/bin/env python
"""``baghera-tool`` command line (ref: baghera_tool/cli.py:14-25; flags as argh derives them from
ref: baghera_tool/command.py:23-38, documented at ref: docs/analysis.rst:19-67).  argh is replaced by
argparse; ``--device`` / ``--seed`` / ``--trace-dtype`` are additions, nothing is renamed."""
from __future__ import annotations

import argparse
import sys

from . import command as cmd
from . import gene_regression as gr
from .logging import setup_console_logger


def _common(p, with_snp_thr):
    p.add_argument("--sweeps", type=int, default=1000, help="number of samples for each chain")
    p.add_argument("-b", "--burnin", type=int, default=1000, help="number of burnin samples")
    p.add_argument("--n-chains", type=int, default=4, help="number of chains of the sampler")
    p.add_argument("--n-cores", type=int, default=4, help="number of parallel cores to use (advisory on the GPU path)")
    p.add_argument("-N", "--N-1kG", dest="N_1kG", type=int, default=1290028,
                   help="number of SNPs on which the LD-score is calculated")
    p.add_argument("-c", "--chromosome", default="all", help="chromosome on which the analysis is run")
    if with_snp_thr:
        p.add_argument("--snp-thr", type=int, default=10, help="threshold for the minimum number of SNPs in a gene")
    p.add_argument("--sep", default=",", help="separator for the input files, use t for tab separated (not \\t)")
    p.add_argument("-m", "--model", default="normal", help="regression model, one between normal/gamma")
    p.add_argument("-f", "--fix-intercept", action="store_true", default=False)
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    p.add_argument("--devices", default=None,
                   help="comma-separated CUDA ordinals: the chains run as independent replicas over these GPUs "
                        "(one host thread, engine and sampler kernel per device); default: --device only")
    p.add_argument("--seed", type=int, default=20211, help="Philox seed (the reference is unseeded)")
    p.add_argument("--trace-dtype", default="float64", choices=["float64", "float32"])
    p.add_argument("--stream-stats", default="auto", choices=["auto", "on", "off"],
                   help="accumulate the per-gene mean / variance / P while sampling and keep only an fp32 trace for the quantiles "
                        "(auto: when the fp64 trace would exceed 8 GiB)")


def build_parser():
    ap = argparse.ArgumentParser(prog="baghera-tool", description="BAGHERA gene-level heritability analysis (B200 path)")
    sub = ap.add_subparsers(dest="command")
    g = sub.add_parser("gene-heritability", help="Performs bayesian gene-level heritability regression analysis.")
    g.add_argument("input_snp_filename", metavar="input-snp-filename", help="Data Input, use the SNPs file from dataParse")
    g.add_argument("output_genes_filename", metavar="output-genes-filename", help="output file for gene-level results, use .csv")
    g.add_argument("output_summary_filename", metavar="output-summary-filename",
                   help="output file for the genomewide results summary, use .csv")
    g.add_argument("logger_filename", metavar="logger-filename", help="file for the logger, use a txt")
    _common(g, True)
    w = sub.add_parser("gw-heritability", help="Computes the genome-wide estimate heritability using Bayesian regression.")
    w.add_argument("input_snp_filename", metavar="input-snp-filename")
    w.add_argument("output_summary_filename", metavar="output-summary-filename")
    w.add_argument("logger_filename", metavar="logger-filename")
    _common(w, False)
    for name in ("create-files", "generate-snp-file"):
        sub.add_parser(name, help="offline data preparation: not part of the B200 hot path, use the reference tool")
    return ap


def main(argv=None):
    setup_console_logger()
    ap = build_parser()
    a = ap.parse_args(argv)
    if a.command is None:
        ap.print_help()
        return 1
    if a.command in ("create-files", "generate-snp-file"):
        print("%s is offline data preparation (ref: baghera_tool/preprocess.py) and is out of scope of baghera_b200; "
              "run it with the reference tool." % a.command, file=sys.stderr)
        return 2
    devices = [int(x) for x in a.devices.split(",")] if a.devices else None
    n_dev = len(devices) if devices else 1
    if (a.n_chains + n_dev - 1) // n_dev > 128:
        # the persistent sampler kernel runs one CTA per chain and at most 128 chains per context (bgh_nuts_run)
        print("--n-chains %d: at most 128 chains per device (use --devices to spread the chains over more GPUs)" % a.n_chains, file=sys.stderr)
        return 2
    gr.OPTIONS.update(device=devices[0] if devices else a.device, seed=a.seed, trace_dtype=a.trace_dtype, devices=devices,
                      stream_stats={"auto": None, "on": True, "off": False}[a.stream_stats])
    chrom = a.chromosome if a.chromosome == "all" else a.chromosome
    if a.command == "gene-heritability":
        cmd.gene_heritability(a.input_snp_filename, a.output_genes_filename, a.output_summary_filename, a.logger_filename,
                              sweeps=a.sweeps, burnin=a.burnin, n_chains=a.n_chains, n_cores=a.n_cores, N_1kG=a.N_1kG,
                              chromosome=chrom, snp_thr=a.snp_thr, sep=a.sep, model=a.model, fix_intercept=a.fix_intercept)
    else:
        cmd.gw_heritability(a.input_snp_filename, a.output_summary_filename, a.logger_filename, sweeps=a.sweeps,
                            burnin=a.burnin, n_chains=a.n_chains, n_cores=a.n_cores, N_1kG=a.N_1kG, chromosome=chrom,
                            sep=a.sep, model=a.model, fix_intercept=a.fix_intercept)
    return 0


if __name__ == "__main__":
    sys.exit(main())
