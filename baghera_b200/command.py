"""``gene_heritability`` / ``gw_heritability`` commands (mirror of ref: baghera_tool/command.py:23-212)."""
from __future__ import annotations

import logging

import numpy as np

from . import gene_regression as gr
from . import genes as g
from . import heritability
from . import logging as log
from . import snps as s


def _prelude(snps, input_snp_filename, sep, output_logger, chromosome, fmt_gene):
    snps.read_table(input_snp_filename, separator=sep)
    snps.generate_stats()
    snps.update_summary()
    output_logger.info(" Sample size " + str(snps.n_patients) + "\n")
    snps.apply_filter_table(s.baghera_filter)
    snps.update_summary()
    output_logger.info(fmt_gene % (str(snps.n_snps), str(snps.n_genes)))
    snps.rename_non_annotated(name="NonCoding")
    if chromosome != "all":
        snps.apply_filter_table(s.cut_single_chrom, **{"chromosome": chromosome})     # works here (Q4)
        output_logger.info("Analysis restricted to chr %s" % str(chromosome))
        snps.update_summary()
        output_logger.info("Analysis. Number of SNPs: %s\n,  Number of genes: %s\n" % (str(snps.n_snps), str(snps.n_genes)))


def gene_heritability(input_snp_filename, output_genes_filename, output_summary_filename, logger_filename,
                      sweeps=1000, burnin=1000, n_chains=4, n_cores=4, N_1kG=1290028, chromosome="all", snp_thr=10,
                      sep=",", model="normal", fix_intercept=False):
    """Bayesian gene-level heritability analysis (ref: command.py:23-130)."""
    output_logger = log.setup_logger("output_logger", logger_filename)
    try:
        log.initialise_log(output_logger, "gene level regression, model: %s" % model, [input_snp_filename],
                           [output_genes_filename, output_summary_filename], sweeps, burnin, chromosome=str(chromosome),
                           other_params_diz={"chains": n_chains, "cores": n_cores, "SNP threshold": snp_thr})
        logging.info("Start Analysis")
        snps = s.Snps()
        logging.info("Reading SNP file: %s,\n\t with %s delimiter" % (input_snp_filename, sep))
        _prelude(snps, input_snp_filename, sep, output_logger, chromosome,
                 "After baghera init filter.\n\t Number of SNPs: %s\n\t Number of genes: %s\n")
        genes = g.Genes()
        genes.initialise_genes(snps.table.copy(), snps_thr=snp_thr)
        output_logger.info("Output gene table initialised:\nNumber of genes: %s\n" % (str(genes.n_genes)))
        snps.set_non_annotated(genes.cut_genes, "NonCoding")

        if model == "gamma":
            result = gr.analyse_gamma(snps, output_summary_filename, output_logger, sweeps, burnin, n_chains, n_cores,
                                      N_1kG, fix_intercept)
        else:
            result = gr.analyse_normal(snps, output_summary_filename, output_logger, sweeps, burnin, n_chains, n_cores,
                                       N_1kG, fix_intercept)

        logging.info("Saving genes table")
        genes.table = genes.table.merge(result, left_index=False, left_on="name", right_on="name")
        k = genes.table.n_snps / float(N_1kG)
        genes.table["h2g"] = genes.table.bg_mean.astype("float") * k
        # the reference sorts P as a string column (Q5); P stays numeric here and sorts numerically
        genes.table = genes.table.sort_values(by=["P", "bg_median"])
        genes.save_table(output_genes_filename)

        non_coding = genes.table[genes.table.name == "NonCoding"]
        h2g_tot = np.sum(genes.table["h2g"].values) - non_coding["h2g"].values
        output_logger.info(" Non coding heritability : " + str(non_coding["h2g"].values) + "\n")
        output_logger.info(" Coding heritability : " + str(h2g_tot) + "\n")
        return genes.table
    finally:
        log.close_logger(output_logger)


def gw_heritability(input_snp_filename, output_summary_filename, logger_filename, sweeps=1000, burnin=1000, n_chains=4,
                    n_cores=4, N_1kG=1290028, chromosome="all", sep=",", model="normal", fix_intercept=False):
    """Genome-wide heritability (ref: command.py:139-212); every `model` value runs gw_normal (:202-211)."""
    output_logger = log.setup_logger("output_logger", logger_filename)
    try:
        log.initialise_log(output_logger, "genome-wide regression, model: %s" % model, [input_snp_filename],
                           [output_summary_filename], sweeps, burnin, chromosome=str(chromosome),
                           other_params_diz={"chains": n_chains, "cores": n_cores})
        logging.info("Start Analysis")
        snps = s.Snps()
        _prelude(snps, input_snp_filename, sep, output_logger, chromosome,
                 "After baghera init filter.\nNumber of SNPs: %s\nNumber of genes: %s\n")
        if model not in ("normal", "gamma"):
            logging.info("Normal model by default")
        intercept, slope = heritability.gw_normal(snps, output_summary_filename, output_logger, sweeps, burnin, n_chains,
                                                  n_cores, N_1kG, fix_intercept)
        logging.info("Analysis complete")
        return intercept, slope
    finally:
        log.close_logger(output_logger)
