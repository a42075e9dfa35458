"""Console / results-log format of the reference (ref: baghera_tool/logging.py:5-53).

The file logger IS one of the three user-facing output artefacts, so record formats and
message texts are kept as the reference emits them (including its spelling).
"""
import datetime
import logging


def setup_console_logger():
    logging.basicConfig(level=logging.INFO, format=" [%(levelname)s]: %(message)s")


def setup_logger(name, log_file, level=logging.INFO):
    handler = logging.FileHandler(log_file)
    handler.setFormatter(logging.Formatter("[%(levelname)s]: %(message)s"))
    logger = logging.getLogger(name)
    logger.setLevel(level)
    logger.addHandler(handler)
    return logger


def close_logger(logger):
    """Flush and detach the file handlers (the reference leaves them open until interpreter exit)."""
    for h in list(logger.handlers):
        h.flush()
        h.close()
        logger.removeHandler(h)


def initialise_log(logger_object, analysis, input_filenames, output_filenames, sweeps, tune, chromosome="all",
                   other_params_diz=None):
    now = datetime.datetime.now()
    logger_object.info("BAGHERA resutls\n " "--------------------------------\n" + "Current date & time "
                       + now.strftime("%Y-%m-%d %H:%M \n"))
    logger_object.info("Analysis: %s \n" % str(analysis))
    logger_object.info(" Input files: %s \n" % (",".join(input_filenames)))
    logger_object.info(" Output files: %s \n" % (",".join(output_filenames)))
    logger_object.info("- sweeps: %d, \n- burnin: %d, \n- chr: %s\nOther params:\n" % (sweeps, tune, str(chromosome)))
    for k, val in (other_params_diz or {}).items():
        logger_object.info("\t- %s: %s\n" % (str(k), str(val)))
