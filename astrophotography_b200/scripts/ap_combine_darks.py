#!/usr/bin/env python3
"""ap_combine_darks: master dark / bias / flat from a directory of raw calibration frames.

Same command line as the reference's ``scripts/ap_combine_darks.py`` (:48-98):
``ap_combine_darks RAW_CAL_DIR MASTER_CAL_FILENAME [--exclude PAT] [--telescop NAME]
[--temptol DEG] [-l LEVEL]``; the defaults reproduce its combine settings (:394-399).
Additive options select the other combine modes of the GPU stack reducer.
"""
import argparse
import logging

from astrophotography_b200 import ApMasterCal


def command_line_opts(argv):
    parser = argparse.ArgumentParser(
        prog="ap_combine_darks",
        description=("Generates a master dark or master bias file from all calibration FITS files in a"
                     " given directory."))
    parser.add_argument("rawcaldir", metavar="RAW_CAL_DIR",
                        help="The directory in which the raw calibration files to be combined are to be found.")
    parser.add_argument("master_filename", metavar="MASTER_CAL_FILENAME", help="Output file name for master calibration file.")
    p_temptol, p_telescop, p_exclude = 0.5, "UNKNOWN", "master*"
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    parser.add_argument("--exclude", dest="exclude_pattern", default=p_exclude, metavar="FILE_PATTERN",
                        help=f'Unix-style pattern of files in the directory to skip. Default: "{p_exclude}"')
    parser.add_argument("--telescop", default=p_telescop, metavar="TELESCOPE_NAME",
                        help=f"TELESCOP value to write when the inputs have none. Default: {p_telescop}")
    parser.add_argument("--temptol", default=p_temptol, type=float, metavar="DEGREES_C",
                        help=f"Allowed |CCD-TEMP - SET-TEMP|. Default: {p_temptol} C.")
    # additive (not in the reference): the other settings of the stack reducer
    parser.add_argument("--method", default="average", choices=["average", "median", "min", "max"])
    parser.add_argument("--no_sigma_clip", action="store_true", help="Disable clipping.")
    parser.add_argument("--kappa_low", type=float, default=5.0)
    parser.add_argument("--kappa_high", type=float, default=5.0)
    parser.add_argument("--maxiters", type=int, default=1, help="Clip passes (ccdproc: 1; astropy sigma_clip: 5).")
    parser.add_argument("--cenfunc", default="median", choices=["median", "mean"])
    parser.add_argument("--devfunc", default="mad_std", choices=["mad_std", "std"])
    parser.add_argument("--out_dtype", default="float64", choices=["float64", "float32"])
    parser.add_argument("--gpus", type=int, default=1,
                        help="Shard the rows of the stack over this many GPUs of the box (one process each). Default: 1")
    return parser.parse_args(argv)


def main(args=None):
    p = command_line_opts(args)
    logger = logging.getLogger(__name__)
    try:
        mkcal = ApMasterCal(p.rawcaldir, p.exclude_pattern, p.telescop, p.temptol, p.loglevel,
                            method=p.method, sigma_clip=not p.no_sigma_clip,
                            sigma_clip_low_thresh=p.kappa_low, sigma_clip_high_thresh=p.kappa_high,
                            maxiters=p.maxiters, cenfunc=p.cenfunc, devfunc=p.devfunc, out_dtype=p.out_dtype)
        mkcal.make_master(p.master_filename, gpus=p.gpus)
    except RuntimeError as rte:
        logger.error(f"Shutting down due to exception raised by ApMasterCal: {rte}")
        return 1
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)
