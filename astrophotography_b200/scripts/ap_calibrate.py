#!/usr/bin/env python3
"""ap_calibrate: bias / dark / flat calibration (+ bad-pixel repair) of one raw frame.

Same command line as the reference's ``scripts/ap_calibrate.py`` (:40-122):
``ap_calibrate RAW BIAS DARK OUT [--master_flat F] [--master_badpix M] [--normflat N]
[--deltapix 2] [--fixcosmic] [--dark_still_biased] [-l LEVEL]``.
"""
import argparse
import logging

import astrophotography_b200 as ap


def command_line_opts(argv):
    parser = argparse.ArgumentParser(
        prog="ap_calibrate",
        description=("Performs calibration of raw astronomical images by applying bias and dark frame"
                     " subtraction, along with optional (but recommended) flat fielding, bad pixel"
                     " correction, and cosmic ray removal."))
    parser.add_argument("raw_image", metavar="INPUT_IMAGE.FITS", help="Path/name of the raw (uncalibrated) input image.")
    parser.add_argument("master_bias", metavar="MBIAS.FITS", help="Path/name of the master bias file.")
    parser.add_argument("master_dark", metavar="MDARK.FITS", help="Path/name of the master dark file.")
    parser.add_argument("calibrated_image", metavar="CALIBRATED_IMAGE.FITS", help="Path/name of the output calibrated image.")
    p_delta = 2
    parser.add_argument("--master_flat", metavar="MFLAT.FITS", default=None,
                        help="Path/name of the master flat file. If given, flat fielding is applied.")
    parser.add_argument("--master_badpix", metavar="BADPIX.FITS", default=None,
                        help=("Path/name of the master badpixel file (zero at good pixels, non-zero at bad"
                              " pixels, e.g. from ap_find_badpix). If given, bad pixels are repaired."))
    parser.add_argument("--normflat", metavar="NORMALIZED_FLAT.FITS", default=None,
                        help="Optional output of the normalized flat used.")
    parser.add_argument("--deltapix", default=p_delta, type=int,
                        help=("Half-width of the box around a bad pixel from which the median of the good"
                              f" pixels is taken (1: 8 neighbours, 2: 24). Default: {p_delta} pixels."))
    parser.add_argument("--fixcosmic", default=False, action="store_true",
                        help="If specified, cosmic ray removal will be performed (needs ccdproc).")
    parser.add_argument("--dark_still_biased", default=False, action="store_true",
                        help=("The master dark has NOT had the bias subtracted: subtract it before scaling"
                              " the dark by the exposure time ratio."))
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    return parser.parse_args(argv)


def main(args=None):
    p = command_line_opts(args)
    calibrator = ap.ApCalibrate(p.master_bias, p.master_dark, p.master_flat, p.master_badpix,
                                p.loglevel, p.dark_still_biased)
    calibrator.calibrate(p.raw_image, p.calibrated_image, p.deltapix, p.normflat, p.fixcosmic)
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:                                 # the reference uses a bare except here, which also logs on --help
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)
