#!/usr/bin/env python3
"""ap_fix_badpix: patch an image given a bad pixel mask (reference ``scripts/ap_fix_badpix.py`` :34-93)."""
import argparse
import logging

import astrophotography_b200 as ap


def command_line_opts(argv):
    parser = argparse.ArgumentParser(
        prog="ap_fix_badpix",
        description=("Patches an image given a bad pixel mask, replacing bad pixels with the median"
                     " value of the surrounding good pixels."))
    parser.add_argument("raw_image", metavar="INPUT_IMAGE.FITS", help="Path/name of the input image to patch.")
    parser.add_argument("master_badpix", metavar="BADPIX.FITS",
                        help="Path/name of the master badpixel file (zero at good pixels, non-zero at bad pixels).")
    parser.add_argument("fixed_image", metavar="OUTPUT_IMAGE.FITS", help="Path/name of the output patched image.")
    p_delta = 2
    parser.add_argument("--deltapix", default=p_delta, type=int,
                        help=f"Half-width of the box of donor pixels around a bad pixel. Default: {p_delta} pixels.")
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    return parser.parse_args(argv)


def main(args=None):
    p = command_line_opts(args)
    fixpix = ap.ApFixBadPixels(p.loglevel)
    fixpix.fix_files(p.raw_image, p.master_badpix, p.fixed_image, p.deltapix)
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)
