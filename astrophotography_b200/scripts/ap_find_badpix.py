#!/usr/bin/env python3
"""ap_find_badpix: bad pixel mask from a master dark/bias (reference ``scripts/ap_find_badpix.py`` :35-94)."""
import argparse
import logging

import astrophotography_b200 as ap


def command_line_opts(argv):
    parser = argparse.ArgumentParser(
        prog="ap_find_badpix",
        description="Generates a bad pixel mask given a master dark or master bias FITS file.")
    parser.add_argument("masterdark", metavar="IN_MASTER_DARK.FITS", help="Path/name of the input master dark/bias to use.")
    parser.add_argument("badpixfile", metavar="OUT_BADPIX.FITS", help="Path/name of the output badpix file to generate.")
    p_sigma = 4.0
    parser.add_argument("--sigma", metavar="NSIGMA", default=p_sigma, type=float,
                        help=("Number of standard deviations from the sigma-clipped median beyond which a"
                              f" pixel is marked bad. Default: {p_sigma:.2f}"))
    parser.add_argument("--user_badpix", metavar="USER_BADPIX.YML", default=None,
                        help=("Optional YaML file with user-defined bad rows, columns or rectangles"
                              " (format: etc/user_badpixels.yml of the reference)."))
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    return parser.parse_args(argv)


def main(args=None):
    p = command_line_opts(args)
    mkbadpix = ap.ApFindBadPixels(p.masterdark, p.sigma, p.loglevel)
    if p.user_badpix is not None:
        mkbadpix.add_user_badpix(p.user_badpix)
    mkbadpix.write_mask(p.badpixfile)
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)
