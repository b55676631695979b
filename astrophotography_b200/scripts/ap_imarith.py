#!/usr/bin/env python3
"""ap_imarith: arithmetic on a FITS image with a constant or a second image.

Same command line as the reference's ``scripts/ap_imarith.py`` (:34-87): ``ap_imarith INPUT_IMAGE.FITS OPERATION
VALUE_OR_IMAGE OUTPUT_IMAGE.FITS [--units U] [-l LEVEL]`` with OPERATION one of ADD, SUB, MUL, DIV."""
import argparse
import logging

import astrophotography_b200 as ap


def command_line_opts(argv):
    allowed_ops = ["ADD", "SUB", "MUL", "DIV"]
    parser = argparse.ArgumentParser(
        prog="ap_imarith",
        description=("Perform arithmetic operations on a FITS file, using either a constant value applied to the"
                     " whole primary array or the pixels values in another FITS file of the same shape."))
    parser.add_argument("input_image", metavar="INPUT_IMAGE.FITS", help="Path/name of the input image to perform arithmetic on.")
    parser.add_argument("operation", metavar="OPERATION",
                        help=f"The mathematical operation to perform, a three letter uppercase string. Allowed values are: {allowed_ops}")
    parser.add_argument("value", metavar="VALUE_OR_IMAGE",
                        help="A floating point value OR the path to another fits image of the same shape as the input image.")
    parser.add_argument("output_image", metavar="OUTPUT_IMAGE.FITS", help="Path/name of the output image.")
    parser.add_argument("--units", default=None, help='Output image units, e.g. "adu/s": sets the BUNIT keyword.')
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    return parser.parse_args(argv)


def main(args=None):
    p = command_line_opts(args)
    ap.ApImArith(p.loglevel).process_files(p.input_image, p.operation, p.value, p.output_image, p.units)
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)
