"""``ap_*`` command-line entry points for the FITS-reduction hot path (same
program names, positional arguments and options as the reference's
``AstroPhotography/scripts/ap_*.py``)."""
