"""astrophotography_b200: the FITS-reduction hot path of DaveStrickland/AstroPhotography
(master-frame combine, science-frame calibration, bad-pixel repair) on B200 GPUs.

``import astrophotography_b200 as ap`` exposes the same class names the reference's
``import AstroPhotography as ap`` does for this path (``ap.ApCalibrate``,
``ap.ApFixBadPixels``, ``ap.ApFindBadPixels``, ``ap.ApImArith``) plus ``ApMasterCal`` (which the
reference keeps inside ``scripts/ap_combine_darks.py``).  Importing the package
does not need a GPU; constructing a class or calling a kernel does, and fails
loudly without one -- there is no CPU fallback.
"""
from .version import __version__
from .core import ApCalibrate, ApFindBadPixels, ApFixBadPixels, ApImArith, ApMasterCal

__all__ = ["__version__", "ApCalibrate", "ApFindBadPixels", "ApFixBadPixels", "ApImArith", "ApMasterCal"]
