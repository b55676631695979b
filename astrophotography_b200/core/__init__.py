"""Host-side mirrors of the reference's Ap* classes for the FITS-reduction hot path."""
from .ApCalibrate import ApCalibrate
from .ApFindBadPixels import ApFindBadPixels
from .ApFixBadPixels import ApFixBadPixels
from .ApImArith import ApImArith
from .ApMasterCal import ApMasterCal

__all__ = ["ApCalibrate", "ApFindBadPixels", "ApFixBadPixels", "ApImArith", "ApMasterCal"]
