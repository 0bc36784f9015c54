"""cet_pick/detectors/detector_factory.py:8-13 restricted to the hot path ('semi')."""
from .tomo_det import TomodetDetector

detector_factory = {
    "semi": TomodetDetector,
}
