"""cet_pick/detectors/detector_factory.py:8-13 restricted to the hot path ('semi') and the semiclass
tile path ('semiclass', detectors/tomo_det_classify.py)."""
from .tomo_det import TomodetDetector
from .tomo_det_classify import TomoClassdetDetector

detector_factory = {
    "semi": TomodetDetector,
    "semiclass": TomoClassdetDetector,
}
