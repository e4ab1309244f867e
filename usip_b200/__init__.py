"""usip_b200 -- B200-native (sm_100a) drop-in for the USIP detector / descriptor hot path.

Layout mirrors the reference's own import names so that a caller can switch by putting this package
directory first on sys.path:
    import index_max, ball_query                      -> usip_b200/index_max.py, usip_b200/ball_query.py
    from models import networks, losses, layers       -> usip_b200/models/
    from models.keypoint_detector import ModelDetector
    from util import som                              -> usip_b200/util/som.py
All arithmetic is done by libusip_b200.so (usip_b200/csrc, C ABI in include/usip_b200.h).
"""
__version__ = "0.1.0"
