"""TEST INFRASTRUCTURE - re-export of the synthetic weight / input generator (it lives in the package because bench.py uses it too)."""
from multiyolov5_b200.synth import *  # noqa: F401,F403
from multiyolov5_b200.synth import GOLDEN_DIR, CFG_DIR  # noqa: F401
