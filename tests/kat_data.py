"""Known-answer inputs taken from the reference's own unit tests (test_monolidar_fusion.cpp)."""
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

# Histogram.FilterPointsMinDistBlob golden vector (test_monolidar_fusion.cpp:306-374)
HIST_DEPTHS = [2.2, 3.5, 4.2, 5.2, 5.2, 6.2, 7.2, 8.2, 8.3, 8.4, 9.2, 10.2, 10.5]
HIST_BIN_WIDTH = 1.0
HIST_MIN_COUNT = 3
HIST_EXPECTED = [8.2, 8.3, 8.4]

_ransac_cloud = None


def ransac_kat_cloud() -> np.ndarray:
    """(18000, 4) float32: the cloud of RansacPlane.CalculateInlersPlane (:376-409), regenerated with
    the same <random> calls (tests/golden/gen_ransac_kat.cpp)."""
    global _ransac_cloud
    if _ransac_cloud is None:
        with tempfile.TemporaryDirectory() as td:
            exe = Path(td) / "gen"
            subprocess.check_call(["/usr/bin/g++", "-O1", "-o", str(exe), str(ROOT / "tests" / "golden" / "gen_ransac_kat.cpp")])
            raw = subprocess.check_output([str(exe)])
        _ransac_cloud = np.frombuffer(raw, np.float32).reshape(-1, 4).copy()
        assert _ransac_cloud.shape == (18000, 4)
    return _ransac_cloud
