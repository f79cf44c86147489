import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# GPU tests: fill the engine's device lambda buffer with NaN before every (partial) upload, so a band render
# that read a row it did not upload could not pass (read once by the library, at its first host-API call).
os.environ.setdefault("FG_B200_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
