import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
DISK_FILE = os.path.join(GOLDEN_DIR, "disk_aniso_brushed_aluminium_1_rgb.npz")
SPH_FILE = os.path.join(GOLDEN_DIR, "spherical_aniso_brushed_aluminium_1_rgb.npz")
BSDF_FILE = os.path.join(GOLDEN_DIR, "bsdf_0.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden_ids():
    return [os.path.basename(f)[:-4] for f in GOLDEN_FILES]


@pytest.fixture(scope="session")
def built_lib():
    """Build libbsdfdiff.so if needed (nvcc cross-compiles on CPU) and return the package."""
    # build.py is loaded by file path: importing the package itself requires an up-to-date library
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bsdfdiff_build", os.path.join(ROOT, "bsdf_diffusion_sampling_b200", "build.py"))
    _build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_build)
    _build.build()
    import bsdf_diffusion_sampling_b200 as pkg
    return pkg
