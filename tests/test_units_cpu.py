"""Host-side unit tests of device helpers that are plain integer logic: tests/units/*.cu are compiled by nvcc and
run on the CPU (no CUDA API call, no GPU)."""
import os
import shutil
import subprocess
import tempfile

import pytest

from conftest import ROOT


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        pytest.skip("nvcc not available")
    return exe


def test_flat_staging_slot_map_equals_the_segment_enumeration():
    """tile_entry_of_slot (csrc/tile.cuh, SPHB_FLAT_STAGING) against the enumeration of SPHB_TILE_FOR_STAGED on
    20 000 random segment tables with empty segments anywhere."""
    src = os.path.join(ROOT, "tests", "units", "tile_units.cu")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "tile_units")
        subprocess.run([_nvcc(), "-std=c++17", "-O1", "-DSPHB_FLAT_STAGING=1", "-arch=sm_100a", src, "-o", exe],
                       check=True, capture_output=True)
        r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
