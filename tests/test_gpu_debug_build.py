"""The parity suite against the bounds-checking build (-DK3D_DEBUG): every data-dependent
index into the shared-memory carve-up of both kernels is checked and a violation traps the
kernel.  compute-sanitizer is not available on the GPU pool; this is what stands in for
memcheck.  Runs in a subprocess (K3D_LIB selects the library a process loads)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, "pygeostatistics_b200", "csrc", "libk3d_debug.so")

# edge cases (every kernel instantiation, list overflow, dense tiles, bricks, listed locations,
# block kriging, small ndmax on wide CTAs) rather than the full-size runs
SELECTED = [
    "tests/test_gpu_parity.py::test_scaled_configs_against_oracle",
    "tests/test_gpu_parity.py::test_edge_cases_against_oracle",
    "tests/test_gpu_parity.py::test_dense_tile_falls_back_to_global_walk",
    "tests/test_gpu_parity.py::test_unordered_selection_path_and_plane_runs",
    "tests/test_gpu_parity.py::test_ked_and_nonstationary_sk",
    "tests/test_gpu_parity.py::test_cross_validation_against_oracle",
    "tests/test_gpu_parity.py::test_jackknife_locations_and_grid_equivalence",
    "tests/test_gpu_robustness.py::test_small_ndmax_on_large_tiles",
    "tests/test_gpu_robustness.py::test_bricks_of_super_blocks",
    "tests/test_gpu_robustness.py::test_largest_supported_system",
    "tests/test_gpu_robustness.py::test_duplicate_samples_are_singular",
    "tests/test_gpu_robustness.py::test_random_parameter_sets_against_oracle",
]


def test_parity_cases_on_the_bounds_checking_build():
    if not os.path.exists(DEBUG_LIB):
        from pygeostatistics_b200.csrc import build
        build.build_debug()
    env = dict(os.environ, K3D_LIB=DEBUG_LIB)
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                          "-m", "gpu"] + SELECTED, cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=3000)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert "k3d debug:" not in out.stdout + out.stderr, tail
    assert out.returncode == 0, tail
    assert " passed" in out.stdout, tail
