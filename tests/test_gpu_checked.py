"""The whole path through the bounds-checked build of the library.

compute-sanitizer is not available on the GPU pool, so `make checked` compiles the same
sources with every table / tile-buffer / output index turned into a device-side assert
(CGB_CHECK_INDEX, cgb_common.cuh) and this test runs scripts/sanitize_target.py -- every
kernel, odd and tile-boundary sizes, both gas kernels, host streaming, the deferred
long-tail pass, ICs + resampling -- against it in a process of its own.  A control run
proves the asserts are live: an out-of-range probe index must trap there and must not in
the shipped library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cluster_generator_b200", "lib")
CHECKED = os.path.join(LIBDIR, "libcgb200_checked.so")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def checked_library():
    """build() makes the checked library; on a tree where it is missing, make it here (nvcc is in
    the image), and skip -- not fail -- when that is impossible."""
    if not os.path.exists(CHECKED):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cluster_generator_b200", "csrc"), "-j4", "checked"],
                       capture_output=True, timeout=1200)
    if not os.path.exists(CHECKED):
        pytest.skip("libcgb200_checked.so is not built (make -C cluster_generator_b200/csrc checked)")


def _run(code_or_script, lib, script=False):
    env = dict(os.environ, CGB200_LIB=lib)
    cmd = [sys.executable, code_or_script] if script else [sys.executable, "-c", code_or_script]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)


PROBE = ("import sys; sys.path.insert(0, '.'); from cluster_generator_b200 import _native as N; "
         "L = N.lib(); print('enabled', L.cgb_bounds_check_enabled()); "
         "print('in', L.cgb_test_bounds_check(2, 3)); print('out', L.cgb_test_bounds_check(3, 3))")


def test_checked_build_traps_an_out_of_range_index_and_the_shipped_one_does_not():
    r = _run(PROBE, CHECKED)
    assert "enabled 1" in r.stdout and "in 0" in r.stdout, r.stdout + r.stderr
    assert "out 0" not in r.stdout, "the out-of-range probe did not trap in the checked build"
    r = _run(PROBE, os.path.join(LIBDIR, "libcgb200.so"))
    assert "enabled 0" in r.stdout and "in 0" in r.stdout and "out 0" in r.stdout, r.stdout + r.stderr


def test_every_kernel_on_the_path_runs_clean_under_the_bounds_checked_build():
    r = _run(os.path.join(ROOT, "scripts", "sanitize_target.py"), CHECKED, script=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "bounds checks enabled: 1" in r.stdout and "\nok " in r.stdout, r.stdout[-2000:]
    assert "Assertion" not in r.stderr and "device-side assert" not in r.stderr, r.stderr[-4000:]
