"""Kernel logic on a machine without a GPU: a subset of the GPU parity tests (tests/test_gpu_parity.py, unchanged) run in a
subprocess against tests/emu/_build/libsvbfm_emu.so -- csrc/*.cu compiled for the host on a lockstep CUDA emulator (every
CUDA thread a fiber, warp primitives and __syncthreads as barriers; tests/emu/include/cuda_runtime.h). This checks index
arithmetic, warp-synchronous code and the engine's launch order against the oracle; it says nothing about the B200 (speed,
memory system, races between CTAs): the `-m gpu` tests do. The package itself never loads the emulated library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# quick ones that between them reach k_stream (VB, MCMC, small tiles, x != 1), the general per-run schedule, k_predict /
# k_predict2, the finalize / hyper-parameter kernels and the device ingest
SUBSET = ["test_vb_two_field_onehot", "test_vb_ragged_multihot", "test_vb_groups_and_small_tiles", "test_mcmc_als_exact",
          "test_vb_values_no_reorder", "test_stream_schedule_sorted_input", "test_vb_k_zero_and_empty_test_columns"]


@pytest.fixture(scope="module")
def emu_lib(built):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    return build_emu.build()


def test_emulated_library_exports_the_abi(emu_lib):
    import ctypes
    import svbfm_b200 as sv
    L = ctypes.CDLL(emu_lib)
    for name in sv.ABI_SYMBOLS:
        assert hasattr(L, name), name


def test_parity_subset_on_the_emulator(emu_lib):
    env = dict(os.environ, SVBFM_LIB=emu_lib)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           "-k", " or ".join(SUBSET)]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "passed" in r.stdout
