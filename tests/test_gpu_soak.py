"""Soak test of the persistent kernels' barrier protocols (tools/stress.py): random geometries, forward + backward launched
back to back without host synchronisation, run-to-run identical forward outputs, every 8th case against the SIMT
verification kernels.  Round 2 found a skipped-phase wait in the forward this way (an MMA warp whose tile was absent from
an item did not observe O_FREE for it and could overwrite an accumulator the epilogue was still reading) that none of the
single-launch parity tests hit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [2, 11])
def test_back_to_back_launches_on_random_geometries(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress.py"), "8", str(seed)], capture_output=True,
                       text=True, timeout=110)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-12:])
    assert r.returncode == 0, tail
    assert " 0 bad" in r.stdout, tail


@pytest.mark.gpu
def test_decode_path_is_bit_identical_run_to_run():
    """tools/stress_decode.py: prefill + graph-replayed fused decode steps (dependent-launch chained linear layers) on random
    batch sizes and prompt lengths, every scenario twice: identical tokens and caches; every 4th against HF's layers."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_decode.py"), "8", "3"], capture_output=True,
                       text=True, timeout=110)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-12:])
    assert r.returncode == 0, tail
    assert " 0 bad" in r.stdout, tail
