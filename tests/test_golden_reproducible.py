"""The committed golden vectors ARE the reference's outputs: where the reference tree is available (the build
container, not the GPU box), re-run the generator into a scratch directory and compare every array bit for bit."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

REF = os.environ.get("STITCH_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "core")), reason="reference tree not present (GPU box)")
def test_generator_reproduces_committed_golden_vectors(tmp_path):
    env = dict(os.environ, STITCH_GOLDEN_OUT=str(tmp_path), STITCH_REFERENCE=REF)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py")], capture_output=True,
                       text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    committed = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    assert len(committed) >= 27
    for f in committed:
        new = os.path.join(str(tmp_path), os.path.basename(f))
        assert os.path.exists(new), f"generator no longer writes {os.path.basename(f)}"
        a, b = np.load(f), np.load(new)
        assert sorted(a.files) == sorted(b.files), os.path.basename(f)
        for k in a.files:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (os.path.basename(f), k)
            assert np.array_equal(a[k], b[k], equal_nan=True), f"{os.path.basename(f)}:{k} is not what the reference computes"
