"""Shared assertions of the parity tests (tolerances are stated where used)."""
import numpy as np


def check_inputs(gold, *tensors):
    """The golden outputs belong to these exact inputs: verify the RNG reproduced them."""
    import cases
    cs = cases.checksum(*tensors)
    ref = float(gold["inputs_checksum"])
    assert abs(cs - ref) <= 1e-9 * max(1.0, abs(ref)), (
        f"seeded inputs differ from the ones the golden file was made from ({cs} vs {ref}); "
        "torch's CPU generator changed — regenerate with tests/golden/make_golden.py")


def max_abs(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b)
    d[both_nan] = 0.0
    return float(np.nanmax(d)) if d.size else 0.0


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    n = int((a != b).sum())
    assert n == 0, f"{what}: {n} of {a.size} elements differ"


def unpack_bits(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)
