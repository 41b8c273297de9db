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
    na, nb = np.isnan(a), np.isnan(b)
    one_sided = int((na != nb).sum())
    # a NaN on one side only is a mismatch, never something to skip over
    assert one_sided == 0, f"{one_sided} of {a.size} elements are NaN on one side only"
    # infinities must agree exactly (inf - inf would be NaN and hide a sign error)
    ia, ib = np.isinf(a), np.isinf(b)
    bad_inf = int(((ia | ib) & ~na & (a != b)).sum())
    assert bad_inf == 0, f"{bad_inf} of {a.size} elements are infinite on one side only (or differ in sign)"
    fin = ~(na | ia)
    return float(np.abs(a[fin] - b[fin]).max()) if fin.any() else 0.0


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    n = int((a != b).sum())
    assert n == 0, f"{what}: {n} of {a.size} elements differ"


def unpack_bits(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)
