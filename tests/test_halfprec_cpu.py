"""CPU checks of the half-precision extension's oracle (the definition of its semantics) and of the
host-side dtype plumbing.  No GPU needed."""
import numpy as np
import pytest

import oracle as O

ml_dtypes = pytest.importorskip("ml_dtypes")
DTYPES = [np.dtype(np.float16), np.dtype(ml_dtypes.bfloat16)]


@pytest.mark.parametrize("dt", DTYPES)
def test_lookup_lowp_accumulates_in_float32(dt):
    base = np.zeros((4, 40), np.float32)
    base[:, 0] = 1.0
    base[:, 1:] = 2.0 ** -9
    base = np.asfortranarray(base.astype(dt))
    I = np.arange(1, 41).reshape(40, 1)
    got = O.lookup_lowp(base, I)
    assert got.dtype == dt and got.shape == (4, 1)
    assert np.all(got.astype(np.float64) == np.float64(np.float32(1.0 + 39 * 2.0 ** -9).astype(dt)))
    # a gather is a bit copy
    Iv = np.array([3, 1, 3])
    assert np.array_equal(O.lookup_lowp(base, Iv).view(np.uint16), base[:, Iv - 1].view(np.uint16))


@pytest.mark.parametrize("dt", DTYPES)
def test_update_lowp_known_answer(dt):
    # row 2 receives columns 1 and 3 (in that order), row 5 column 2; eta = 0.5; all values exact in both types
    data = np.asfortranarray(np.arange(12, dtype=np.float32).reshape(2, 6, order="F").astype(dt))
    delta = np.asfortranarray(np.array([[1.0, 4.0, 0.25], [2.0, -8.0, 0.5]], np.float32).astype(dt))
    before = data.astype(np.float32).copy()
    O.update_lowp(data, delta, np.array([2, 5, 2]), 0.5)
    after = data.astype(np.float32)
    want = before.copy()
    want[:, 1] -= 0.5 * np.array([1.25, 2.5], np.float32)
    want[:, 4] -= 0.5 * np.array([4.0, -8.0], np.float32)
    assert np.array_equal(after, want)


def test_bfloat16_bridge_round_trips():
    torch = pytest.importorskip("torch")
    import embtab.darray as D
    assert D.bfloat16 == np.dtype(ml_dtypes.bfloat16)
    a = np.arange(-8, 8, dtype=np.float32).astype(D.bfloat16)
    t = D._np_to_torch(a)
    assert t.dtype == torch.bfloat16 and torch.equal(t.float(), torch.arange(-8, 8, dtype=torch.float32))
    assert np.array_equal(D._torch_to_np(t).view(np.uint16), a.view(np.uint16))
    assert D._NP2ELT[D.bfloat16] == 5 and D._NP2ELT[np.dtype(np.float16)] == 4
