"""CPU check of the oracle's row-wise Adagrad (the definition of etb_adagrad_update's result) against the
textbook formula in Float64.  No GPU needed."""
import numpy as np
import pytest

import oracle as O


@pytest.mark.parametrize("dim", [128, 80, 6, 4])
def test_adagrad_oracle_matches_textbook(dim):
    rng = np.random.default_rng(dim)
    data = np.asfortranarray(rng.standard_normal((dim, 50)).astype(np.float32))
    state = np.zeros(50, np.float32)
    ref, rs = data.astype(np.float64), np.zeros(50)
    for _ in range(3):
        I = rng.integers(1, 51, (4, 60))
        delta = np.asfortranarray(rng.standard_normal((dim, 60)).astype(np.float32))
        O.adagrad_update(data, state, delta, I, 0.05, 1e-6)
        g = {}
        for p, r in enumerate(I.reshape(-1, order="F")):
            g[r] = g.get(r, 0) + delta[:, p // 4].astype(np.float64)
        for r, v in g.items():
            rs[r - 1] += np.mean(v * v)
            ref[:, r - 1] -= 0.05 / (np.sqrt(rs[r - 1]) + 1e-6) * v
    assert np.allclose(data, ref, rtol=1e-5, atol=1e-6) and np.allclose(state, rs, rtol=1e-5)


def test_adagrad_layout_rule():
    # (vectors per row, lanes per row, elements per vector) as the kernels choose them for dense aligned rows
    assert O.adagrad_layout(128, 4) == (32, 32, 4)
    assert O.adagrad_layout(128, 2) == (16, 16, 8)
    assert O.adagrad_layout(80, 4) == (20, 32, 4)
    assert O.adagrad_layout(6, 2) == (3, 4, 2)
    assert O.adagrad_layout(6, 8) == (3, 4, 2)
    assert O.adagrad_layout(4, 4) == (1, 1, 4)
