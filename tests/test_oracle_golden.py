"""Pins the CPU oracle against every golden vector / worked example the reference holds for
the hot path (tests/golden/reference_known_answers.json), and against an independent numpy
statement of the same definitions.  CPU only."""
import numpy as np
import pytest

import oracle as O


def F(rows, dtype):
    return np.asfortranarray(np.array(rows, dtype=dtype))


@pytest.fixture(autouse=True, scope="module")
def _built():
    O.build()


def test_columns_traversal_order(golden):
    # reference test/misc.jl:9-10: matrix traversal is column-major, yielding (column, item)
    g = golden["columns"]
    y = F(g["matrix"]["rows"], np.int64)
    got = [[c + 1, int(v)] for c in range(y.shape[1]) for v in y[:, c]]
    assert got == g["matrix"]["expect"]
    # ... and that is what the oracle's `map` records: delta column = flat position // bag
    cum, mp = O.index(y, 4)
    b = O.buckets(cum, mp)
    assert b == {1: [1], 3: [1], 2: [2], 4: [2]}


def test_histogram_known_answer(golden):
    g = golden["histogram"]
    for _ in range(2):  # reference runs it twice (shallow_empty!), test/misc.jl:57-71
        nnz, order, count = O.histogram_dense(g["A"], g["maxindex"])
        assert nnz == len(g["expect"])
        for k, (o, c) in g["expect"].items():
            assert (order[int(k) - 1], count[int(k) - 1]) == (o, c)
        seen = sorted(range(len(order)), key=lambda i: order[i] or 10**9)[:nnz]
        assert [i + 1 for i in seen] == g["keys_in_insertion_order"]


@pytest.mark.parametrize("dense", [False, True])
def test_index_known_answer(golden, dense):
    g = golden["index"]
    for _ in range(2):
        cum, mp = O.index(g["A"], g["maxindex"], dense=dense)
        assert [list(c) for c in cum] == g["cumulative"]
        assert mp.tolist() == g["map"]


def test_readme_lookup(golden):
    g = golden["readme_lookup"]
    A = O.Table(F(g["data_rows"], np.int64))
    assert O.lookup(A, g["gather"]["inds"]).tolist() == g["gather"]["expect_rows"]
    assert O.lookup(A, F(g["pooled"]["inds_rows"], np.int64)).tolist() == g["pooled"]["expect_rows"]


@pytest.mark.parametrize("strategy", ["default", "simple_parallel", "preallocation"])
def test_readme_maplookup(golden, strategy):
    g = golden["readme_maplookup"]
    tables = [O.Table(F(g["A_rows"], np.int64)), O.Table(F(g["B_rows"], np.int64))]
    for I in ([g["iA"], g["iB"]], np.stack([g["iA"], g["iB"]], axis=1)):
        out = O.maplookup(strategy, tables, I, nthreads=2)
        if strategy == "preallocation":
            assert out.tolist() == g["expect_A_rows"] + g["expect_B_rows"]
        else:
            assert out[0].tolist() == g["expect_A_rows"] and out[1].tolist() == g["expect_B_rows"]


@pytest.mark.parametrize("static", [False, True])
def test_readme_update(golden, static):
    g = golden["readme_update"]
    A = O.Table(np.zeros(g["table_shape"], np.float32, order="F"), static=static)
    O.update(A, F(g["adjoint_rows"], np.float32), g["inds"], g["eta"])
    # README.md:226-231 prints -0.9 where the current source (eta converted to Float32,
    # src/sparseupdate.jl:173) gives Float32(0.1)*9 = 0.90000004: the README example predates
    # that conversion (it also constructs an Int table).  Every entry agrees within 1 ulp,
    # far inside the 1e-5 tolerance of the north star; all other entries are bit-equal.
    expect = F(g["expect_rows"], np.float32)
    assert np.all(np.abs(A.data - expect) <= np.spacing(np.abs(expect)))
    assert np.count_nonzero(A.data != expect) <= 1


# ---- oracle vs an independent numpy statement of README.md:13-25 -------------------------

DIMS = [5, 16, 32, 64, 80, 128, 256, 1504]


def np_lookup(A, I):
    I = np.asarray(I)
    if I.ndim == 1:
        return A[:, I - 1]
    out = A[:, I[0] - 1].copy()
    for i in range(1, I.shape[0]):  # sequential bag order (reference src/lookup.jl:134-147)
        out = out + A[:, I[i] - 1]
    return out


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("static", [False, True])
def test_lookup_matches_numpy(dim, dtype, static):
    rng = np.random.default_rng(dim)
    ncols = 200
    base = (rng.random((dim, ncols)) * 100).astype(dtype, order="F")
    for shard in (None, 30):
        t = O.Table(base, static=static, cols_per_shard=shard)
        I = rng.integers(1, ncols + 1, ncols)
        assert np.array_equal(O.lookup(t, I), np_lookup(base, I))
        I = rng.integers(1, ncols + 1, (12, ncols - 1))
        assert np.array_equal(O.lookup(t, I), np_lookup(base, I))


def test_avx512_and_portable_paths_agree():
    rng = np.random.default_rng(7)
    base = rng.random((64, 300), dtype=np.float32).astype(np.float32, order="F")
    I = rng.integers(1, 301, (33, 257))
    delta = np.asfortranarray(rng.standard_normal((64, 257), dtype=np.float32))
    outs, tabs = [], []
    for portable in (False, True):
        O.force_portable(portable)
        t = O.Table(base.copy(order="F"), static=True)
        outs.append(O.lookup(t, I))
        O.update(t, delta, I, 10.0)
        tabs.append(t.data.copy())
    O.force_portable(False)
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(tabs[0], tabs[1])


def np_update(base, delta, I, eta, fma):
    """dense restatement: accumulate per row from 0 in occurrence order, then one epilogue."""
    I = np.asarray(I)
    flat = I.ravel(order="F")
    bag = 1 if I.ndim == 1 else I.shape[0]
    out = base.copy(order="F")
    acc = {}
    for p, c in enumerate(flat):
        a = acc.setdefault(int(c), np.zeros(base.shape[0], base.dtype))
        a += delta[:, p // bag]
    for c, a in acc.items():
        if fma:
            out[:, c - 1] = (out[:, c - 1].astype(np.float64) - np.float64(np.float32(eta)) * a.astype(np.float64)).astype(base.dtype)
        else:
            out[:, c - 1] = out[:, c - 1] - (base.dtype.type(eta) * a)
    return out


@pytest.mark.parametrize("dim,static", [(64, True), (80, True), (256, True), (64, False), (16, True)])
@pytest.mark.parametrize("reducing", [False, True])
@pytest.mark.parametrize("dense", [False, True])
def test_update_matches_numpy(dim, static, reducing, dense):
    # shapes of reference test/update.jl:1-2,122-162 (f in 64/80/256, 100 rows, eta = 10)
    rng = np.random.default_rng(dim + reducing)
    ncols = 100
    base = np.asfortranarray(rng.standard_normal((dim, ncols), dtype=np.float32))
    I = rng.integers(1, ncols + 1, (10, ncols) if reducing else ncols)
    delta = np.asfortranarray(rng.standard_normal((dim, ncols), dtype=np.float32))
    t = O.Table(base.copy(order="F"), static=static)
    O.update(t, delta, I, 10.0, dense=dense)
    # f32 fma(-eta, acc, row) == round(row - eta*acc) computed exactly; float64 holds the exact
    # product and sum of two f32 values' magnitudes here, so the numpy statement is bit-exact
    fma = static and dim % 16 == 0 and dim * 4 <= 512
    assert np.array_equal(t.data, np_update(base, delta, I, 10.0, fma))
    # uncompress (reference test/update.jl:44-45)
    dense_grad = O.uncompress(delta, I, ncols)
    ref = np.zeros_like(base)
    flat = np.asarray(I).ravel(order="F")
    bag = 10 if reducing else 1
    for p, c in enumerate(flat):
        ref[:, c - 1] += delta[:, p // bag]
    assert np.array_equal(dense_grad, ref)


def test_update_partitions_equal_full():
    # reference test/update.jl:90-120: four IndexerView partial updates == the full update
    rng = np.random.default_rng(3)
    base = np.asfortranarray(rng.standard_normal((16, 100), dtype=np.float32))
    delta = np.asfortranarray(rng.standard_normal((16, 512), dtype=np.float32))
    inds = rng.integers(1, 101, 512)
    A = O.Table(base.copy(order="F"), static=True)
    O.update(A, delta, inds, 1.0)
    B = O.Table(base.copy(order="F"), static=True)
    for s in range(1, 5):
        O.update(B, delta, inds, 1.0, split=(4, s))
    assert np.array_equal(A.data, B.data)


@pytest.mark.parametrize("nthreads", [1, 3])
def test_ensemble_update_equals_single(nthreads):
    rng = np.random.default_rng(11)
    bases = [np.asfortranarray(rng.standard_normal((32, 50), dtype=np.float32)) for _ in range(5)]
    Is = [rng.integers(1, 51, (4, 64)) for _ in bases]
    big = np.asfortranarray(rng.standard_normal((8 + 5 * 32, 64), dtype=np.float32))
    deltas = [big[8 + 32 * k: 8 + 32 * (k + 1), :] for k in range(5)]  # Preallocation pullback views
    single = [O.Table(b.copy(order="F"), static=True) for b in bases]
    for t, d, i in zip(single, deltas, Is):
        O.update(t, d, i, 0.5)
    ens = [O.Table(b.copy(order="F"), static=True) for b in bases]
    O.update_ensemble(ens, deltas, Is, 0.5, nthreads=nthreads)
    for a, b in zip(single, ens):
        assert np.array_equal(a.data, b.data)


@pytest.mark.parametrize("strategy", ["default", "simple_parallel", "preallocation"])
@pytest.mark.parametrize("form", ["vecvec", "matrix", "vecmat", "3d"])
def test_maplookup_forms(strategy, form):
    # reference test/map.jl:14-100: 10 tables x (16|64|512) x 100, batch 64, bag 10
    rng = np.random.default_rng(5)
    for nrows in (16, 64, 512):
        base = [np.asfortranarray(rng.standard_normal((nrows, 100), dtype=np.float32)) for _ in range(10)]
        tables = [O.Table(b, static=True) for b in base]
        if form == "vecvec":
            I = [rng.integers(1, 101, 64) for _ in base]
        elif form == "matrix":
            I = rng.integers(1, 101, (64, 10))
        elif form == "vecmat":
            I = [rng.integers(1, 101, (10, 64)) for _ in base]
        else:
            I = rng.integers(1, 101, (10, 64, 10))
        ref = np.concatenate([np_lookup(b, i) for b, i in zip(base, O.colwrap(tables, I))], axis=0)
        out = O.maplookup(strategy, tables, I, nthreads=4)
        got = out if strategy == "preallocation" else np.concatenate(out, axis=0)
        assert np.array_equal(got, ref)
    # prependrows: rows 1..prepend untouched (reference src/lookup.jl:311-313, 334-340)
    out = np.full((20 + 10 * 512, 64), 7.0, np.float32, order="F")
    O.maplookup("preallocation", tables, I, prependrows=20, nthreads=2, out=out)
    assert np.all(out[:20] == 7.0) and np.array_equal(out[20:], ref)
