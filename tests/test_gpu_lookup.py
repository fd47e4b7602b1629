"""GPU parity: lookup / lookup! / maplookup -- the reference's test/lookup.jl, test/map.jl and
test/constructors.jl re-expressed against the C-ABI library, with the CPU oracle as checker.
Bit-exact (`==`) everywhere: gathers are bit copies and pooled sums keep the reference's
left-to-right association."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import embtab
    return embtab


@pytest.fixture(scope="module")
def O():
    import oracle
    oracle.build()
    return oracle


NROWS = [32, 64, 128, 256, 512, 1024, 1504]  # reference test/lookup.jl:67
NCOLS = 1000


def make_tables(E, O, base, kind, shard=None):
    if kind == "dynamic":
        return E.SimpleEmbedding(base), O.Table(base, static=False)
    if kind == "static":
        return E.SimpleEmbedding(base, E.Static(base.shape[0])), O.Table(base, static=True)
    return E.SplitEmbedding(base, shard), O.Table(base, cols_per_shard=shard)


def test_readme_examples(E, golden):
    g = golden["readme_lookup"]
    A = E.SimpleEmbedding(np.array(g["data_rows"], np.int64))
    assert E.lookup(A, g["gather"]["inds"]).numpy().tolist() == g["gather"]["expect_rows"]
    assert E.lookup(A, np.array(g["pooled"]["inds_rows"])).numpy().tolist() == g["pooled"]["expect_rows"]
    g = golden["readme_maplookup"]
    tables = [E.SimpleEmbedding(np.array(g["A_rows"], np.int64)), E.SimpleEmbedding(np.array(g["B_rows"], np.int64))]
    res = E.maplookup(tables, [g["iA"], g["iB"]])
    assert res[0].numpy().tolist() == g["expect_A_rows"] and res[1].numpy().tolist() == g["expect_B_rows"]
    res2 = E.maplookup(tables, np.stack([g["iA"], g["iB"]], axis=1))
    assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(res, res2))


def test_constructors(E, golden):
    # reference test/constructors.jl:1-25
    even, odd = np.random.rand(64, 10).astype(np.float32), np.random.rand(65, 10).astype(np.float32)
    x = E.SimpleEmbedding(even, E.Static(64))
    assert x.size() == even.shape
    with pytest.raises(E.ArgumentError):
        E.SimpleEmbedding(even, E.Static(32))
    with pytest.raises(E.ArgumentError):
        E.SimpleEmbedding(even, E.Static(64.0))
    x = E.SimpleEmbedding(odd)
    assert x.size() == odd.shape and isinstance(x.lookup_type, E.Dynamic)
    x = E.SimpleEmbedding(odd, E.Static(65))
    assert x.size() == odd.shape and x.lookup_type == E.Static(65)


def test_custom_table_type(E, O):
    # reference test/constructors.jl:34-54: a user-defined table implementing only the contract
    class DummyEmbedding(E.AbstractEmbeddingTable):
        def __init__(self, data):
            self.data = E.DeviceArray.from_numpy(data)
            self.lookup_type, self.dtype = E.Dynamic(), self.data.dtype

        def size(self, d=None):
            return self.data.size(d)

        def columnpointer(self, i, ctx=None):
            return E.columnpointer(self.data, i)

        def example(self):
            return self.data

        def descriptor(self):
            d = self.data
            from embtab import _lib
            return _lib.Table(d.ptr, None, d.shape[1], 0, d.shape[0], d.ld, d.elt, 0)

    rng = np.random.default_rng(0)
    base = rng.standard_normal((10, 10)).astype(np.float32)
    table = DummyEmbedding(base)
    inds = rng.integers(1, 11, (10, 10))
    assert np.array_equal(E.lookup(table, inds).numpy(), O.lookup(O.Table(base), inds))


@pytest.mark.parametrize("rows", NROWS)
@pytest.mark.parametrize("kind", ["dynamic", "static"])
def test_simple_lookup(E, O, rows, kind):
    rng = np.random.default_rng(rows)
    base = rng.random((rows, NCOLS), dtype=np.float32)
    table, ref = make_tables(E, O, base, kind)
    assert table == base and len(table) == base.size          # `table == baseline`, test/lookup.jl:78
    for _ in range(3):
        I = rng.permutation(NCOLS) + 1                         # no repeats, test/lookup.jl:14-20
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))
        I = rng.integers(1, NCOLS + 1, NCOLS)                  # repeats, :23-29
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))
        I = np.stack([rng.permutation(NCOLS - 1) + 2 for _ in range(12)])  # bag 12, batch 999, :42-48
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))
        I = rng.integers(1, NCOLS + 1, (12, NCOLS))            # :51-56
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))


@pytest.mark.parametrize("rows", NROWS)
@pytest.mark.parametrize("shard", [10, 30, 50])
def test_split_lookup(E, O, rows, shard):
    # reference test/lookup.jl:110-138 (ragged last chunk for 30)
    rng = np.random.default_rng(rows + shard)
    base = rng.random((rows, NCOLS), dtype=np.float32)
    table, ref = make_tables(E, O, base, "split", shard)
    assert table == base
    I = rng.integers(1, NCOLS + 1, NCOLS)
    assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))
    I = rng.integers(1, NCOLS + 1, (12, NCOLS))
    assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("dim", [1, 3, 5, 6, 12, 16, 20, 33, 48, 80, 96, 100, 130, 160, 192, 320, 384, 640])
def test_odd_dims_and_dtypes(E, O, dim, dtype):
    # feature sizes that force the 8- and 4-byte vector paths and partially filled groups;
    # integer tables must be bit-exact (wrapping adds)
    rng = np.random.default_rng(dim)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        base = rng.integers(info.min, info.max, (dim, 300), dtype=dtype)
    else:
        base = (rng.standard_normal((dim, 300)) * 1e3).astype(dtype)
    table, ref = E.SimpleEmbedding(base), O.Table(base)
    for bag in (1, 2, 7, 32, 33, 70):
        I = rng.integers(1, 301, (bag, 257))
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I)), (dim, dtype, bag)
    I = rng.integers(1, 301, 1001)
    assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(ref, I))


def test_int32_indices_and_negative_zero(E, O):
    rng = np.random.default_rng(1)
    base = rng.standard_normal((64, 50)).astype(np.float32)
    base[:, 0] = -0.0   # a bag of only -0.0 rows must give -0.0 (accumulator seeded with row 1)
    base[:, 1] = 0.0
    table, ref = E.SimpleEmbedding(base, E.Static(64)), O.Table(base, static=True)
    I = rng.integers(1, 51, (9, 40))
    I[:, 0] = 1
    I[:, 1] = [1, 2, 1, 2, 1, 2, 1, 2, 1]
    got32 = E.lookup(table, I.astype(np.int32)).numpy()
    got64 = E.lookup(table, I).numpy()
    want = O.lookup(ref, I)
    assert np.array_equal(got32.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got64.view(np.uint32), want.view(np.uint32))
    assert np.all(np.signbit(got64[:, 0])) and not np.any(np.signbit(got64[:, 1]))


def test_empty_and_single(E, O):
    base = np.arange(40, dtype=np.float32).reshape(8, 5, order="F")
    table = E.SimpleEmbedding(base)
    assert E.lookup(table, np.zeros(0, np.int64)).shape == (8, 0)
    assert E.lookup(table, np.zeros((3, 0), np.int64)).shape == (8, 0)
    assert np.array_equal(E.lookup(table, [5]).numpy(), base[:, 4:5])
    assert np.array_equal(E.lookup(table, np.array([[5], [1]])).numpy(), base[:, 4:5] + base[:, 0:1])


def test_lookup_inplace_strided_destination(E, O):
    # lookup!(view(dst, rows, :), ...) -- what the PreallocationStrategy does per table
    rng = np.random.default_rng(2)
    base = rng.standard_normal((32, 100)).astype(np.float32)
    table = E.SimpleEmbedding(base, E.Static(32))
    big = E.DeviceArray.from_numpy(np.full((100, 64), 7.0, np.float32))
    I = rng.integers(1, 101, (5, 64))
    E.lookup_(big.rows(10, 42), table, I)
    got = big.numpy()
    assert np.array_equal(got[10:42], O.lookup(O.Table(base, static=True), I))
    assert np.all(got[:10] == 7.0) and np.all(got[42:] == 7.0)


# ---------------------------------------------------------------------------- maplookup
def strategies(E):
    return [E.DefaultStrategy(), E.SimpleParallelStrategy(), E.PreallocationStrategy()]


@pytest.mark.parametrize("nrows", [16, 64, 512])
@pytest.mark.parametrize("form", ["vecvec", "matrix", "vecmat", "3d"])
def test_maplookup_forms(E, O, nrows, form):
    # reference test/map.jl:14-100
    rng = np.random.default_rng(nrows)
    ncols, ntables, nlookups, batch = 100, 10, 10, 64
    for rep in range(3):
        base = [rng.standard_normal((nrows, ncols)).astype(np.float32) for _ in range(ntables)]
        tables = [E.SimpleEmbedding(b, E.Static(nrows)) for b in base]
        refs = [O.Table(b, static=True) for b in base]
        if form == "vecvec":
            inds = [rng.integers(1, ncols + 1, batch) for _ in range(ntables)]
        elif form == "matrix":
            inds = rng.integers(1, ncols + 1, (batch, ntables))
        elif form == "vecmat":
            inds = [rng.integers(1, ncols + 1, (nlookups, batch)) for _ in range(ntables)]
        else:
            inds = rng.integers(1, ncols + 1, (nlookups, batch, ntables))
        reference = np.concatenate([O.lookup(r, i) for r, i in zip(refs, O.colwrap(refs, inds))], axis=0)
        for s in strategies(E):
            out = E.maplookup(s, tables, inds)
            got = out.numpy() if isinstance(s, E.PreallocationStrategy) else np.concatenate([o.numpy() for o in out], axis=0)
            assert np.array_equal(got, reference), (form, type(s).__name__)
    out = E.maplookup(tables, inds)  # default strategy when omitted
    assert np.array_equal(np.concatenate([o.numpy() for o in out], axis=0), reference)


def test_preallocation_prependrows_untouched(E, O):
    rng = np.random.default_rng(9)
    dims = [16, 64, 5, 128]  # mixed feature sizes -> several kernel classes in one call
    base = [rng.standard_normal((d, 77)).astype(np.float32) for d in dims]
    tables = [E.SimpleEmbedding(b) for b in base]
    inds = [rng.integers(1, 78, (4, 50)) for _ in dims]
    ref = np.concatenate([O.lookup(O.Table(b), i) for b, i in zip(base, inds)], axis=0)
    for prepend in (0, 20, 3):
        dst = E.DeviceArray.from_numpy(np.full((prepend + sum(dims), 50), -3.0, np.float32))
        out = E.maplookup_(E.PreallocationStrategy(prepend), dst, tables, inds)
        assert out is dst
        got = dst.numpy()
        assert np.all(got[:prepend] == -3.0)          # reference src/lookup.jl:311-313: rows left alone
        assert np.array_equal(got[prepend:], ref)
    out = E.maplookup(E.PreallocationStrategy(20), tables, inds)
    assert out.shape == (20 + sum(dims), 50) and np.array_equal(out.numpy()[20:], ref)


def test_maplookup_split_tables_many(E, O):
    # more tables than one launch holds (kMaxItems = 256), split storage, ragged chunks
    rng = np.random.default_rng(4)
    base = [rng.standard_normal((32, 64)).astype(np.float32) for _ in range(300)]
    tables = [E.SplitEmbedding(b, 24) for b in base]
    inds = rng.integers(1, 65, (3, 40, 300))
    out = E.maplookup(E.PreallocationStrategy(), tables, inds).numpy()
    ref = np.concatenate([O.lookup(O.Table(b, cols_per_shard=24), inds[:, :, t]) for t, b in enumerate(base)], axis=0)
    assert np.array_equal(out, ref)


def test_missing_library_fails_loudly(E, monkeypatch):
    from embtab import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libembtab_b200.so")
    with pytest.raises(E.EmbTabError):
        E.lookup(E.SimpleEmbedding(np.zeros((4, 4), np.float32)), [1])


def test_direct_c_entry_points(E, O):
    # the single-table C entry points a Julia `ccall` would use (the mirror itself goes through etb_maplookup):
    # etb_gather, etb_pooled_sum, etb_index_and_update, runtime alloc/copy/stream
    import ctypes as C

    from embtab import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(77)
    dim, nrows, bag, batch = 64, 500, 6, 300
    base = np.asfortranarray(rng.standard_normal((dim, nrows)).astype(np.float32))
    I = np.asfortranarray(rng.integers(1, nrows + 1, (bag, batch)))
    stream = C.c_void_p()
    _lib.check(lib.etb_stream_create(C.byref(stream)))
    ptrs = []

    def dmalloc(nbytes):
        p = C.c_void_p()
        _lib.check(lib.etb_malloc(C.byref(p), nbytes))
        ptrs.append(p)
        return p

    d_table, d_idx, d_out = dmalloc(base.nbytes), dmalloc(I.nbytes), dmalloc(dim * batch * 4)
    _lib.check(lib.etb_memcpy_h2d(d_table, base.ctypes.data, base.nbytes, stream))
    _lib.check(lib.etb_memcpy_h2d(d_idx, I.ctypes.data, I.nbytes, stream))
    t = _lib.Table(d_table.value, None, nrows, 0, dim, dim, _lib.F32, 0)
    out = np.empty((dim, batch), np.float32, order="F")
    _lib.check(lib.etb_pooled_sum(d_out, dim, C.byref(t), d_idx, _lib.I64, bag, batch, bag, stream))
    _lib.check(lib.etb_memcpy_d2h(out.ctypes.data, d_out, out.nbytes, stream))
    _lib.check(lib.etb_stream_sync(stream))
    assert np.array_equal(out, O.lookup(O.Table(base), I))
    _lib.check(lib.etb_gather(d_out, dim, C.byref(t), d_idx, _lib.I64, batch, stream))   # first `batch` indices as a vector
    _lib.check(lib.etb_memcpy_d2h(out.ctypes.data, d_out, out.nbytes, stream))
    _lib.check(lib.etb_stream_sync(stream))
    assert np.array_equal(out, O.lookup(O.Table(base), I.ravel(order="F")[:batch]))
    # update!: etb_index_workspace_bytes + etb_index_and_update
    delta = np.asfortranarray(rng.standard_normal((dim, batch)).astype(np.float32))
    d_delta = dmalloc(delta.nbytes)
    _lib.check(lib.etb_memcpy_h2d(d_delta, delta.ctypes.data, delta.nbytes, stream))
    item = _lib.UpdateItem(t, d_delta.value, dim, d_idx.value, batch, bag, bag, _lib.I64, _lib.UPDATE_FMA)
    need = C.c_size_t()
    _lib.check(lib.etb_index_workspace_bytes(C.byref(item), 1, C.byref(need)))
    ws = dmalloc(need.value)
    assert lib.etb_index_and_update(ws, need.value - 1, C.byref(item), 1, 0.5, 0, stream) == 3   # ETB_ERR_WORKSPACE
    assert b"workspace" in lib.etb_last_error()
    _lib.check(lib.etb_index_and_update(ws, need.value, C.byref(item), 1, 0.5, 0, stream))
    got = np.empty_like(base)
    _lib.check(lib.etb_memcpy_d2h(got.ctypes.data, d_table, got.nbytes, stream))
    _lib.check(lib.etb_stream_sync(stream))
    ref = O.Table(base.copy(order="F"), static=True)
    O.update(ref, delta, I, 0.5)
    assert np.array_equal(got, ref.data)
    for p in ptrs:
        _lib.check(lib.etb_free(p))
    _lib.check(lib.etb_stream_destroy(stream))


def test_strided_upload_download_of_row_slices(E):
    # one table's rows of a concatenated (column-major) matrix move as a single 2-D copy, both directions,
    # without touching the neighbouring rows (etb_memcpy2d_h2d / etb_memcpy2d_d2h)
    import torch
    rng = np.random.default_rng(5)
    host = E.pinned_empty((40, 33), np.float32)
    host[...] = rng.standard_normal((40, 33)).astype(np.float32)
    dev = E.DeviceArray.zeros((40, 33))
    dev.rows(8, 24).upload(host[8:24])                       # strided on both sides
    torch.cuda.synchronize()
    got = dev.numpy()
    assert np.array_equal(got[8:24], host[8:24]) and np.all(got[:8] == 0) and np.all(got[24:] == 0)
    dense = np.asfortranarray(rng.standard_normal((16, 33)).astype(np.float32))
    dev.rows(24, 40).upload(dense)                           # contiguous host, strided device
    back = E.pinned_empty((40, 33), np.float32)
    back[...] = -7.0
    dev.rows(24, 40).download(back[2:18])                    # strided device, strided host
    torch.cuda.synchronize()
    assert np.array_equal(back[2:18], dense) and np.all(back[:2] == -7.0) and np.all(back[18:] == -7.0)
    with pytest.raises(AssertionError):
        dev.rows(0, 8).upload(host[::2][:8])                 # not a row-slice view


def test_preallocation_eltype_override(E, O):
    # PreallocationStrategy{U}(prependrows) with U != eltype(tables) (reference src/lookup.jl:293-294, 312): the output
    # matrix has element type U; a non-reducing lookup is convert(U, x) exactly, a pooled sum is the tables'-type sum
    # converted once; the prepend rows stay the caller's
    rng = np.random.default_rng(77)
    base = [rng.standard_normal((32, 200)).astype(np.float32) for _ in range(3)]
    tables = [E.SimpleEmbedding(b, E.Static(32)) for b in base]
    refs = [O.Table(b, static=True) for b in base]
    for I in (rng.integers(1, 201, (50, 3)), rng.integers(1, 201, (6, 50, 3))):
        out = E.maplookup(E.PreallocationStrategy(4, eltype=np.float64), tables, I)
        assert out.dtype == np.float64 and out.shape == (4 + 96, 50)
        want = O.maplookup("preallocation", refs, I, prependrows=4, out=np.zeros((100, 50), np.float32, order="F"))
        assert np.array_equal(out.numpy()[4:], want[4:].astype(np.float64))
    pre = E.DeviceArray.from_numpy(np.full((100, 50), -3.0, np.float64))
    E.maplookup_(E.PreallocationStrategy(4, eltype=np.float64), pre, tables, rng.integers(1, 201, (50, 3)))
    assert np.all(pre.numpy()[:4] == -3.0)
