"""Table-wise sharded ensembles over the GPUs of one NVSwitch box.

The reference is single-process (no NCCL/MPI anywhere); this layer is the
multi-GPU form the north star asks for: tables are block-partitioned over ranks (one process per
GPU), every rank looks its tables up for the GLOBAL batch, and one all-to-all turns
"all samples x my tables" into "my samples x all tables" -- the concatenated DLRM feature matrix
of maplookup(PreallocationStrategy(prependrows), ...) for this rank's batch slice.  The backward
pass is the reverse all-to-all of the cotangent's row blocks, after which each owner runs the
ordinary ensemble update! on its tables.  Arithmetic per output column is exactly the 1-GPU
path's, so results are bit-identical to it (only placement changes).

Layout trick (K6): the lookup kernel writes straight into the all-to-all SEND layout (block p =
my_rows x cols[p], dense; end to end the blocks are just the column-major my_rows x B_global
matrix), so there is no pack pass in the forward; in the
backward the RECEIVE buffer already is the (my_rows x global_batch) cotangent of my tables, so
there is no unpack pass there.  The remaining strided copies (forward unpack, backward pack) are
etb_a2a_unpack / etb_a2a_pack.

ShardPlan and exchange() are pure host logic over torch tensors of any device; the CPU tests
drive them with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .darray import DeviceArray, as_device_indices, current_stream_ptr
from .lookup import _item, _run
from .sparseupdate import Indexer, SparseEmbeddingUpdate, update_
from .sparseupdate import prefetch_index as _prefetch_index
from .tables import featuresize


class ShardPlan:
    """Who owns which tables, which rows of the feature matrix they fill, and which batch
    columns each rank keeps.

    dims          featuresize of every table of the GLOBAL ensemble, in ensemble order
    world, rank   torch.distributed world
    prependrows   PreallocationStrategy.prependrows of the feature matrix
    batch_global  global batch (columns of the index arrays every owner receives)
    """

    def __init__(self, dims, world: int, rank: int, prependrows: int, batch_global: int):
        self.dims = [int(d) for d in dims]
        self.world, self.rank = int(world), int(rank)
        self.prependrows, self.batch_global = int(prependrows), int(batch_global)
        T = len(self.dims)
        # block partition of tables: rank q owns tables [tlo[q], thi[q])
        bounds = np.floor(np.arange(self.world + 1) * T / self.world + 0.5).astype(int)
        self.tlo, self.thi = bounds[:-1].tolist(), bounds[1:].tolist()
        self.rows = [sum(self.dims[a:b]) for a, b in zip(self.tlo, self.thi)]       # feature rows per owner
        self.row_off = [self.prependrows + sum(self.rows[:q]) for q in range(self.world)]
        self.total_rows = self.prependrows + sum(self.dims)
        # block partition of the batch: rank p keeps columns [clo[p], chi[p])
        cb = np.floor(np.arange(self.world + 1) * self.batch_global / self.world + 0.5).astype(int)
        self.clo, self.chi = cb[:-1].tolist(), cb[1:].tolist()
        self.cols = [b - a for a, b in zip(self.clo, self.chi)]

    @property
    def my_tables(self):
        return range(self.tlo[self.rank], self.thi[self.rank])

    @property
    def my_rows(self):
        return self.rows[self.rank]

    @property
    def my_cols(self):
        return self.cols[self.rank]

    # element counts of the all-to-all splits
    def fwd_send_splits(self):   # to peer p: my tables' rows x p's columns
        return [self.my_rows * c for c in self.cols]

    def fwd_recv_splits(self):   # from owner q: q's rows x my columns
        return [r * self.my_cols for r in self.rows]

    def bwd_send_splits(self):
        return self.fwd_recv_splits()

    def bwd_recv_splits(self):
        return self.fwd_send_splits()

    def send_block_offset(self, p: int) -> int:
        """element offset of peer p's block inside the forward send buffer"""
        return self.my_rows * self.clo[p]

    def recv_block_offset(self, q: int) -> int:
        return self.my_cols * sum(self.rows[:q])


def exchange(recv: torch.Tensor, send: torch.Tensor, recv_splits, send_splits, group=None):
    """One all-to-all (NCCL on GPUs, gloo in the CPU tests)."""
    dist.all_to_all_single(recv, send, list(recv_splits), list(send_splits), group=group)
    return recv


class ShardedEnsemble:
    """This rank's tables of a table-wise sharded ensemble + the exchange buffers.

    forward(I)       I = indices of MY tables for the GLOBAL batch (list of (bag, B_global) or
                     (B_global,) arrays, or an N-d container) -> my slice of the feature matrix,
                     (prependrows + sum(dims)) x my_cols; rows 1..prependrows are left untouched.
    backward(delta)  delta = cotangent of that matrix -> SparseEmbeddingUpdates of my tables
                     (delta views into the received (my_rows x B_global) buffer + global indices).
    update_(opt, grads)
    """

    def __init__(self, tables_local, plan: ShardPlan, group=None, fused: bool = False, table_groups: int = 1,
                 peer_barrier: bool = True, copy_engine: bool = False):
        """fused=False: NCCL all-to-all + pack/unpack kernels.  fused=True: the lookup kernel stores
        straight into the peers' feature matrices and the backward scatter straight into the owners'
        cotangent buffers over NVLink (CUDA-IPC mapped peer memory); ranks meet at a hand-written
        barrier over peer-memory flags (etb_peer_barrier; peer_barrier=False: a one-element NCCL
        all-reduce instead).  table_groups > 1 (fused only): backward_update_ sends the cotangent
        table group by table group and every owner updates group g while group g+1 is on the wire."""
        self.tables, self.plan, self.group, self.fused = list(tables_local), plan, group, bool(fused)
        self.peer_barrier = bool(peer_barrier) and self.fused
        # copy_engine (fused only): the kernels write LOCAL staging blocks and the GPU's copy engines push them into
        # the peers' buffers over NVLink (etb_memcpy2d_d2d) while the SMs look the next peer's columns up
        self.copy_engine = bool(copy_engine) and self.fused
        assert len(self.tables) == len(plan.my_tables)
        if not self.tables:
            raise ValueError("every rank must own at least one table (fewer tables than ranks)")
        assert [featuresize(t) for t in self.tables] == [plan.dims[t] for t in plan.my_tables]
        self.dtype = self.tables[0].dtype
        p = plan
        self.send = DeviceArray.empty((max(1, p.my_rows * p.batch_global),), self.dtype)
        self.recv = DeviceArray.empty((max(1, (p.total_rows - p.prependrows) * p.my_cols),), self.dtype)
        self.delta_global = DeviceArray.empty((max(1, p.my_rows), p.batch_global), self.dtype)
        self.out = DeviceArray.empty((p.total_rows, p.my_cols), self.dtype)
        self.indexer = Indexer()
        # table groups of the pipelined backward: every owner q cuts ITS tables into the same number of contiguous
        # groups; group_rows[q][g] = (first feature row inside q's block, rows) of q's group g
        G = max(1, min(int(table_groups), min(b - a for a, b in zip(p.tlo, p.thi)))) if self.fused else 1
        self.group_bounds, self.group_rows = [], []
        for q in range(p.world):
            nq = p.thi[q] - p.tlo[q]
            b = [round(g * nq / G) for g in range(G + 1)]
            self.group_bounds.append(b)
            offs = np.concatenate([[0], np.cumsum(p.dims[p.tlo[q]:p.thi[q]])]).astype(int)
            self.group_rows.append([(int(offs[b[g]]), int(offs[b[g + 1]] - offs[b[g]])) for g in range(G)])
        self.n_groups = G
        self.group_indexers = [Indexer() for _ in range(G)] if G > 1 else [self.indexer]
        # an ensemble built with table groups can still be driven whole (one index!, one exchange, one update!): the
        # `grouped` argument of prefetch_index / forward / update_ / backward_update_ chooses per call (bench.py: the
        # device-timed step runs whole, the host-buffer step by groups); None = by groups iff there are groups
        self._index_grouped = G > 1
        self._upd_stream = None
        self._I = None
        self._rows = (C.c_int64 * p.world)(*p.rows)
        self._row_off = (C.c_int64 * p.world)(*p.row_off)
        self.launches = 0
        self.index_launches = 0
        self.update_launches = 0
        if self.fused:
            self._setup_peer_memory()

    # ---- fused mode: peer-mapped buffers -------------------------------------------------
    def _setup_peer_memory(self):
        p, lib, es = self.plan, _lib.lib(), np.dtype(self.dtype).itemsize
        max_cols = max(p.cols)
        sizes = [p.total_rows * max_cols * es, max(1, p.my_rows) * p.batch_global * es, 256]   # out, cotangent, flags
        self._raw = []
        for nbytes in sizes:                               # raw cudaMalloc: an IPC handle maps a whole allocation
            ptr = C.c_void_p()
            _lib.check(lib.etb_malloc(C.byref(ptr), max(nbytes, 256)))
            self._raw.append(ptr.value)
        _lib.check(lib.etb_memset(self._raw[2], 0, 256, C.c_void_p(current_stream_ptr())))   # barrier flags start at 0
        torch.cuda.current_stream().synchronize()
        NB = len(sizes)
        handles = np.zeros(NB * 64, np.uint8)
        for k, ptr in enumerate(self._raw):
            _lib.check(lib.etb_ipc_export(ptr, handles[64 * k:].ctypes.data))
        mine = torch.from_numpy(handles).cuda()
        gathered = [torch.empty_like(mine) for _ in range(p.world)]
        dist.all_gather(gathered, mine, group=self.group)
        self._peer_out, self._peer_dglob, self._peer_flags, self._imported = [], [], [], []
        for q in range(p.world):
            if q == p.rank:
                self._peer_out.append(self._raw[0]); self._peer_dglob.append(self._raw[1]); self._peer_flags.append(self._raw[2])
                continue
            h = gathered[q].cpu().numpy()
            ptrs = []
            for k in range(NB):
                ptr = C.c_void_p()
                _lib.check(lib.etb_ipc_import(np.ascontiguousarray(h[64 * k:64 * k + 64]).ctypes.data, C.byref(ptr)))
                ptrs.append(ptr.value)
                self._imported.append(ptr.value)
            self._peer_out.append(ptrs[0]); self._peer_dglob.append(ptrs[1]); self._peer_flags.append(ptrs[2])
        self._flag_ptrs = (C.c_void_p * p.world)(*self._peer_flags)
        self._epoch = 0
        # my own buffers, as DeviceArrays over the raw allocations
        self.out = DeviceArray.from_pointer(self._raw[0], (p.total_rows, p.my_cols), self.dtype)
        self.delta_global = DeviceArray.from_pointer(self._raw[1], (max(1, p.my_rows), p.batch_global), self.dtype)
        self._peer_out_arrays = [DeviceArray.from_pointer(self._peer_out[q], (p.total_rows, p.cols[q]), self.dtype)
                                 for q in range(p.world)]
        self._flag = torch.zeros(1, device="cuda")
        # owner q's cotangent buffer is rows[q] x B_global; my columns start at clo[rank]
        # Destinations are visited in rotated order (rank+1, rank+2, ..., rank): if every rank started
        # with peer 0 all of them would store into the same GPU at the same time (measured on 8 GPUs:
        # 0.79 ms for the scatter vs 0.25 ms at link rate).
        self._order = [(p.rank + 1 + k) % p.world for k in range(p.world)]
        self._scatter_ptrs = (C.c_void_p * p.world)(*[self._peer_dglob[q] + p.clo[p.rank] * p.rows[q] * es
                                                      for q in self._order])
        self._scatter_rows = (C.c_int64 * p.world)(*[p.rows[q] for q in self._order])
        self._scatter_row_off = (C.c_int64 * p.world)(*[p.row_off[q] for q in self._order])
        # the same per table group: row block (group_rows) of owner q inside its row block, destination = that row
        # block inside q's (rows[q] x B_global) cotangent buffer (leading dimension rows[q])
        self._gscatter = []
        for g in range(self.n_groups):
            ptrs = (C.c_void_p * p.world)(*[self._peer_dglob[q] + (self.group_rows[q][g][0] + p.clo[p.rank] * p.rows[q]) * es
                                            for q in self._order])
            lds = (C.c_int64 * p.world)(*[max(1, p.rows[q]) for q in self._order])
            rows = (C.c_int64 * p.world)(*[self.group_rows[q][g][1] for q in self._order])
            offs = (C.c_int64 * p.world)(*[p.row_off[q] + self.group_rows[q][g][0] for q in self._order])
            self._gscatter.append((ptrs, lds, rows, offs))
        self._upd_stream = torch.cuda.Stream()
        self._copy_streams = [torch.cuda.Stream() for _ in range(2)]
        dist.barrier(group=self.group)

    def _barrier(self):
        """stream-ordered barrier: completes on this rank's stream only after every rank's preceding
        kernels (whose completion makes their peer stores visible) have finished."""
        if self.peer_barrier:
            self._epoch = (self._epoch + 1) & 0xffffffff
            _lib.check(_lib.lib().etb_peer_barrier(self._flag_ptrs, self.plan.rank, self.plan.world, self._epoch,
                                                   C.c_void_p(current_stream_ptr())))
        else:
            dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self.fused:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for ptr in self._imported:
                _lib.lib().etb_ipc_close(ptr)
            for ptr in self._raw:
                _lib.lib().etb_free(ptr)
            self._imported, self._raw, self.fused = [], [], False

    def _forward_copy(self, Is):
        """forward with the copy engines: one lookup launch per peer (all my tables, that peer's columns) into the local
        send layout; as soon as a launch is done its block travels to the peer as ONE strided device-to-device copy,
        beside the next peer's lookup.  My own columns come last and go straight into my feature matrix."""
        p, lib, es = self.plan, _lib.lib(), np.dtype(self.dtype).itemsize
        main = torch.cuda.current_stream()
        send = DeviceArray(self.send.buf, (p.my_rows, p.batch_global), 0, p.my_rows, self.dtype)
        self.launches = 0
        order = self._order[:-1] + [p.rank] if self._order[-1] == p.rank else self._order
        for k, q in enumerate(order):
            if p.cols[q] == 0:
                continue
            items, off = [], 0
            for t, i in zip(self.tables, Is):
                f = featuresize(t)
                if q == p.rank:
                    dst = self.out.rows(p.row_off[p.rank] + off, p.row_off[p.rank] + off + f)
                else:
                    dst = send.rows(off, off + f).cols(p.clo[q], p.chi[q])
                items.append(_item(t, i.cols(p.clo[q], p.chi[q]), dst))
                off += f
            _run(items)
            self.launches += lib.etb_last_launch_count()
            if q != p.rank:
                cs = self._copy_streams[k % len(self._copy_streams)]
                cs.wait_event(main.record_event())
                _lib.check(lib.etb_memcpy2d_d2d(self._peer_out[q] + p.row_off[p.rank] * es, p.total_rows * es,
                                                send.ptr + p.clo[q] * p.my_rows * es, p.my_rows * es, p.my_rows * es, p.cols[q],
                                                C.c_void_p(cs.cuda_stream)))
        for cs in self._copy_streams:
            main.wait_stream(cs)
        self._barrier()
        return self.out

    def _backward_copy(self, delta: DeviceArray):
        """reverse exchange with the copy engines: the row block of every owner goes into that owner's cotangent
        buffer as one strided copy; no SM is involved"""
        p, lib, es = self.plan, _lib.lib(), np.dtype(self.dtype).itemsize
        main = torch.cuda.current_stream()
        ev = main.record_event()
        for k, q in enumerate(self._order):
            if p.rows[q] == 0 or p.my_cols == 0:
                continue
            cs = self._copy_streams[k % len(self._copy_streams)]
            cs.wait_event(ev)
            _lib.check(lib.etb_memcpy2d_d2d(self._peer_dglob[q] + p.clo[p.rank] * p.rows[q] * es, p.rows[q] * es,
                                            delta.ptr + p.row_off[q] * es, delta.ld * es, p.rows[q] * es, p.my_cols,
                                            C.c_void_p(cs.cuda_stream)))
        for cs in self._copy_streams:
            main.wait_stream(cs)
        self._barrier()

    def _forward_fused(self, Is, cols=None):
        """cols = (c0, c1): only local columns c0..c1 of every rank's slice (the e2e path looks the batch up in column
        chunks so that a chunk can travel to the host while the next one is computed)"""
        p = self.plan
        items, off = [], 0
        for t, i in zip(self.tables, Is):
            f = featuresize(t)
            for q in self._order:                         # my rows of peer q's feature matrix, q's columns
                c0, c1 = (0, p.cols[q]) if cols is None else (min(cols[0], p.cols[q]), min(cols[1], p.cols[q]))
                if c1 > c0:
                    dst = self._peer_out_arrays[q].rows(p.row_off[p.rank] + off, p.row_off[p.rank] + off + f).cols(c0, c1)
                    items.append(_item(t, i.cols(p.clo[q] + c0, p.clo[q] + c1), dst))
            off += f
        if items:
            _run(items)
            self.launches = _lib.lib().etb_last_launch_count()
        self._barrier()
        return self.out

    def _backward_fused(self, delta: DeviceArray):
        p = self.plan
        _lib.check(_lib.lib().etb_a2a_scatter(self._scatter_ptrs, delta.ptr, delta.ld, self._scatter_rows,
                                              self._scatter_row_off, p.world, p.my_cols, delta.elt,
                                              C.c_void_p(current_stream_ptr())))
        self._barrier()

    def distribute_indices(self, I_local: DeviceArray, wire_dtype=None):
        """Data-parallel input: `I_local` holds MY samples' indices for ALL tables,
        (bag, my_cols, T_total) [or (my_cols, T_total) for non-reducing lookups].  One all-to-all sends
        each owner the slices of its tables; returns the list forward()/backward() expect: for each of
        MY tables the (bag, B_global) [or (B_global,)] indices of the global batch.
        `wire_dtype=np.int32` halves the bytes on the wire (the kernels take int32 indices natively)."""
        p = self.plan
        reducing = I_local.ndim == 3
        bag = I_local.shape[0] if reducing else 1
        assert I_local.shape[-1] == len(p.dims) and I_local.shape[-2] == p.my_cols and I_local.is_dense
        n_local = int(np.prod(I_local.shape))
        src = I_local.buf[I_local.offset:I_local.offset + n_local]
        if wire_dtype is not None and np.dtype(wire_dtype) != I_local.dtype:
            src = src.to(torch.int32 if np.dtype(wire_dtype) == np.int32 else torch.int64)
        t_mine = len(p.my_tables)
        send_splits = [bag * p.my_cols * (p.thi[q] - p.tlo[q]) for q in range(p.world)]  # last-dim slices are contiguous
        recv_splits = [bag * p.cols[q] * t_mine for q in range(p.world)]
        recv = torch.empty(sum(recv_splits), dtype=src.dtype, device=src.device)
        exchange(recv, src, recv_splits, send_splits, self.group)
        # received block of peer q is [table][q's columns][bag]; tables want [all columns][bag] each
        glob = torch.empty(t_mine * bag * p.batch_global, dtype=src.dtype, device=src.device)
        gv = glob.view(t_mine, p.batch_global * bag)
        off = 0
        for q in range(p.world):
            if p.cols[q]:
                gv[:, bag * p.clo[q]:bag * p.chi[q]] = recv[off:off + recv_splits[q]].view(t_mine, bag * p.cols[q])
            off += recv_splits[q]
        np_dtype = np.dtype(np.int32) if glob.dtype == torch.int32 else np.dtype(np.int64)
        shape = (bag, p.batch_global) if reducing else (p.batch_global,)
        per = bag * p.batch_global
        return [DeviceArray(glob, shape, t * per, None, np_dtype) for t in range(t_mine)]

    def prefetch_index(self, grouped=None):
        """index! of my tables for the indices of the last forward(), on the side stream, starting behind whatever the
        current stream holds so far.  forward() calls it before the lookup (index! then runs beside the lookup, both
        HBM-bound); called after forward(..., prefetch_index=False) it runs beside the backward exchange instead, which
        is NVLink-bound and leaves HBM idle."""
        p = self.plan
        self.index_launches = 0
        self._index_grouped = (self.n_groups > 1) if grouped is None else (bool(grouped) and self.n_groups > 1)
        if not self._index_grouped:
            _prefetch_index(self.indexer, self.tables, self._I)
            self.index_launches = _lib.lib().etb_last_launch_count()
            return
        b = self.group_bounds[p.rank]
        for g, ix in enumerate(self.group_indexers):
            _prefetch_index(ix, self.tables[b[g]:b[g + 1]], self._I[b[g]:b[g + 1]])
            self.index_launches += _lib.lib().etb_last_launch_count()

    def forward(self, I, out: DeviceArray = None, prefetch_index: bool = True, cols=None, grouped=None) -> DeviceArray:
        p = self.plan
        Is = [as_device_indices(i) for i in (I if isinstance(I, (list, tuple)) else
                                             [I.lastdim(t) for t in range(I.shape[-1])])]
        self._I = Is
        if prefetch_index:   # index! needs only the indices: run it beside the lookup and the exchange
            self.prefetch_index(grouped)
        if self.fused:
            assert out is None, "fused mode writes into the peer-mapped self.out"
            if self.copy_engine and cols is None:
                return self._forward_copy(Is)
            return self._forward_fused(Is, cols)
        assert cols is None
        out = self.out if out is None else out
        # Peer p's send block is (my_rows x cols[p]) dense at element offset my_rows*clo[p]: the
        # blocks laid end to end ARE the column-major (my_rows x B_global) matrix of my tables'
        # lookups, so one item per table writes the whole send layout.
        send = DeviceArray(self.send.buf, (p.my_rows, p.batch_global), 0, p.my_rows, self.dtype)
        items, off = [], 0
        for t, i in zip(self.tables, Is):
            f = featuresize(t)
            if p.batch_global > 0:
                items.append(_item(t, i, send.rows(off, off + f)))
            off += f
        if items:
            _run(items)                                        # K6: lookups written in send layout
            self.launches = _lib.lib().etb_last_launch_count()
        n_send, n_recv = p.my_rows * p.batch_global, (p.total_rows - p.prependrows) * p.my_cols
        exchange(self.recv.buf[:n_recv], self.send.buf[:n_send], p.fwd_recv_splits(), p.fwd_send_splits(), self.group)
        _lib.check(_lib.lib().etb_a2a_unpack(out.ptr, out.ld, self.recv.ptr, self._rows, self._row_off, p.world,
                                             p.my_cols, out.elt, C.c_void_p(current_stream_ptr())))
        self.launches += 1
        return out

    def backward(self, delta: DeviceArray):
        p = self.plan
        if self.fused:
            if self.copy_engine:
                self._backward_copy(delta)
            else:
                self._backward_fused(delta)
            return self._grads()
        n_send, n_recv = (p.total_rows - p.prependrows) * p.my_cols, p.my_rows * p.batch_global
        # pack my cotangent's row blocks by owner (reuses the forward receive buffer)
        _lib.check(_lib.lib().etb_a2a_pack(self.recv.ptr, delta.ptr, delta.ld, self._rows, self._row_off, p.world,
                                           p.my_cols, delta.elt, C.c_void_p(current_stream_ptr())))
        exchange(self.delta_global.buf[:n_recv], self.recv.buf[:n_send], p.bwd_recv_splits(), p.bwd_send_splits(),
                 self.group)
        return self._grads()

    def _grads(self):
        grads, off = [], 0
        for t, i in zip(self.tables, self._I):
            f = featuresize(t)
            grads.append(SparseEmbeddingUpdate(t.lookup_type, self.delta_global.rows(off, off + f), i))
            off += f
        return grads

    def update_(self, opt, grads, grouped=None):
        if grouped is None:
            grouped = self._index_grouped    # whichever Indexers the last prefetch filled (update_ re-indexes otherwise)
        if self.n_groups == 1 or not grouped:
            update_(opt, self.tables, grads, [self.indexer])
            return
        b = self.group_bounds[self.plan.rank]
        for g, ix in enumerate(self.group_indexers):
            update_(opt, self.tables[b[g]:b[g + 1]], grads[b[g]:b[g + 1]], [ix])

    def scatter_group(self, delta: DeviceArray, g: int):
        """backward exchange of table group g alone: my cotangent's rows of every owner's group g go into that
        owner's buffer (fused mode), followed by the barrier"""
        p = self.plan
        ptrs, lds, rows, offs = self._gscatter[g]
        _lib.check(_lib.lib().etb_a2a_scatter_ld(ptrs, lds, delta.ptr, delta.ld, rows, offs, p.world, p.my_cols, delta.elt,
                                                 C.c_void_p(current_stream_ptr())))
        self._barrier()

    def update_group_(self, opt, g: int):
        """update! of my tables of group g on the update stream, as soon as the main stream's barrier for that
        group has passed; call join_updates() before the tables are used again"""
        b = self.group_bounds[self.plan.rank]
        grads = self._grads()[b[g]:b[g + 1]]
        ev = torch.cuda.current_stream().record_event()
        self._upd_stream.wait_event(ev)
        with torch.cuda.stream(self._upd_stream):
            update_(opt, self.tables[b[g]:b[g + 1]], grads, [self.group_indexers[g]])
        self.update_launches += _lib.lib().etb_last_launch_count()

    def join_updates(self):
        torch.cuda.current_stream().wait_stream(self._upd_stream)

    def backward_update_(self, opt, delta: DeviceArray, grouped=None):
        """backward + update!, pipelined over table groups (fused mode): while the owners update group g the
        cotangent of group g+1 crosses NVLink.  Same results as backward() followed by update_()."""
        self.update_launches = 0
        if grouped is None:
            grouped = self._index_grouped
        if not self.fused or self.n_groups == 1 or not grouped:
            self.update_(opt, self.backward(delta), grouped=False)
            self.update_launches = _lib.lib().etb_last_launch_count()
            return
        for g in range(self.n_groups):
            self.scatter_group(delta, g)
            self.update_group_(opt, g)
        self.join_updates()
