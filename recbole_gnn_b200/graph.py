"""``GraphHandle`` — the engine's sparse-matrix object, a stand-in for ``torch_sparse.SparseTensor`` on
the reference's ``enable_sparse`` path (recbole_gnn/data/dataset.py:41-47,68-75;
recbole_gnn/model/abstract_recommender.py:15-20; ngcf.py:79-87; sgl.py:120-122).

A handle describes a matrix ``A`` of shape ``sparse_sizes = (n_rows, n_cols)``; propagation computes
``out[r] = sum_c A[r, c] * x[c]`` exactly like ``torch_sparse.matmul(A, x, reduce='add')``.

Two states:

* *described* (any device, picklable): COO ids (or the raw ``inter_feat`` columns of a bipartite
  graph) plus a list of pending operations (``t()``, ``gcn_norm``).  Nothing is computed — the
  reference builds and normalises on the CPU, this engine defers both to the GPU.
* *resident* (CUDA): CSR keyed by row (int64 rowptr, int32 col, fp32 val or None for unit weights), a
  hub-row plan for skewed graphs and a lazily built transpose for the backward product.

``.to(cuda_device)`` turns the first into the second (H2D copy of the ids, radix-sort CSR build and
normalisation on the device through libb200gcn).  There is no CPU compute path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor

LONG_ROW = 4096      # rows with more entries than this take the hub path
HUB_CHUNK = 8192     # entries per CTA on the hub path (a 9 M-entry row becomes ~1150 chunks)
_HUB_CAP = 1 << 16


def _as_device(device) -> torch.device:
    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


class GraphHandle:
    # ------------------------------------------------------------------ construction
    def __init__(self, row: Optional[Tensor] = None, col: Optional[Tensor] = None,
                 value: Optional[Tensor] = None, sparse_sizes: Optional[Tuple[int, int]] = None, *,
                 rowptr: Optional[Tensor] = None, symmetric: bool = False):
        """``GraphHandle(row=, col=, value=, sparse_sizes=(m, n))`` mirrors the ``SparseTensor``
        constructor used at dataset.py:43-46 and ngcf.py:84-86: entry k is ``A[row[k], col[k]] = value[k]``
        (value None = ones); duplicates are kept as parallel entries."""
        if sparse_sizes is None:
            raise ValueError("sparse_sizes=(n_rows, n_cols) is required")
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        self._symmetric = bool(symmetric)
        self._ops: List[str] = []          # pending ops in the described state
        self._inter = None                 # (uid, iid, user_num, item_num) for bipartite descriptions
        self._t_cache: Optional["GraphHandle"] = None
        self._hubs: Optional[Tensor] = None
        self._n_hubs = 0
        self._long_row = LONG_ROW
        if rowptr is not None:             # resident CSR handed in directly
            _lib.require_cuda(rowptr, col, value, what="CSR array")
            self._rowptr, self._col, self._val = rowptr, col, value
            self._row = None
            self._resident = True
            self._plan_hubs()
        else:
            if row is None or col is None:
                raise ValueError("row and col are required")
            if row.dtype != torch.int64 or col.dtype != torch.int64:
                raise TypeError("row/col must be int64 (the reference's edge_index dtype)")
            if row.dim() != 1 or row.shape != col.shape:
                raise ValueError("row/col must be 1-D tensors of equal length")
            if value is not None and (value.dtype != torch.float32 or value.shape != row.shape):
                raise TypeError("value must be float32 with one entry per edge")
            self._row, self._col, self._val = row, col, value
            self._rowptr = None
            self._resident = False

    @classmethod
    def from_interactions(cls, uid: Tensor, iid: Tensor, user_num: int, item_num: int) -> "GraphHandle":
        """Description of the symmetric ``[[0, R], [R^T, 0]]`` adjacency of
        ``get_norm_adj_mat`` (dataset.py:60-66) straight from the ``inter_feat`` columns; the
        int64 ``[2, 2E]`` edge_index is never materialised."""
        if uid.dtype != torch.int64 or iid.dtype != torch.int64 or uid.shape != iid.shape or uid.dim() != 1:
            raise TypeError("uid/iid must be 1-D int64 tensors of equal length")
        n = int(user_num) + int(item_num)
        h = cls.__new__(cls)
        h._sizes, h._symmetric, h._ops = (n, n), True, []
        h._inter = (uid, iid, int(user_num), int(item_num))
        h._row = h._col = h._val = h._rowptr = None
        h._resident = False
        h._t_cache, h._hubs, h._n_hubs, h._long_row = None, None, 0, LONG_ROW
        return h

    # ------------------------------------------------------------------ SparseTensor surface
    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sizes

    def size(self, dim: int) -> int:
        return self._sizes[dim]

    def nnz(self) -> int:
        if self._inter is not None and not self._resident:
            return 2 * self._inter[0].numel()
        return int(self._col.numel())

    @property
    def is_resident(self) -> bool:
        return self._resident

    @property
    def is_symmetric(self) -> bool:
        return self._symmetric

    @property
    def device(self) -> torch.device:
        if self._resident:
            return self._rowptr.device
        return (self._inter[0] if self._inter is not None else self._row).device

    @property
    def is_cuda(self) -> bool:
        return self.device.type == "cuda"

    def t(self) -> "GraphHandle":
        """Transpose (dataset.py:47, ngcf.py:79,87)."""
        if self._symmetric:
            return self
        if self._resident:
            t = self._t_cache() if isinstance(self._t_cache, weakref.ref) else self._t_cache
            if t is None:
                t = self._transpose_resident()
                self._t_cache = t
                t._t_cache = weakref.ref(self)      # back-pointer is weak: no reference cycle, freed by refcount
            return t
        h = self._clone_description()
        if not h._ops:   # nothing order-dependent recorded yet: swap the COO roles for free
            h._row, h._col = h._col, h._row
            h._sizes = (h._sizes[1], h._sizes[0])
        else:
            h._ops.append("t")
            h._sizes = (h._sizes[1], h._sizes[0])
        return h

    def gcn_norm(self) -> "GraphHandle":
        """``gcn_norm(adj_t, None, N, add_self_loops=False)`` on a square handle (dataset.py:74, sgl.py:121)."""
        if self._sizes[0] != self._sizes[1]:
            raise ValueError("gcn_norm needs a square matrix")
        if self._resident:
            val = torch.empty(self.nnz(), dtype=torch.float32, device=self.device)
            lib = _lib.load()
            with torch.cuda.device(self.device):
                _lib.check(lib.b200gcn_gcn_norm_csr(_lib.ptr(self._rowptr), _lib.ptr(self._col), _lib.ptr(self._val),
                                                    _lib.ptr(val), None, self._sizes[0], _lib.stream_ptr(self.device)))
            h = GraphHandle(rowptr=self._rowptr, col=self._col, value=val, sparse_sizes=self._sizes,
                            symmetric=self._symmetric)
            return h
        h = self._clone_description()
        h._ops.append("gcn_norm")
        return h

    def to(self, device, *args, **kwargs) -> "GraphHandle":
        """``adj_t.to(device)`` (abstract_recommender.py:18, sgl.py:122).  Moving to a CUDA device makes
        the handle resident: ids are copied, the CSR is built and normalised on the device."""
        d = _as_device(device)
        if d.type == "cpu":
            if self._resident:
                raise RuntimeError("a resident GraphHandle cannot move to the CPU; use .coo() to export it")
            return self
        if d.type != "cuda":
            raise RuntimeError(f"GraphHandle supports CUDA devices only, got {d}")
        if self._resident:
            if self.device == d:
                return self
            h = GraphHandle(rowptr=self._rowptr.to(d), col=self._col.to(d),
                            value=None if self._val is None else self._val.to(d),
                            sparse_sizes=self._sizes, symmetric=self._symmetric)
            return h
        return self._materialise(d)

    def cuda(self, device=None) -> "GraphHandle":
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def coo(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        """``(row, col, value)`` sorted by (row, col) like ``SparseTensor.coo()`` (ngcf.py:79)."""
        self._need_resident("coo()")
        nnz = self.nnz()
        row = torch.empty(nnz, dtype=torch.int64, device=self.device)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.b200gcn_csr_row_ids(_lib.ptr(self._rowptr), self._sizes[0], nnz, _lib.ptr(row),
                                               _lib.stream_ptr(self.device)))
        return row, self._col.to(torch.int64), self._val

    def csr(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        self._need_resident("csr()")
        return self._rowptr, self._col, self._val

    def masked(self, keep: Tensor, symmetric: bool = False) -> "GraphHandle":
        """Edge dropout on the resident CSR without a re-sort: keeps entry e (CSR order == ``coo()``
        order) iff ``keep[e]``; PyG ``dropout_adj`` semantics, no rescale (ngcf.py:81,89).  ``symmetric=True``
        declares that the mask keeps (r, c) and (c, r) together (node dropout), so the result is its own transpose."""
        self._need_resident("masked()")
        _lib.require_cuda(keep, what="keep mask")
        nnz, n = self.nnz(), self._sizes[0]
        if keep.numel() != nnz:
            raise ValueError("keep must have one flag per entry")
        keep8 = keep.to(torch.uint8).contiguous()
        lib = _lib.load()
        dev = self.device
        with torch.cuda.device(dev):
            need = C.c_size_t(0)
            _lib.check(lib.b200gcn_csr_mask_workspace(nnz, n, C.byref(need)))
            ws = _workspace(need.value, dev)
            rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
            col = torch.empty(nnz, dtype=torch.int32, device=dev)
            val = None if self._val is None else torch.empty(nnz, dtype=torch.float32, device=dev)
            kept = C.c_int64(0)
            _lib.check(lib.b200gcn_csr_mask(_lib.ptr(self._rowptr), _lib.ptr(self._col), _lib.ptr(self._val),
                                            _lib.ptr(keep8), n, nnz, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val),
                                            C.byref(kept), _lib.ptr(ws), need.value, _lib.stream_ptr(dev)))
        k = kept.value
        return GraphHandle(rowptr=rowptr, col=col[:k], value=None if val is None else val[:k],
                           sparse_sizes=self._sizes, symmetric=bool(symmetric) and self._symmetric)

    def __repr__(self) -> str:
        state = "resident" if self._resident else "described"
        return f"GraphHandle({state}, sizes={self._sizes}, nnz={self.nnz()}, device={self.device})"

    # pickling (the reference pickles its dataset, dataset.py:30-39): only descriptions travel
    def __getstate__(self):
        if self._resident:
            raise RuntimeError("a resident GraphHandle holds device state and is not picklable; "
                               "keep resident handles on the model, not on the dataset")
        return self.__dict__

    # ------------------------------------------------------------------ internals
    def _need_resident(self, what: str) -> None:
        if not self._resident:
            raise RuntimeError(f"GraphHandle.{what} needs a resident handle: call .to('cuda') first "
                               "(the engine has no CPU path)")

    def _clone_description(self) -> "GraphHandle":
        h = GraphHandle.__new__(GraphHandle)
        h.__dict__.update(self.__dict__)
        h._ops = list(self._ops)
        h._t_cache = None
        return h

    def _plan_hubs(self) -> None:
        lib = _lib.load()
        dev = self.device
        n = self._sizes[0]
        with torch.cuda.device(dev):
            long_row = LONG_ROW
            while True:
                hubs = torch.empty(_HUB_CAP, dtype=torch.int64, device=dev)
                cnt = C.c_int32(0)
                _lib.check(lib.b200gcn_plan_hubs(_lib.ptr(self._rowptr), n, long_row, _lib.ptr(hubs), _HUB_CAP,
                                                 C.byref(cnt), _lib.stream_ptr(dev)))
                if cnt.value <= _HUB_CAP:
                    break
                long_row *= 2
        self._long_row, self._n_hubs = long_row, cnt.value
        self._hubs = hubs[:cnt.value].sort().values if cnt.value > 0 else None
        self._n_chunks, self._hub_chunk_ptr, self._chunk_beg, self._chunk_end = 0, None, None, None
        self._hub_scratch = {}
        if cnt.value > 0:
            # chunk plan (device-side index arithmetic; one small sync for the chunk count)
            beg = self._rowptr[self._hubs]
            end = self._rowptr[self._hubs + 1]
            n_ch = (end - beg + HUB_CHUNK - 1) // HUB_CHUNK
            ptr = torch.zeros(cnt.value + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(n_ch, 0)
            total = int(ptr[-1].item())
            hub_of_chunk = torch.repeat_interleave(torch.arange(cnt.value, device=dev), n_ch)
            k = torch.arange(total, device=dev) - ptr[hub_of_chunk]
            self._chunk_beg = (beg[hub_of_chunk] + k * HUB_CHUNK).contiguous()
            self._chunk_end = torch.minimum(self._chunk_beg + HUB_CHUNK, end[hub_of_chunk]).contiguous()
            self._hub_chunk_ptr = ptr.to(torch.int32).contiguous()
            self._n_chunks = total

    def _hub_plan(self, dim: int):
        """ctypes ``b200gcn_hub_plan`` for this graph with a [n_chunks, dim] scratch (cached per dim)."""
        sc = self._hub_scratch.get(dim)
        if sc is None:
            sc = torch.empty(self._n_chunks, dim, dtype=torch.float32, device=self.device)
            self._hub_scratch[dim] = sc
        hp = _lib.HubPlan()
        hp.n_hubs, hp.n_chunks = self._n_hubs, self._n_chunks
        hp.hub_rows, hp.hub_chunk_ptr = self._hubs.data_ptr(), self._hub_chunk_ptr.data_ptr()
        hp.chunk_beg, hp.chunk_end, hp.scratch = self._chunk_beg.data_ptr(), self._chunk_end.data_ptr(), sc.data_ptr()
        return hp

    def _transpose_resident(self) -> "GraphHandle":
        lib = _lib.load()
        dev = self.device
        n_rows, n_cols = self._sizes
        nnz = self.nnz()
        with torch.cuda.device(dev):
            need = C.c_size_t(0)
            _lib.check(lib.b200gcn_csr_transpose_workspace(nnz, n_rows, n_cols, C.byref(need)))
            ws = _workspace(need.value, dev)
            rowptr_t = torch.empty(n_cols + 1, dtype=torch.int64, device=dev)
            col_t = torch.empty(nnz, dtype=torch.int32, device=dev)
            val_t = None if self._val is None else torch.empty(nnz, dtype=torch.float32, device=dev)
            _lib.check(lib.b200gcn_csr_transpose(_lib.ptr(self._rowptr), _lib.ptr(self._col), _lib.ptr(self._val),
                                                 n_rows, n_cols, nnz, _lib.ptr(rowptr_t), _lib.ptr(col_t),
                                                 _lib.ptr(val_t), _lib.ptr(ws), need.value, _lib.stream_ptr(dev)))
        return GraphHandle(rowptr=rowptr_t, col=col_t, value=val_t, sparse_sizes=(n_cols, n_rows))

    def _materialise(self, dev: torch.device, check: bool = True) -> "GraphHandle":
        lib = _lib.load()
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            need = C.c_size_t(0)
            if self._inter is not None:
                uid, iid, U, I = self._inter
                uid_d, iid_d = uid.to(dev, non_blocking=True), iid.to(dev, non_blocking=True)
                E, n = uid.numel(), U + I
                _lib.check(lib.b200gcn_csr_from_interactions_workspace(E, U, I, C.byref(need)))
                ws = _workspace(need.value, dev)
                rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
                col = torch.empty(2 * E, dtype=torch.int32, device=dev)
                _lib.check(lib.b200gcn_csr_from_interactions(_lib.ptr(uid_d), _lib.ptr(iid_d), E, U, I,
                                                             _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(ws),
                                                             need.value, int(check), st))
                val = None
                sizes = (n, n)
            else:
                n_rows, n_cols = self._sizes
                # undo the size swaps of recorded "t" ops to get the sizes of the COO as described
                for op in self._ops:
                    if op == "t":
                        n_rows, n_cols = n_cols, n_rows
                row_d, col_d = self._row.to(dev, non_blocking=True), self._col.to(dev, non_blocking=True)
                w_d = None if self._val is None else self._val.to(dev, non_blocking=True)
                nnz = row_d.numel()
                _lib.check(lib.b200gcn_csr_from_coo_workspace(nnz, n_rows, n_cols, C.byref(need)))
                ws = _workspace(need.value, dev)
                rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
                col = torch.empty(nnz, dtype=torch.int32, device=dev)
                val = None if w_d is None else torch.empty(nnz, dtype=torch.float32, device=dev)
                # matrix rows are propagation destinations, matrix cols are sources
                _lib.check(lib.b200gcn_csr_from_coo(_lib.ptr(col_d), _lib.ptr(row_d), _lib.ptr(w_d), nnz, n_rows, n_cols,
                                                    _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), None,
                                                    _lib.ptr(ws), need.value, int(check), st))
                sizes = (n_rows, n_cols)
            del ws
            h = GraphHandle(rowptr=rowptr, col=col, value=val, sparse_sizes=sizes, symmetric=self._symmetric)
            for op in self._ops:
                h = h.gcn_norm() if op == "gcn_norm" else h.t()
        return h


# the reference's name for the same role
SparseTensor = GraphHandle


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=False,
             flow="source_to_target", dtype=None):
    """Drop-in for PyG ``gcn_norm`` as the reference calls it (dataset.py:74,77; sgl.py:121,124):
    ``add_self_loops=False`` only.  A GraphHandle gets the normalisation recorded (or applied, when
    resident); a CUDA ``edge_index`` gets ``(edge_index, dis[row]*w*dis[col])`` computed on the device."""
    if add_self_loops or improved or flow != "source_to_target":
        raise NotImplementedError("only gcn_norm(add_self_loops=False, flow='source_to_target') is on this path")
    if isinstance(edge_index, GraphHandle):
        return edge_index.gcn_norm()
    _lib.require_cuda(edge_index, edge_weight, what="edge_index/edge_weight")
    if num_nodes is None:
        raise ValueError("num_nodes is required")
    lib = _lib.load()
    dev = edge_index.device
    nnz = edge_index.size(1)
    src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
    w = None if edge_weight is None else edge_weight.contiguous().float()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        need = C.c_size_t(0)
        _lib.check(lib.b200gcn_csr_from_coo_workspace(nnz, num_nodes, num_nodes, C.byref(need)))
        ws = _workspace(need.value, dev)
        rowptr = torch.empty(num_nodes + 1, dtype=torch.int64, device=dev)
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = None if w is None else torch.empty(nnz, dtype=torch.float32, device=dev)
        perm = torch.empty(nnz, dtype=torch.int64, device=dev)
        _lib.check(lib.b200gcn_csr_from_coo(_lib.ptr(src), _lib.ptr(dst), _lib.ptr(w), nnz, num_nodes, num_nodes,
                                            _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(perm),
                                            _lib.ptr(ws), need.value, 1, st))
        out = torch.empty(nnz, dtype=torch.float32, device=dev)
        _lib.check(lib.b200gcn_gcn_norm_csr(_lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(out), None,
                                            num_nodes, st))
    w_coo = torch.empty_like(out)
    w_coo[perm] = out      # back to the caller's edge order
    return edge_index, w_coo
