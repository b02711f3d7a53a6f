"""Graph build behind the reference's dataset API (recbole_gnn/data/dataset.py:24-106):
``GeneralGraphDataset.get_norm_adj_mat`` / ``edge_index_to_adj_t`` / ``get_bipartite_inter_mat``.

The three methods live in :class:`GraphBuildMixin`, which reads the same attributes the reference reads
from RecBole's ``Dataset`` (``inter_feat``, ``uid_field``, ``iid_field``, ``user_num``, ``item_num``,
``num(field)``).  ``GeneralGraphDataset`` is the mixin on top of RecBole's ``Dataset`` when ``recbole`` is
importable; :class:`InteractionDataset` is a RecBole-free carrier of the same attributes (synthetic graphs,
tests, bench).

Where the reference builds and normalises on the CPU (PyG ``gcn_norm`` over int64 COO), this engine
defers to the GPU: ``enable_sparse=True`` returns a *described* :class:`GraphHandle` that becomes a CSR
when ``GeneralGraphRecommender`` calls ``.to(device)``; the dense-edge return (``enable_sparse`` falsy)
is computed by the device kernels and returned as CUDA tensors.  No CPU arithmetic path exists.
"""
from __future__ import annotations

import torch

from . import _lib
from .graph import GraphHandle, gcn_norm

try:  # pragma: no cover - recbole is absent in the build container
    from recbole.data.dataset import Dataset as _RecBoleDataset
    HAVE_RECBOLE = True
except Exception:  # ImportError or its transitive failures
    _RecBoleDataset = object
    HAVE_RECBOLE = False

is_sparse = True  # the engine's sparse object is always available (reference: torch_sparse importable)


def _device_of(ds) -> torch.device:
    dev = None
    cfg = getattr(ds, "config", None)
    if cfg is not None:
        try:
            dev = cfg["device"]
        except Exception:
            dev = None
    dev = torch.device(dev) if dev is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise RuntimeError("recbole_gnn_b200 builds graphs on a CUDA device only (config['device'] is %s)" % dev)
    if not torch.cuda.is_available():
        raise RuntimeError("recbole_gnn_b200: no CUDA device available and no CPU path exists")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


class GraphBuildMixin:
    @staticmethod
    def edge_index_to_adj_t(edge_index, edge_weight, m_num_nodes, n_num_nodes):
        """dataset.py:41-47: ``SparseTensor(row=edge_index[0], col=edge_index[1], value, (m, n)).t()``."""
        adj = GraphHandle(row=edge_index[0], col=edge_index[1], value=edge_weight,
                          sparse_sizes=(m_num_nodes, n_num_nodes))
        return adj.t()

    def get_norm_adj_mat(self, enable_sparse=False):
        r"""dataset.py:49-79: :math:`\hat A = D^{-1/2} A D^{-1/2}` of the symmetric user-item graph.

        ``enable_sparse`` truthy -> ``(GraphHandle, None)``; otherwise ``(edge_index, edge_weight)`` in the
        reference's edge order ([user->item block | item->user block]), as CUDA tensors."""
        self.is_sparse = is_sparse
        row = self.inter_feat[self.uid_field]
        col = self.inter_feat[self.iid_field]
        num_nodes = self.user_num + self.item_num
        if enable_sparse:
            adj_t = GraphHandle.from_interactions(row, col, self.user_num, self.item_num)
            adj_t = gcn_norm(adj_t, None, num_nodes, add_self_loops=False)
            return adj_t, None
        dev = _device_of(self)
        row = row.to(dev)
        col = col.to(dev) + self.user_num
        edge_index1 = torch.stack([row, col])
        edge_index2 = torch.stack([col, row])
        edge_index = torch.cat([edge_index1, edge_index2], dim=1)
        edge_index, edge_weight = gcn_norm(edge_index, None, num_nodes, add_self_loops=False)
        return edge_index, edge_weight

    def get_bipartite_inter_mat(self, row='user', row_norm=True):
        """dataset.py:81-106: rectangular COO with ``1/deg_row`` or ``deg_row^-1/2 deg_col^-1/2`` weights."""
        if row == 'user':
            row_field, col_field = self.uid_field, self.iid_field
        else:
            row_field, col_field = self.iid_field, self.uid_field
        dev = _device_of(self)
        r = self.inter_feat[row_field].to(dev).contiguous()
        c = self.inter_feat[col_field].to(dev).contiguous()
        edge_index = torch.stack([r, c])
        n_row, n_col = self.num(row_field), self.num(col_field)
        w = torch.empty(r.numel(), dtype=torch.float32, device=dev)
        ws = torch.empty(max((n_row + n_col) * 4, 4), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200gcn_bipartite_norm_coo(
                r.data_ptr(), c.data_ptr(), r.numel(), n_row, n_col, int(bool(row_norm)), w.data_ptr(),
                ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
        return edge_index, w


class InteractionDataset(GraphBuildMixin):
    """RecBole-free carrier of the attributes the graph build reads.  ``uid``/``iid`` are the remapped
    int64 ``inter_feat`` columns (0 = [PAD]); ``user_num``/``item_num`` include the [PAD] id."""

    def __init__(self, uid: torch.Tensor, iid: torch.Tensor, user_num: int, item_num: int, device=None):
        self.uid_field, self.iid_field = "user_id", "item_id"
        self.inter_feat = {self.uid_field: uid, self.iid_field: iid}
        self.user_num, self.item_num = int(user_num), int(item_num)
        self.config = {"device": device} if device is not None else None

    def num(self, field):
        return self.user_num if field == self.uid_field else self.item_num

    @classmethod
    def from_inter_file(cls, path: str, device=None) -> "InteractionDataset":
        """Read a RecBole atomic ``.inter`` file (tab-separated, header ``user_id:token\titem_id:token...``, the
        format of the reference's ``tests/test_data/test/test.inter``) and remap the tokens to ids in
        first-appearance order starting at 1 — id 0 is RecBole's ``[PAD]`` — which is what RecBole's ``Dataset``
        hands to ``get_norm_adj_mat`` through ``inter_feat`` (dataset.py:60-61).  Parsed by the library's native
        reader (``b200gcn_inter_open``: mmap + one tokenising pass), not by a Python line loop."""
        import ctypes as C
        lib = _lib.load()
        handle, n, un, inn = C.c_void_p(), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _lib.check(lib.b200gcn_inter_open(str(path).encode(), C.byref(handle), C.byref(n), C.byref(un), C.byref(inn)))
        try:
            uid = torch.empty(n.value, dtype=torch.int64)
            iid = torch.empty(n.value, dtype=torch.int64)
            if n.value > 0:
                _lib.check(lib.b200gcn_inter_read(handle, uid.data_ptr(), iid.data_ptr()))
        finally:
            lib.b200gcn_inter_close(handle)
        return cls(uid, iid, un.value, inn.value, device=device)


if HAVE_RECBOLE:  # pragma: no cover
    class GeneralGraphDataset(GraphBuildMixin, _RecBoleDataset):
        def __init__(self, config):
            super().__init__(config)
else:
    class GeneralGraphDataset(InteractionDataset):
        """Without RecBole the class degrades to :class:`InteractionDataset` (same graph-build methods)."""
