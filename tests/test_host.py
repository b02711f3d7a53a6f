"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import pickle
import re
import subprocess

import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(b200gcn_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    # the ctypes table binds exactly the header's symbols
    assert sorted(_lib.SIGNATURES) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (b200gcn_[a-z0-9_]+)", out))
    assert exported == set(declared)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_spmm_args_struct_layout(tmp_path):
    """The ctypes mirrors must have the C layout of include/b200gcn.h: compile the header and compare every offset."""
    a = _lib.SpmmArgs
    assert a.n_rows.offset == 0 and a.dim.offset == 8 and a.rowptr.offset == 16
    assert a.eps.offset + 4 == a.acc_scale.offset and a.seed.offset == a.acc_scale.offset + 4
    fields = [f[0] for f in a._fields_]
    hub_fields = [f[0] for f in _lib.HubPlan._fields_]
    src = tmp_path / "off.c"
    body = "".join(f'printf("{f} %zu\\n", offsetof(b200gcn_spmm_args, {f}));\n' for f in fields)
    body += "".join(f'printf("hub.{f} %zu\\n", offsetof(b200gcn_hub_plan, {f}));\n' for f in hub_fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200gcn.h"\nint main(void){\n' + body +
                   'printf("sizeof %zu %zu\\n", sizeof(b200gcn_spmm_args), sizeof(b200gcn_hub_plan));return 0;}\n')
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = dict(l.rsplit(" ", 1) for l in out if l and not l.startswith("sizeof"))
    for f in fields:
        assert int(got[f]) == getattr(a, f).offset, f
    for f in hub_fields:
        assert int(got["hub." + f]) == getattr(_lib.HubPlan, f).offset, f
    sizes = [l for l in out if l.startswith("sizeof")][0].split()
    assert int(sizes[1]) == ctypes.sizeof(a) and int(sizes[2]) == ctypes.sizeof(_lib.HubPlan)


def test_error_reporting_without_gpu():
    lib = _lib.load()
    need = ctypes.c_size_t(0)
    rc = lib.b200gcn_csr_from_coo_workspace(-1, 4, 4, ctypes.byref(need))
    assert rc == _lib.ERR_INVALID
    assert b"nnz" in lib.b200gcn_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = lib.b200gcn_spmm(None, None)
    assert rc == _lib.ERR_INVALID


def test_no_cpu_fallback():
    x = torch.zeros(5, 8)
    ds = rg.InteractionDataset(torch.tensor([1, 2]), torch.tensor([1, 1]), 3, 2)
    h, w = ds.get_norm_adj_mat(enable_sparse=True)
    assert w is None and not h.is_resident and h.sparse_sizes() == (5, 5) and h.nnz() == 4
    with pytest.raises(RuntimeError):
        rg.LightGCNConv(8)(x, h, None)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        rg.LightGCNConv(8)(x, ei, torch.ones(2))
    with pytest.raises(RuntimeError, match="CUDA"):
        rg.functional.lightgcn_propagate(h, x[:3], x[3:], 2)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ds.get_norm_adj_mat(enable_sparse=False)
        with pytest.raises(RuntimeError):
            ds.get_bipartite_inter_mat()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "recbole_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_handle_description_algebra():
    row = torch.tensor([0, 1, 2, 2]); col = torch.tensor([1, 0, 0, 3])
    h = rg.GraphHandle(row=row, col=col, value=None, sparse_sizes=(3, 4))
    t = h.t()
    assert t.sparse_sizes() == (4, 3) and torch.equal(t._row, col) and torch.equal(t._col, row)
    assert t.t().sparse_sizes() == (3, 4)
    with pytest.raises(ValueError):
        h.gcn_norm()
    sq = rg.GraphHandle(row=row, col=col, sparse_sizes=(4, 4))
    n = rg.gcn_norm(sq, None, 4, add_self_loops=False)
    assert n._ops == ["gcn_norm"] and sq._ops == []
    assert n.t()._ops == ["gcn_norm", "t"]
    with pytest.raises(NotImplementedError):
        rg.gcn_norm(sq, None, 4, add_self_loops=True)
    with pytest.raises(RuntimeError):
        h.coo()
    with pytest.raises(TypeError):
        rg.GraphHandle(row=row.int(), col=col, sparse_sizes=(3, 4))
    # edge_index_to_adj_t mirrors SparseTensor(row=ei[0], col=ei[1]).t(): rows become destinations
    adj_t = rg.GeneralGraphDataset.edge_index_to_adj_t(torch.stack([row, col]), None, 3, 4)
    assert adj_t.sparse_sizes() == (4, 3)
    # descriptions are picklable (the reference pickles its dataset), and .to('cpu') is the identity
    h2 = pickle.loads(pickle.dumps(n))
    assert h2._ops == ["gcn_norm"] and h2.to("cpu") is h2
    sym = rg.GraphHandle.from_interactions(torch.tensor([1]), torch.tensor([1]), 2, 2)
    assert sym.t() is sym and sym.nnz() == 2


def test_layer_reprs_and_state_dict_keys():
    assert repr(rg.LightGCNConv(64)) == "LightGCNConv(64)"
    assert repr(rg.BipartiteGCNConv(32)) == "BipartiteGCNConv(32)"
    m = rg.BiGNNConv(64, 32)
    assert repr(m) == "BiGNNConv(64,32)"
    assert sorted(m.state_dict()) == ["lin1.bias", "lin1.weight", "lin2.bias", "lin2.weight"]
    assert list(rg.LightGCNConv(8).state_dict()) == []


def test_inter_file_ingest(tmp_path, g1):
    """`.inter` atomic file -> remapped ids (0 = [PAD]); the reference fixture's ids are what the goldens hold."""
    p = tmp_path / "toy.inter"
    p.write_text("user_id:token\titem_id:token\trating:float\ttimestamp:float\n"
                 "196\t242\t3\t881250949\n186\t302\t3\t891717742\n196\t302\t1\t1\n22\t377\t1\t878887116\n")
    ds = rg.InteractionDataset.from_inter_file(str(p))
    assert ds.user_num == 4 and ds.item_num == 4
    assert ds.inter_feat["user_id"].tolist() == [1, 2, 1, 3]
    assert ds.inter_feat["item_id"].tolist() == [1, 2, 2, 3]
    ref = "/root/reference/tests/test_data/test/test.inter"
    if os.path.exists(ref):       # only in the build container
        ds = rg.InteractionDataset.from_inter_file(ref)
        assert (ds.user_num, ds.item_num) == (int(g1["U"]), int(g1["I"]))
        assert ds.inter_feat["user_id"].tolist() == g1["uid"].tolist()
        assert ds.inter_feat["item_id"].tolist() == g1["iid"].tolist()


def test_chain_sync_struct_layout(tmp_path):
    c = _lib.ChainSync
    fields = [f[0] for f in c._fields_]
    src = tmp_path / "off.c"
    body = "".join(f'printf("{f} %zu\\n", offsetof(b200gcn_chain_sync, {f}));\n' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200gcn.h"\nint main(void){\n' + body +
                   'printf("sizeof %zu %d %d\\n", sizeof(b200gcn_chain_sync), B200GCN_CHAIN_MAX_PHASES, '
                   'B200GCN_CHAIN_MAX_RANKS);return 0;}\n')
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    got = dict(l.rsplit(" ", 1) for l in out if not l.startswith("sizeof"))
    for f in fields:
        assert int(got[f]) == getattr(c, f).offset, f
    sz = out[-1].split()
    assert int(sz[1]) == ctypes.sizeof(c)
    assert (int(sz[2]), int(sz[3])) == (_lib.CHAIN_MAX_PHASES, _lib.CHAIN_MAX_RANKS)
    assert _lib.load().b200gcn_spmm_chain(None, 0, None, None) == _lib.ERR_INVALID


def test_hub_plan_struct_layout():
    h = _lib.HubPlan
    assert h.n_hubs.offset == 0 and h.n_chunks.offset == 4 and h.hub_rows.offset == 8
    assert h.hub_chunk_ptr.offset == 16 and h.chunk_beg.offset == 24 and h.chunk_end.offset == 32
    assert h.scratch.offset == 40 and ctypes.sizeof(h) == 48
    lib = _lib.load()
    assert lib.b200gcn_spmm_hubs(None, None, None) == _lib.ERR_INVALID


def test_models_refuse_cpu_device():
    """A model built for a CPU device keeps a *described* graph and its forward raises: no silent CPU path."""
    ds = rg.InteractionDataset(torch.tensor([1, 2, 2]), torch.tensor([1, 1, 2]), 3, 3)
    m = rg.LightGCN({"device": "cpu", "enable_sparse": True, "embedding_size": 8, "n_layers": 2}, ds)
    assert m.use_sparse and not m.edge_index.is_resident
    with pytest.raises(RuntimeError):
        m.forward()
    m.fused = False
    with pytest.raises(RuntimeError):
        m.forward()
    with pytest.raises(ValueError):
        rg.LightGCN({"device": "cpu", "enable_sparse": 1.5, "embedding_size": 8, "n_layers": 2}, ds)
    from recbole_gnn_b200.host import HostPropagator
    with pytest.raises(RuntimeError):
        HostPropagator(m.edge_index, 3, 3, 8, 2)


def test_dispatcher_ops_are_registered_with_fake_kernels_and_refuse_cpu():
    """`torch.ops.b200gcn.*` exist, trace under fake tensors (what torch.compile / export need) and have no
    CPU kernel to fall back to."""
    import recbole_gnn_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("spmm", "lightgcn_propagate", "bignn_tail"):
        assert hasattr(torch.ops.b200gcn, name)
    with FakeTensorMode():
        rowptr = torch.empty(101, dtype=torch.int64, device="cuda")
        col, val = torch.empty(500, dtype=torch.int32, device="cuda"), torch.empty(500, device="cuda")
        y = torch.ops.b200gcn.spmm(rowptr, col, val, torch.empty(100, 64, device="cuda", requires_grad=True), 100, True)
        assert y.shape == (100, 64) and y.requires_grad
        u, i = torch.ops.b200gcn.lightgcn_propagate(rowptr, col, None, torch.empty(40, 64, device="cuda"),
                                                    torch.empty(60, 64, device="cuda"), 3)
        assert u.shape == (40, 64) and i.shape == (60, 64)
        t = torch.ops.b200gcn.bignn_tail(*(torch.empty(s, device="cuda") for s in ((9, 64), (9, 64), (32, 64), (32,),
                                                                                   (32, 64), (32,))), 0.2, True)
        assert t.shape == (9, 32)
    with pytest.raises(NotImplementedError):
        torch.ops.b200gcn.spmm(torch.zeros(3, dtype=torch.int64), torch.zeros(1, dtype=torch.int32), None,
                               torch.zeros(2, 4), 2, True)


def test_native_inter_reader_edge_cases(tmp_path):
    """b200gcn_inter_open (host C++): header-driven column order, CRLF, short / blank lines skipped, first-appearance
    remap starting at 1, missing and empty files rejected — against the oracle's plain-Python reader where both apply."""
    from oracle import oracle as O
    p = tmp_path / "a.inter"
    p.write_text("item_id:token\trating:float\tuser_id:token\r\n"
                 "i9\t3\tbob\r\n" "i9\t1\tamy\r\n" "\r\n" "i2\t5\r\n" "i2\t5\tbob\r\n" "i 7\t2\tcy d\n")
    ds = rg.InteractionDataset.from_inter_file(str(p))
    assert ds.inter_feat["user_id"].tolist() == [1, 2, 1, 3] and ds.inter_feat["item_id"].tolist() == [1, 1, 2, 3]
    assert (ds.user_num, ds.item_num) == (4, 4)
    q = tmp_path / "b.inter"
    lines = ["user_id:token\titem_id:token\ttimestamp:float"] + [f"u{(k * 7) % 53}\ti{(k * 11) % 31}\t{k}" for k in range(500)]
    q.write_text("\n".join(lines))                        # no trailing newline
    ds = rg.InteractionDataset.from_inter_file(str(q))
    u, i, U, I = O.load_inter_file(str(q))
    assert torch.equal(ds.inter_feat["user_id"], u) and torch.equal(ds.inter_feat["item_id"], i)
    assert (ds.user_num, ds.item_num) == (U, I)
    with pytest.raises(ValueError):
        rg.InteractionDataset.from_inter_file(str(tmp_path / "missing.inter"))
    e = tmp_path / "empty.inter"
    e.write_text("")
    with pytest.raises(ValueError):
        rg.InteractionDataset.from_inter_file(str(e))
    h = tmp_path / "header_only.inter"
    h.write_text("user_id:token\titem_id:token\n")
    ds = rg.InteractionDataset.from_inter_file(str(h))
    assert ds.inter_feat["user_id"].numel() == 0 and (ds.user_num, ds.item_num) == (1, 1)


def test_round2_entry_points_refuse_cpu_tensors():
    """No CPU fallback anywhere on the widened path: training step, full-sort evaluation, graph re-sampling."""
    from recbole_gnn_b200 import augment, functional as F_, train
    u, it = torch.randn(4, 8), torch.randn(6, 8)
    ids = torch.tensor([1, 2])
    with pytest.raises(RuntimeError):
        F_.full_sort_topk(u, it, 2)
    with pytest.raises(RuntimeError):
        F_.full_sort_scores(u, it)
    with pytest.raises(RuntimeError):
        train.bpr_loss_fused(u, it, u, it, ids, ids, ids, reg_weight=1e-5)
    with pytest.raises(RuntimeError):
        train.adam_step(u, u.clone(), u.clone(), u.clone(), lr=1e-3, step=1)
    with pytest.raises(RuntimeError):
        augment.SGLAugmenter(ids, ids, 4, 6, "cpu")
    with pytest.raises(RuntimeError):
        augment.sept_norm_edge_weight(torch.tensor([[0, 1], [1, 0]]), 2)
    with pytest.raises(RuntimeError):
        F_.bignn_tail(u, u, it[:, :8], it[:, 0], it[:, :8], it[:, 0])
