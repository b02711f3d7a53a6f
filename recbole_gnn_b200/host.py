"""Host-buffer entry point: LightGCN propagation whose inputs and result live in (pinned) HOST memory.

``HostPropagator.submit(hu, hi, out_u, out_i)`` copies the two embedding tables host->device, runs the fused
K-layer propagation and copies the result device->host, on three CUDA streams (copy-in / compute / copy-out)
with ``depth`` device buffer sets, so that consecutive submissions overlap: the PCIe transfers of step k+1 and
step k-1 ride under the kernels of step k.  Every step still moves its own inputs and result across PCIe; the
reference does the equivalent of none of this (its tables live on the device, `quick_start.py:41`) — this is
the engine's serving-style entry for callers that keep the tables on the host.
"""
from __future__ import annotations

from typing import List

import torch

from . import functional as F_
from .graph import GraphHandle


class HostPropagator:
    def __init__(self, g: GraphHandle, user_num: int, item_num: int, dim: int, n_layers: int, depth: int = 2):
        if not g.is_resident:
            raise RuntimeError("HostPropagator needs a resident GraphHandle")
        self.g, self.L, self.depth = g, int(n_layers), int(depth)
        dev = g.device
        self.dev = dev
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots: List[dict] = []
        for _ in range(self.depth):
            self.slots.append({
                "du": torch.empty(user_num, dim, dtype=torch.float32, device=dev),
                "di": torch.empty(item_num, dim, dtype=torch.float32, device=dev),
                "in_ready": torch.cuda.Event(), "cmp_done": torch.cuda.Event(), "out_done": torch.cuda.Event(),
                "res": None,
            })
        self.k = 0
        self.h2d_bytes_per_step = (user_num + item_num) * dim * 4
        self.d2h_bytes_per_step = (user_num + item_num) * dim * 4

    def submit(self, hu: torch.Tensor, hi: torch.Tensor, out_u: torch.Tensor, out_i: torch.Tensor) -> None:
        for t in (hu, hi, out_u, out_i):
            if t.is_cuda:
                raise ValueError("HostPropagator takes host tensors (pinned memory for asynchronous copies)")
        s = self.slots[self.k % self.depth]
        first_use = self.k < self.depth
        self.k += 1
        with torch.cuda.stream(self.s_in):
            if not first_use:
                self.s_in.wait_event(s["cmp_done"])        # the kernels that read du/di last time are done
            s["du"].copy_(hu, non_blocking=True)
            s["di"].copy_(hi, non_blocking=True)
            s["in_ready"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(s["in_ready"])
            if not first_use:
                self.s_cmp.wait_event(s["out_done"])       # previous result of this slot has left the device
            with torch.no_grad():
                u, i = F_.lightgcn_propagate(self.g, s["du"], s["di"], self.L)
            s["res"] = (u, i)
            s["cmp_done"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s["cmp_done"])
            out_u.copy_(u, non_blocking=True)
            out_i.copy_(i, non_blocking=True)
            u.record_stream(self.s_out)
            i.record_stream(self.s_out)
            s["out_done"].record(self.s_out)

    def synchronize(self) -> None:
        for st in (self.s_in, self.s_cmp, self.s_out):
            st.synchronize()
