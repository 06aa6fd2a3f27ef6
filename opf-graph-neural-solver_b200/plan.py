"""Topology plan: host wrapper of ``gns_plan_*`` (include/gns_b200.h).

Replaces the per-call index tensors of the reference (ref GNS/main.py:35-36, 85-86,
144, 153, 184-185) with a one-time CSR build in the native library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class TopologyPlan:
    """CSR-by-receiver / CSR-by-sender / generator-by-bus plan for one grid topology."""

    def __init__(self, f_bus, t_bus, gen_bus, n_bus: int, device: int = 0):
        lib = _lib.load_library()
        self.f_bus = np.ascontiguousarray(f_bus, dtype=np.int32)
        self.t_bus = np.ascontiguousarray(t_bus, dtype=np.int32)
        self.gen_bus = np.ascontiguousarray(gen_bus, dtype=np.int32)
        self.n_bus, self.n_line, self.n_gen = int(n_bus), int(self.f_bus.size), int(self.gen_bus.size)
        self.device = int(device)
        handle = C.c_void_p()
        rc = lib.gns_plan_create(self.n_bus, self.n_line, self.n_gen,
                                 self.f_bus.ctypes.data, self.t_bus.ctypes.data,
                                 self.gen_bus.ctypes.data if self.n_gen else None,
                                 self.device, C.byref(handle))
        if rc != 0:
            msg = _lib.last_error()
            # same exception class the reference raises for out-of-range bus ids
            raise IndexError(msg) if rc == -1 else RuntimeError(msg)
        self._h = handle
        self._lib = lib

    @classmethod
    def from_tensors(cls, lines: torch.Tensor, generators: torch.Tensor, n_bus: int, device: int = 0):
        """Build from the packed tensors (first grid of a batch); 1-based float ids like
        ref GNS/main.py:35-36,144."""
        l0 = lines[0] if lines.dim() == 3 else lines
        g0 = generators[0] if generators.dim() == 3 else generators
        ft = l0[:, :2].detach().to("cpu", torch.float32).numpy()
        gb = g0[:, 0].detach().to("cpu", torch.float32).numpy()
        return cls(ft[:, 0].astype(np.int64) - 1, ft[:, 1].astype(np.int64) - 1,
                   gb.astype(np.int64) - 1, n_bus, device)

    def key(self):
        return (self.n_bus, self.f_bus.tobytes(), self.t_bus.tobytes(), self.gen_bus.tobytes())

    @property
    def handle(self):
        return self._h

    def export(self, name: str) -> np.ndarray:
        n = self._lib.gns_plan_export(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(_lib.last_error())
        out = np.zeros(n, dtype=np.int32)
        self._lib.gns_plan_export(self._h, name.encode(), out.ctypes.data, n)
        return out

    def matches(self, lines: torch.Tensor, generators: torch.Tensor) -> bool:
        """Device-side check that every grid of the batch carries this topology (synchronises)."""
        S = lines.shape[0]
        rc = self._lib.gns_check_topology(self._h, lines.data_ptr(), generators.data_ptr(), S,
                                          torch.cuda.current_stream(lines.device).cuda_stream)
        if rc < 0:
            raise RuntimeError(_lib.last_error())
        return rc == 0

    def matches_host(self, lines: torch.Tensor, generators: torch.Tensor) -> bool:
        """Same check on HOST tensors (reference-style CPU inputs): no device work, no stream synchronisation."""
        if getattr(self, "_expect_host", None) is None:
            self._expect_host = (torch.from_numpy(self.f_bus.astype(np.float32) + 1), torch.from_numpy(self.t_bus.astype(np.float32) + 1),
                                 torch.from_numpy(self.gen_bus.astype(np.float32) + 1))
        ef, et, eg = self._expect_host
        l, g = lines.detach(), generators.detach()
        return bool((l[..., 0] == ef).all()) and bool((l[..., 1] == et).all()) and bool((g[..., 0] == eg).all())

    def check_async(self, lines: torch.Tensor, generators: torch.Tensor, flag: torch.Tensor):
        """Device-side check without synchronisation: ORs 1 into the int32 device tensor `flag` on a mismatch."""
        rc = self._lib.gns_check_topology_async(self._h, lines.data_ptr(), generators.data_ptr(), lines.shape[0],
                                                flag.data_ptr(), torch.cuda.current_stream(lines.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(_lib.last_error())

    def launch_info(self, S, K, latent_dim, hidden_dim, multiple_phi, backward=False):
        out = (C.c_int32 * 8)()
        rc = self._lib.gns_launch_info(self._h, S, K, latent_dim, hidden_dim, int(multiple_phi),
                                       int(backward), out)
        _lib.check(rc, "gns_launch_info")
        keys = ["grids_per_cta", "threads", "smem_bytes", "ctas", "vector_width", "cta_batches", "sms", "tmax"]
        return dict(zip(keys, list(out)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.gns_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
