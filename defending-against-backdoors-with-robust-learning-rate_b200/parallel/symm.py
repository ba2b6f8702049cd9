"""Symmetric (peer-mapped) device memory.

Every rank allocates one slab of the same size; after ``rendezvous`` each rank holds a device pointer to every
peer's slab (NVLink/NVSwitch P2P) and, when the fabric supports it, one NVLS *multicast* pointer whose stores land
in all slabs at once.  Two providers:

1. ``torch.distributed._symmetric_memory`` (CUDA VMM + multicast objects) -- preferred: gives ``multicast_ptr``.
2. CUDA IPC (``ops/csrc/binding.cpp``: cudaMalloc + cudaIpcGetMemHandle / OpenMemHandle) -- no multicast; the
   aggregation kernel then issues one P2P store per peer.

Tensors carved from the slab are plain torch tensors; kernels receive raw peer pointers (base + byte offset).
"""
from __future__ import annotations

import torch

from .. import ops


class SymmetricBuffer:
    def __init__(self, ctx, nbytes: int, provider: str = "auto"):
        self.ctx = ctx
        self.nbytes = (int(nbytes) + 4095) // 4096 * 4096
        self.provider = None
        self.multicast_ptr = 0
        self.peer_ptrs: list[int] = []
        self._hdl = None
        self._ipc_opened: list[int] = []
        self._ipc_base = 0
        dev = ctx.device
        if dev.type != "cuda" or not ctx.is_dist:
            self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
            self.peer_ptrs = [self.local.data_ptr()] * max(1, ctx.world)
            self.provider = "local"
            return
        errs = []
        if provider in ("auto", "torch"):
            try:
                self._init_torch_symm()
            except Exception as e:  # noqa: BLE001
                errs.append(f"torch symm_mem: {type(e).__name__}: {e}")
        if self.provider is None and provider in ("auto", "ipc"):
            try:
                self._init_ipc()
            except Exception as e:  # noqa: BLE001
                errs.append(f"cuda ipc: {type(e).__name__}: {e}")
        if self.provider is None:
            raise RuntimeError("no symmetric-memory provider available: " + " | ".join(errs))
        if errs and ctx.is_main:
            print("[symm] fell back to", self.provider, "after:", " | ".join(errs))

    # ---- providers -------------------------------------------------------------------------------------------
    def _init_torch_symm(self):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.ctx.device)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        self.local = t
        self.local.zero_()
        self._hdl = hdl
        self.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
        try:
            self.multicast_ptr = int(hdl.multicast_ptr) if hdl.has_multicast_support(self.ctx.device.type, self.ctx.device.index) else 0
        except Exception:  # noqa: BLE001 - API differs across torch versions
            self.multicast_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self.provider = "torch_symm"
        torch.cuda.synchronize(self.ctx.device)
        self.ctx.barrier()

    def _init_ipc(self):
        e = ops.ext()
        dev = self.ctx.device.index
        base, handle = e.ipc_alloc(self.nbytes, dev)
        self._ipc_base = base
        handles = self.ctx.all_gather_object(bytes(handle))
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.ctx.rank:
                ptrs.append(base)
            else:
                p = e.ipc_open(h, dev)
                self._ipc_opened.append(p)
                ptrs.append(p)
        self.peer_ptrs = ptrs
        self.local = e.tensor_from_ptr(base, [self.nbytes], torch.uint8, dev)
        self.provider = "cuda_ipc"
        self.ctx.barrier()

    # ---- carving ---------------------------------------------------------------------------------------------
    def tensor(self, byte_offset: int, numel: int, dtype):
        """Local tensor view over ``[byte_offset, byte_offset + numel*itemsize)`` of this rank's slab."""
        item = torch.empty((), dtype=dtype).element_size()
        assert byte_offset % 16 == 0 and byte_offset + numel * item <= self.nbytes
        return self.local[byte_offset:byte_offset + numel * item].view(dtype)

    def peer_ptr(self, rank: int, byte_offset: int) -> int:
        return self.peer_ptrs[rank] + byte_offset

    def mc_ptr(self, byte_offset: int) -> int:
        return self.multicast_ptr + byte_offset if self.multicast_ptr else 0

    def close(self):
        if self.provider == "cuda_ipc":
            e = ops.ext()
            torch.cuda.synchronize(self.ctx.device)
            self.ctx.barrier()
            for p in self._ipc_opened:
                e.ipc_close(p)
            self.ctx.barrier()
            e.ipc_free(self._ipc_base)
            self._ipc_opened, self._ipc_base = [], 0
            self.provider = "closed"
