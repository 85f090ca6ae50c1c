"""Peer-mapped slab buffers for the multi-GPU halo transport (``kx_peer_*`` / ``kx_halo_pull`` of ``include/kx_b200.h``).

One process per GPU: every rank allocates its slab with ``kx_peer_alloc`` (a plain ``cudaMalloc``, so the block can be
exported with CUDA IPC whatever torch's caching allocator is configured to do), the 64-byte handles travel once through
``torch.distributed.all_gather_object`` and every rank maps its two neighbours' buffers.  From then on the halo rows move
with one-sided loads over NVLink inside ``halo_pull_kernel``; torch only sees the local buffer, wrapped as a tensor
through ``__cuda_array_interface__``.

``PeerGroup.local(world, nbytes)`` builds the same structure for several emulated ranks inside ONE process on ONE GPU
(the "peer" pointers are ordinary local pointers): this is how the flag protocol and the pull kernel are tested on a
single-GPU box.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

READY_FROM_UP, READY_FROM_DOWN, DONE_FROM_UP, DONE_FROM_DOWN = 0, 1, 2, 3   # 64-bit flag slots at the buffer start
COUNTER_BYTE = 64                                                          # 32-bit scratch word of the pull kernel


class _DeviceBlock:
    """Raw device memory as a ``__cuda_array_interface__`` object (uint8), so torch can alias it without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2,
                                         "strides": None}


class PeerGroup:
    """This rank's buffer plus the mapped base pointers of the others (``None`` for ranks that were not mapped)."""

    def __init__(self, rank, world, nbytes, local_ptr, peer_ptrs, device, owner=True, opened=()):
        self.rank, self.world, self.nbytes, self.device = rank, world, int(nbytes), device
        self.local_ptr, self.peer_ptrs = int(local_ptr), list(peer_ptrs)
        self._owner, self._opened = owner, list(opened)
        self._block = _DeviceBlock(self.local_ptr, self.nbytes)
        self.bytes = torch.as_tensor(self._block, device=device)   # uint8 view of the whole local buffer

    # ---- construction ------------------------------------------------------------------------------------
    @classmethod
    def create(cls, nbytes: int, device, group=None):
        """Collective over ``group``: allocate, exchange IPC handles, map rank-1 and rank+1."""
        import torch.distributed as dist
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        with torch.cuda.device(device):
            ptr = C.c_void_p()
            _lib.check(lib.kx_peer_alloc(C.c_int64(nbytes), C.byref(ptr)))
            handle = (C.c_ubyte * _lib.KX_PEER_HANDLE_BYTES)()
            _lib.check(lib.kx_peer_export(ptr, handle))
            handles = [None] * world
            dist.all_gather_object(handles, (bytes(handle), int(nbytes)), group=group)
            if any(h[1] != nbytes for h in handles):
                raise ValueError("peer buffers must have the same size on every rank")
            peer_ptrs, opened = [None] * world, []
            peer_ptrs[rank] = ptr.value
            for r in (rank - 1, rank + 1):
                if 0 <= r < world:
                    raw = (C.c_ubyte * _lib.KX_PEER_HANDLE_BYTES).from_buffer_copy(handles[r][0])
                    mapped = C.c_void_p()
                    _lib.check(lib.kx_peer_open(raw, C.byref(mapped)))
                    peer_ptrs[r] = mapped.value
                    opened.append(mapped.value)
        dist.barrier(group=group)       # nobody touches a neighbour's flags before every mapping exists
        return cls(rank, world, nbytes, ptr.value, peer_ptrs, device, owner=True, opened=opened)

    @classmethod
    def local(cls, world: int, nbytes: int, device):
        """``world`` emulated ranks in this process on ``device`` (tests): peers are plain local pointers."""
        lib = _lib.load()
        ptrs = []
        with torch.cuda.device(device):
            for _ in range(world):
                ptr = C.c_void_p()
                _lib.check(lib.kx_peer_alloc(C.c_int64(nbytes), C.byref(ptr)))
                ptrs.append(ptr.value)
        return [cls(r, world, nbytes, ptrs[r], list(ptrs), device, owner=True) for r in range(world)]

    # ---- views ---------------------------------------------------------------------------------------------
    def tensor(self, byte_offset: int, shape, dtype=torch.float32):
        n = int(torch.tensor([], dtype=dtype).element_size())
        for s in shape:
            n *= int(s)
        return self.bytes[byte_offset:byte_offset + n].view(dtype).view(tuple(shape))

    def flag_ptr(self, rank: int, slot: int) -> int:
        return self.peer_ptrs[rank] + 8 * slot

    def close(self):
        """Unmap the neighbours and free the local block.  Collective in spirit: call it on every rank after a barrier
        (freeing a block a neighbour still reads is undefined).  Not called from ``__del__``: process exit releases
        everything, and tensors returned to the user may still alias the block."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.kx_peer_close(C.c_void_p(p))
            self._opened = []
            if self._owner and self.local_ptr:
                lib.kx_peer_free(C.c_void_p(self.local_ptr))
                self._owner = False
