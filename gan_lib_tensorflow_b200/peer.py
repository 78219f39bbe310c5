"""Peer-memory exchanges between the GPUs of one node (csrc/peer.cu): small all-reduces -- cross-GPU BatchNorm moments,
the backward sums of a synced batch norm -- as ONE single-CTA kernel per call over NVLink / NVSwitch peer memory instead
of a library collective, so that they can sit inside captured CUDA graphs and cost a few microseconds each.

torch supplies the plumbing only: one symmetric allocation per rank (torch.distributed._symmetric_memory: CUDA VMM
handles exchanged through the process group's store) mapped into every process; all arithmetic and the signalling run
in libganb200 kernels."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int64, c_void_p

import torch

from . import kernels as K
from .cabi import check, ptr


class PeerComm:
    """Symmetric buffer + per-call-site regions.  Every rank must issue the same calls with the same keys; a key names
    one call site (one exchange per training step), e.g. 'Generator/G.Block.1/N1/n128/fwd'."""

    def __init__(self, group=None, nbytes: int = 16 << 20, device=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("PeerComm covers the GPUs of one node (world <= 8)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nbytes = int(nbytes)
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert ptrs[self.rank] == self.buf.data_ptr(), "symmetric memory: own mapping differs from the local allocation"
        self._ptrs = (c_void_p * self.world)(*ptrs)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # every buffer is zeroed (flags, epochs) before any rank signals
        torch.cuda.synchronize(self.device)
        self._sites: dict[str, tuple[int, int]] = {}
        self._next = 0
        self.max_count = int(K.L().ganb_peer_max_count())
        K.L().ganb_peer_site_bytes.restype = c_int64

    def _site(self, key: str, count: int) -> int:
        hit = self._sites.get(key)
        if hit is not None:
            if hit[1] != count:
                raise ValueError(f"peer site {key!r} was created for {hit[1]} floats, now {count}")
            return hit[0]
        size = int(K.L().ganb_peer_site_bytes(count))
        if self._next + size > self.nbytes:
            raise MemoryError("PeerComm buffer exhausted: raise nbytes")
        off = self._next
        self._next += size
        self._sites[key] = (off, count)
        return off

    def allreduce(self, key: str, x: torch.Tensor, out: torch.Tensor | None = None, scale: float = 1.0) -> torch.Tensor:
        """out = scale * sum over ranks of x (fp32, contiguous, <= max_count elements); bit-identical on every rank."""
        assert x.dtype == torch.float32 and x.is_contiguous()
        out = x if out is None else out
        n = x.numel()
        if n > self.max_count:            # larger messages go in slices (each its own site)
            xf, of = x.reshape(-1), out.reshape(-1)
            for i, s in enumerate(range(0, n, self.max_count)):
                e = min(n, s + self.max_count)
                self.allreduce(f"{key}#{i}", xf[s:e], of[s:e], scale)
            return out
        off = self._site(key, n)
        check(K.L().ganb_peer_allreduce(ptr(x), ptr(out), n, c_float(scale), self._ptrs, self.rank, self.world,
                                        c_int64(off), K._stream()), "ganb_peer_allreduce")
        return out

    def bn_moments(self, key: str, mean: torch.Tensor, rstd: torch.Tensor, eps: float) -> None:
        """(mean, rstd) of the local share -> statistics over all ranks (equal shares), in place, one launch."""
        n = mean.numel()
        assert rstd.numel() == n and mean.dtype == torch.float32 and rstd.dtype == torch.float32
        if 2 * n > self.max_count:
            raise ValueError(f"bn_moments: {n} statistics exceed one exchange ({self.max_count // 2})")
        off = self._site(key, 2 * n)
        check(K.L().ganb_peer_bn_moments(ptr(mean), ptr(rstd), n, c_float(eps), self._ptrs, self.rank, self.world,
                                         c_int64(off), K._stream()), "ganb_peer_bn_moments")


class NcclSync:
    """The same interface over a library collective (eager mode; multi-node or no peer access)."""

    def __init__(self, allreduce_sum, world: int):
        self.allreduce_sum, self.world = allreduce_sum, world

    def allreduce(self, key, x, out=None, scale=1.0):
        out = x if out is None else out.copy_(x)
        self.allreduce_sum(out)
        if scale != 1.0:
            out.mul_(scale)
        return out
