"""Peer-memory communicator of the data-parallel step.

One process per GPU.  Every rank allocates one buffer through the C ABI (`clearvae_peer_alloc`), exports its CUDA IPC
handle, and maps the buffers of all peers of the node (`clearvae_peer_open`, NVLink peer access); `torch.distributed`
only carries the 64-byte handles at setup.  `gather` / `allreduce` are then single kernel launches
(`csrc/peer_comm.cu`) on the current stream — graph-capturable, no NCCL in the step.

`PeerComm.create` returns None when peer mapping is unavailable (different nodes, IPC disabled) or when its self-test
against NCCL disagrees on any rank; callers then keep the NCCL collectives.  Both are GPU paths; there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import sys

import torch

from . import _ops

HEADER = 4096


class PeerComm:
    def __init__(self, group, rank: int, world: int, device, nbytes: int = 64 << 20):
        import torch.distributed as td
        if world > 8:
            raise RuntimeError("peer-memory collectives cover one NVSwitch node (<= 8 ranks)")
        self.group, self.rank, self.world, self.device, self.nbytes = group, rank, world, torch.device(device), int(nbytes)
        lib = ctypes.CDLL(_ops.lib_paths()[0])
        for fn, args in (("clearvae_peer_alloc", [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]),
                         ("clearvae_peer_free", [ctypes.c_void_p]),
                         ("clearvae_peer_export", [ctypes.c_void_p, ctypes.c_char_p]),
                         ("clearvae_peer_open", [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
                         ("clearvae_peer_close", [ctypes.c_void_p]),
                         ("clearvae_peer_error", [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]),
                         ("clearvae_peer_timeline", [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)])):
            getattr(lib, fn).argtypes = args
            getattr(lib, fn).restype = ctypes.c_int
        self._lib = lib
        self._opened = []
        self._local = None
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            rc = lib.clearvae_peer_alloc(self.nbytes, ctypes.byref(p))
            if rc:
                raise RuntimeError(f"clearvae_peer_alloc failed ({rc})")
            self._local = p.value
            handle = ctypes.create_string_buffer(64)
            rc = lib.clearvae_peer_export(p, handle)
            mine = (rc, bytes(handle.raw))
            if world > 1:
                everyone = [None] * world
                td.all_gather_object(everyone, mine, group=group)
            else:
                everyone = [mine]
            if any(r != 0 for r, _ in everyone):
                raise RuntimeError(f"cudaIpcGetMemHandle failed on some rank: {[r for r, _ in everyone]}")
            bases = []
            for r, (_, h) in enumerate(everyone):
                if r == rank:
                    bases.append(self._local)
                    continue
                q = ctypes.c_void_p()
                rc = lib.clearvae_peer_open(h, ctypes.byref(q))
                if rc:
                    raise RuntimeError(f"cudaIpcOpenMemHandle(rank {r}) failed ({rc})")
                self._opened.append(q.value)
                bases.append(q.value)
            self.bases = bases

    # ---- collectives (one launch each, current stream) ---------------------------------------------------------
    def gather(self, pieces):
        """pieces: list (<= 8) of contiguous CUDA tensors [B, ...]; returns the rank-major concatenations [world*B, ...]."""
        src = [t.contiguous() for t in pieces]
        dst = [torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in src]
        _ops.ops().peer_gather(self.bases, self.rank, self.nbytes, src, dst)
        return dst

    def allreduce_(self, tensors):
        """in-place sum over ranks of a list of contiguous fp32 CUDA tensors (fixed rank order: identical bits on all ranks)."""
        _ops.ops().peer_allreduce(self.bases, self.rank, self.nbytes, list(tensors))

    def slot_bytes(self):
        return ((self.nbytes - HEADER) // 2) & ~15

    def fits(self, pieces) -> bool:
        """whether one `gather` of these per-rank pieces fits a slot (callers fall back to the NCCL all-gather otherwise)"""
        return len(pieces) <= 8 and sum(t.numel() * t.element_size() + 16 for t in pieces) * self.world <= self.slot_bytes()

    def allreduce_chunked_(self, tensors):
        """`allreduce_` for any total size: the list is cut into groups that fit one slot (<= 64 tensors each); a tensor larger
        than a slot travels as consecutive flat views.  Same fixed rank order, so results stay bit-identical across ranks."""
        cap = self.slot_bytes() // 4 - 64
        group, used = [], 0

        def flush():
            nonlocal group, used
            if group:
                self.allreduce_(group)
            group, used = [], 0

        for t in tensors:
            flat = t.view(-1)
            o = 0
            while o < flat.numel():
                n = min(flat.numel() - o, cap - used)
                if n <= 0 or len(group) >= 64:
                    flush()
                    continue
                n -= (n % 4) if (o + n < flat.numel()) else 0     # keep 16-byte alignment of the next view
                if n == 0:
                    flush()
                    continue
                group.append(flat[o:o + n])
                used += n + 4
                o += n
        flush()

    def check(self):
        """raise if any collective of this communicator ever timed out waiting for a peer (the kernels record the missing
        rank instead of hanging; results after that are undefined)"""
        e = self.error()
        if e:
            raise RuntimeError(f"clear_vae_b200.peer: a peer-memory collective timed out waiting for rank {e - 1}")

    def error(self) -> int:
        v = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            self._lib.clearvae_peer_error(self._local, ctypes.byref(v))
        return int(v.value)

    def timeline(self):
        """[64, 4] ns stamps (start, staged, peers ready, pulled) of CTA 0 for the last 64 calls (debug)."""
        buf = (ctypes.c_uint64 * (64 * 6))()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            self._lib.clearvae_peer_timeline(self._local, buf)
        return torch.tensor(list(buf), dtype=torch.float64).view(64, 6)[:, :4]

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for q in self._opened:
                self._lib.clearvae_peer_close(q)
            self._opened = []
            if self._local:
                self._lib.clearvae_peer_free(self._local)
                self._local = None

    # ---- guarded construction -------------------------------------------------------------------------------------
    @staticmethod
    def create(group, rank, world, device, nbytes: int = 64 << 20):
        """Collective.  PeerComm when every rank mapped every peer and the self-test matched NCCL, else None (on all ranks)."""
        import torch.distributed as td
        if os.environ.get("CLEARVAE_PEER", "1") == "0" or world < 2 or world > 8:
            return None
        pc, why = None, ""
        try:
            pc = PeerComm(group, rank, world, device, nbytes)
        except Exception as e:  # setup is all-or-nothing across ranks: agree below
            why = repr(e)
        ok = torch.tensor([1 if pc is not None else 0], device=device)
        td.all_reduce(ok, op=td.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            if why:
                print(f"[clear_vae_b200.peer] rank {rank}: peer memory unavailable, keeping NCCL: {why}", file=sys.stderr, flush=True)
            return None
        td.barrier(group=group)
        # self-test against NCCL: two gathers (both slots) and one all-reduce
        good = 1
        try:
            g = torch.Generator(device="cpu").manual_seed(1234 + rank)
            for n in (1000, 37):
                a = torch.randn(n, 5, generator=g).to(device)
                b = torch.randint(0, 1 << 40, (n,), generator=g).to(device)
                ga, gb = pc.gather([a, b])
                ra = torch.empty_like(ga)
                rb = torch.empty_like(gb)
                td.all_gather_into_tensor(ra, a, group=group)
                td.all_gather_into_tensor(rb, b, group=group)
                good &= int(torch.equal(ga, ra) and torch.equal(gb, rb))
            ts = [torch.randn(n, generator=g).to(device) for n in (7, 4096, 1, 130)]
            ref = [t.clone() for t in ts]
            pc.allreduce_(ts)
            for t in ref:
                td.all_reduce(t, group=group)
            good &= int(all(torch.allclose(a, b, rtol=1e-5, atol=1e-5) for a, b in zip(ts, ref)))
            torch.cuda.synchronize()
            good &= int(pc.error() == 0)
        except Exception as e:
            why = repr(e)
            good = 0
        ok = torch.tensor([good], device=device)
        td.all_reduce(ok, op=td.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            print(f"[clear_vae_b200.peer] rank {rank}: peer self-test failed, keeping NCCL {why}", file=sys.stderr, flush=True)
            return None
        return pc
