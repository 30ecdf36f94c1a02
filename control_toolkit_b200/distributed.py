"""Sharding of the rollout population over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The path shards naturally by sample: rank r evaluates the global rollout ids
[offset_r, offset_r + count_r).  Philox counters and the injected-noise buffer are indexed by GLOBAL id, so the
sampled population is independent of the number of shards.  Per tick the only exchange is

* MPPI : the per-shard softmin record [rho_r, a_r, b_z,r[n_ind]] (n_ind + 2 floats).  Default ("p2p"): every shard's
         rollout kernel stores its record straight into the peers' mailboxes over NVLink (CUDA IPC mapped memory) and the
         finisher block of each shard combines them -- the whole sharded tick is ONE kernel launch per GPU, no NCCL call.
         Fallback ("nccl", or CTK_EXCHANGE=nccl): one all-gather of the record between two kernels (staged),
* CEM  : per outer iteration k (ordered-cost, global-id) keys per shard.  Default ("p2p"): the refit kernel of every shard stores
         its k keys into the peers' mailboxes over NVLink and merges the world x k keys it finds in its own -- the sharded tick
         is a chain of asynchronous launches per GPU, no NCCL call, no host round trip.  Fallback ("nccl"): one all-gather of the
         keys per outer iteration (staged).  The elite control rows are regenerated from the counter-based noise on every rank,
         never sent,
* RPGD : none (replicas only).

Every rank then runs the same combine/update kernel on the gathered records, so the optimizer state stays replicated.
torch.distributed is plumbing only (rendezvous + the NCCL all-gather on the handle's stream).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def shard_geometry(n_global: int, rank: int, world_size: int):
    """Contiguous split; the first (n_global % world_size) ranks get one extra rollout.  -> (offset, count)"""
    base, rem = divmod(int(n_global), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class ShardPlan:
    def __init__(self, rank: int = 0, world_size: int = 1, group=None):
        self.rank, self.world_size, self.group = int(rank), int(world_size), group
        self._bufs = {}

    @staticmethod
    def from_env():
        """One process per GPU launched by torchrun: RANK / WORLD_SIZE from the environment."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return ShardPlan(dist.get_rank(), dist.get_world_size())
        return ShardPlan(0, 1)

    def local_count(self, n_global: int) -> int:
        return shard_geometry(n_global, self.rank, self.world_size)[1]

    def local_offset(self, n_global: int) -> int:
        return shard_geometry(n_global, self.rank, self.world_size)[0]

    # -- exchange -------------------------------------------------------------------------------------------------
    def all_gather(self, record):
        """record: 1-D torch tensor (CUDA under NCCL, CPU under gloo) -> [world_size * len(record)] tensor."""
        import torch
        import torch.distributed as dist
        key = (record.device, record.numel(), record.dtype)
        out = self._bufs.get(key)
        if out is None:
            out = torch.empty(self.world_size * record.numel(), dtype=record.dtype, device=record.device)
            self._bufs[key] = out
        if self.world_size == 1:
            out.copy_(record)
        else:
            dist.all_gather_into_tensor(out, record, group=self.group)
        return out

    # -- one sharded tick through the C ABI -----------------------------------------------------------------------
    def attach(self, opt, lib) -> None:
        """Run the handle on torch's current stream (ordering with NCCL / torch events) and, for MPPI on more than one
        shard, connect the fused NVLink exchange: all-gather the 64-byte CUDA IPC handles of the mailboxes."""
        import os

        import torch
        import torch.distributed as dist
        torch.cuda.set_device(opt.device)
        L.check(lib.ctk_set_stream(opt._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        opt._shard_attached = True
        opt._exchange = "none" if self.world_size == 1 else "nccl"
        if self.world_size > 1 and opt._OPT in (L.OPT_MPPI, L.OPT_CEM) and os.environ.get("CTK_EXCHANGE", "p2p") != "nccl":
            mine = (C.c_ubyte * 64)()
            L.check(lib.ctk_exchange_export(opt._h, C.cast(mine, C.c_void_p)))
            dev = f"cuda:{opt.device}"
            t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
            allh = torch.empty(64 * self.world_size, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, t, group=self.group)
            blob = bytes(allh.cpu().numpy().tobytes())
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            rc = lib.ctk_exchange_connect(opt._h, self.rank, self.world_size, C.cast(buf, C.c_void_p))
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # all shards or none
            if int(ok.item()) == 1:
                opt._exchange = "p2p"
            else:
                import warnings
                warnings.warn("fused NVLink exchange unavailable (%s); using the staged NCCL all-gather"
                              % (lib.ctk_last_error() or b"").decode())
                if rc == 0:  # connected here but not everywhere: fall back consistently
                    none = (C.c_ubyte * 64)()
                    lib.ctk_exchange_connect(opt._h, 0, 1, C.cast(none, C.c_void_p))
            dist.barrier(group=self.group)  # every mailbox is mapped before the first tick stores into it

    def run_tick_device(self, opt, lib, s_dev_ptr: int, u_dev_ptr: int) -> None:
        """Asynchronous sharded tick: s and u stay on the device (used by bench.py's device-resident timing)."""
        import torch
        if getattr(opt, "_exchange", "nccl") in ("p2p", "none"):
            L.check(lib.ctk_step_device(opt._h, C.c_void_p(s_dev_ptr), C.c_void_p(u_dev_ptr)))
            return
        while True:
            more = L.check(lib.ctk_step_local(opt._h, C.c_void_p(s_dev_ptr)))
            ptr, n = C.c_void_p(), C.c_size_t()
            L.check(lib.ctk_partials(opt._h, C.byref(ptr), C.byref(n)))
            if n.value:
                rec = torch.as_tensor(_DevArray(ptr.value, n.value), device=f"cuda:{opt.device}")
                gathered = self.all_gather(rec)
                gptr = gathered.data_ptr()
            else:
                gptr = 0
            L.check(lib.ctk_step_finish(opt._h, C.c_void_p(gptr), self.world_size if n.value else 1, C.c_void_p(u_dev_ptr)))
            if not more:
                break

    def run_tick(self, opt, lib, s32: np.ndarray, state_which=None, state_n: int = 0) -> np.ndarray:
        """Sharded tick with host in / host out (the optimizer plugin's step()).  With the fused exchange the [H] warm-start
        sequence comes back through the same host mirror as u (opt._state_buf), as in the unsharded call."""
        import torch
        if not getattr(opt, "_shard_attached", False):
            self.attach(opt, lib)
        if opt._exchange in ("p2p", "none"):  # the C call does the staging, the fused tick and the read-back
            if state_which is not None:
                if opt._state_buf is None or opt._state_buf.size != state_n:
                    opt._state_buf = np.empty(state_n, np.float32)
                L.check(lib.ctk_step_state(opt._h, L.fptr(s32), L.fptr(opt._u_buf), state_which, L.fptr(opt._state_buf), state_n))
            else:
                opt._state_buf = None
                L.check(lib.ctk_step(opt._h, L.fptr(s32), L.fptr(opt._u_buf)))
            return opt._u_buf.copy()
        opt._state_buf = None
        dev = f"cuda:{opt.device}"
        if not hasattr(opt, "_s_pin"):
            opt._s_pin = torch.empty(6, dtype=torch.float32).pin_memory()
            opt._s_dev = torch.empty(6, dtype=torch.float32, device=dev)
            opt._u_dev = torch.zeros(4, dtype=torch.float32, device=dev)
        opt._s_pin.copy_(torch.from_numpy(s32))
        opt._s_dev.copy_(opt._s_pin, non_blocking=True)
        self.run_tick_device(opt, lib, opt._s_dev.data_ptr(), opt._u_dev.data_ptr())
        out = opt._u_dev[:2].cpu().numpy()
        if opt._exchange == "p2p" and out[1] != 0.0:
            raise RuntimeError("cross-GPU exchange timed out: a peer shard did not deliver its record within 2 s")
        return out[:1]
