"""Timeline of a back-to-back chain of SHARDED MPPI/ODE ticks (diagnostics; one process per GPU under torchrun):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/k1_chain_trace_dist.py [N_global] [ticks]
Every rank prints the last launches of its chain: rollouts done, last local record stamp, records polled (local + peers'),
combined, tick finished, period."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from control_toolkit_b200 import _lib as L  # noqa: E402
from control_toolkit_b200.distributed import ShardPlan  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    torch.cuda.set_device(lr)
    lib = L.load()
    plan = ShardPlan(rank, world)
    ctrl, N, H = bench.build_controller("mppi_ode_1m", shard=plan, device=lr, n_override=n)
    opt = ctrl.optimizer
    plan.attach(opt, lib)
    states = torch.from_numpy(bench.synthetic_states(ticks, 0)).to(f"cuda:{lr}")
    u = torch.zeros(ticks, 2, device=f"cuda:{lr}")
    grid = C.c_int()
    L.check(lib.ctk_debug_trace(opt._h, 1, None, 0, C.byref(grid)))
    per = 320 * 8
    buf = np.zeros(4 * per, np.uint64)
    L.check(lib.ctk_step_device_n(opt._h, C.c_void_p(states.data_ptr()), 6, C.c_void_p(u.data_ptr()), 2, 10))
    torch.cuda.synchronize()
    dist.barrier()
    L.check(lib.ctk_exchange_barrier(opt._h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.ctk_step_device_n(opt._h, C.c_void_p(states.data_ptr()), 6, C.c_void_p(u.data_ptr()), 2, ticks))
    e1.record()
    torch.cuda.synchronize()
    L.check(lib.ctk_debug_trace(opt._h, 1, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(grid)))
    g = grid.value
    t = buf.reshape(4, 320, 8)[:, :g, :].astype(np.int64)
    t0 = t[1, :, 0].min()
    t = (t - t0) / 1e3
    lines = [f"rank {rank}: N_local={opt._n_local} grid={g} exchange={opt._exchange} hops={os.environ.get('CTK_EXCHANGE_HOPS', '1')}: "
             f"{e0.elapsed_time(e1) * 1e3 / ticks:.2f} us per tick"]
    for i in range(1, 4):
        a = t[i]
        prev_fin = t[i - 1, 0, 5]
        lines.append(f"   {i}: prologue done {np.median(a[:, 1]):8.2f} | rollouts done median {np.median(a[:, 2]):8.2f} last {a[:, 2].max():8.2f} | last local record {a[:, 4].max():8.2f} | "
                     f"polled +{a[0, 6] - a[:, 4].max():5.2f} | combined +{a[0, 7] - a[0, 6]:5.2f} | finished +{a[0, 5] - a[0, 7]:5.2f} | period {a[0, 5] - prev_fin:6.2f}")
    for r in range(world):
        if r == rank and (rank in (0, world - 1)):
            print("\n".join(lines), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
