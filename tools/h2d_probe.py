"""Aggregate pinned-host -> device copy bandwidth of this box with N ranks copying at the same time (the ceiling of
bench.py's `e2e` figure, whose timed region holds the H2D copy of every stream).

  python tools/h2d_probe.py                                      one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Each rank copies 2 GiB in 256 MiB pieces from its own pinned buffer, all ranks between two barriers; rank 0 prints one
JSON line: per-rank GB/s, the aggregate, and the same with write-combined host memory (cudaHostAllocWriteCombined) and
with the rank's threads bound to its own share of the host cores.
"""
import ctypes as C
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
N, PIECE, REPS = 1 << 30, 1 << 28, 8


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(copy):
    for _ in range(2):
        copy()
    barrier()
    t = time.perf_counter()
    for _ in range(REPS):
        copy()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    barrier()
    return REPS * PIECE / float(dt.item()) / 1e9  # per rank, at the pace of the slowest rank


out = {"n_gpus": world}
y = torch.empty(PIECE, dtype=torch.uint8, device=dev)
x = torch.empty(N, dtype=torch.uint8).pin_memory()
pieces = [x[i * PIECE:(i + 1) * PIECE] for i in range(N // PIECE)]
k = [0]


def copy_pinned():
    y.copy_(pieces[k[0] % len(pieces)], non_blocking=True)
    k[0] += 1


out["pinned_GBs_per_gpu"] = measure(copy_pinned)
# the same with this rank bound to its own slice of the host cores
try:
    cores = sorted(os.sched_getaffinity(0))
    share = max(1, len(cores) // world)
    os.sched_setaffinity(0, set(cores[local * share:(local + 1) * share]) or set(cores))
    out["pinned_affinity_GBs_per_gpu"] = measure(copy_pinned)
    os.sched_setaffinity(0, set(cores))
except Exception as e:  # noqa: BLE001
    out["pinned_affinity_GBs_per_gpu"] = repr(e)
# write-combined host memory
try:
    rt = C.CDLL("libcudart.so.12")
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(4)) == 0  # cudaHostAllocWriteCombined
    C.memset(p, 1, N)
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    st = torch.cuda.current_stream().cuda_stream

    def copy_wc():
        off = (k[0] % (N // PIECE)) * PIECE
        rt.cudaMemcpyAsync(y.data_ptr(), p.value + off, PIECE, 1, st)
        k[0] += 1

    out["write_combined_GBs_per_gpu"] = measure(copy_wc)
    rt.cudaFreeHost(p)
except Exception as e:  # noqa: BLE001
    out["write_combined_GBs_per_gpu"] = repr(e)
for key in list(out):
    if key.endswith("_per_gpu") and isinstance(out[key], float):
        out[key.replace("_per_gpu", "_aggregate")] = out[key] * world
if rank == 0:
    try:
        out["host_cores"] = len(os.sched_getaffinity(0))
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
