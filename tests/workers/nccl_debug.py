"""Stage-by-stage NCCL bring-up check (2 ranks): init, one send/recv, grouped ring exchange."""
import os, sys, ctypes
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpic_b200 import _lib
from fbpic_b200._lib import DeviceArray, call
from fbpic_b200.boundaries import world

dist.init_process_group('gloo')
rank, size = dist.get_rank(), dist.get_world_size()
ctx = _lib.context()
print('rank', rank, 'device', ctx.device, flush=True)
a = DeviceArray.from_numpy(np.full(1 << 16, float(rank + 1)))
b = DeviceArray.zeros(1 << 16, np.float64)
world().init_nccl()
call.b2_device_sync()
print('rank', rank, 'nccl init ok', flush=True)
peer = 1 - rank
call.b2_nccl_group_start()
call.b2_nccl_send(ctx.handle, a.ptr, a.nbytes, peer, None)
call.b2_nccl_recv(ctx.handle, b.ptr, b.nbytes, peer, None)
call.b2_nccl_group_end()
call.b2_device_sync()
assert np.all(b.get() == peer + 1), b.get()[:4]
print('rank', rank, 'sendrecv ok', flush=True)
# sub-array views, 4 ops to the same peer
big = DeviceArray.from_numpy((np.arange(40 * 8, dtype=np.float64) + 1000 * rank).reshape(40, 8).astype(np.complex128))
call.b2_nccl_group_start()
call.b2_nccl_send(ctx.handle, big[4:8].ptr, 4 * 8 * 16, peer, None)
call.b2_nccl_send(ctx.handle, big[32:36].ptr, 4 * 8 * 16, peer, None)
call.b2_nccl_recv(ctx.handle, big[36:40].ptr, 4 * 8 * 16, peer, None)
call.b2_nccl_recv(ctx.handle, big[0:4].ptr, 4 * 8 * 16, peer, None)
call.b2_nccl_group_end()
call.b2_device_sync()
h = big.get().real
ref = (np.arange(40 * 8, dtype=np.float64) + 1000 * peer).reshape(40, 8)
assert np.array_equal(h[36:40], ref[4:8]) and np.array_equal(h[0:4], ref[32:36])
print('rank', rank, 'ring slabs ok', flush=True)
dist.barrier()
if rank == 0:
    print('NCCL_DEBUG_OK')
