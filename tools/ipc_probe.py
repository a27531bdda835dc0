"""Does CUDA IPC work between the ranks of this box?  torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from cytvdn_b200 import _lib
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
lib = _lib.load()
n = 1 << 20
p = C.c_void_p()
_lib.check(lib.cytvdn_malloc(C.byref(p), n * 4))
src = torch.full((n,), float(rank + 1), device=dev)
_lib.check(lib.cytvdn_memcpy(p, src.data_ptr(), n * 4, None))
_lib.check(lib.cytvdn_stream_synchronize(None))
h = (C.c_ubyte * 64)()
_lib.check(lib.cytvdn_ipc_get_handle(p, h))
handles = [None] * world
dist.all_gather_object(handles, bytes(h))
peer = (rank + 1) % world
ph = (C.c_ubyte * 64)(*handles[peer])
q = C.c_void_p()
_lib.check(lib.cytvdn_ipc_open(ph, C.byref(q)))
dst = torch.zeros(n, device=dev)
_lib.check(lib.cytvdn_memcpy(dst.data_ptr(), q, n * 4, None))
_lib.check(lib.cytvdn_stream_synchronize(None))
print(f"rank {rank}: read {float(dst[0])} .. {float(dst[-1])} from rank {peer} (expected {peer + 1}.0)", flush=True)
dist.barrier()
_lib.check(lib.cytvdn_ipc_close(q))
dist.destroy_process_group()
