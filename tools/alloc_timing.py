"""How expensive are cudaMalloc / cudaFree for the state of one denoise call? (tools/, not product code)"""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytvdn_b200 import _lib
lib = _lib.load()
GB = 1 << 30
def t(f):
    t0 = time.perf_counter(); r = f(); lib.cytvdn_stream_synchronize(None); return r, (time.perf_counter() - t0) * 1e3
lib.cytvdn_set_device(0)
p0 = C.c_void_p(); lib.cytvdn_malloc(C.byref(p0), 1 << 20)   # context creation
for n, sz in ((20, 4 * GB), (1, 80 * GB), (20, 4 * GB)):
    ptrs = []
    def alloc():
        for _ in range(n):
            p = C.c_void_p(); _lib.check(lib.cytvdn_malloc(C.byref(p), sz)); ptrs.append(p)
    _, ta = t(alloc)
    def touch():
        for p in ptrs: lib.cytvdn_memset(p, 0, sz, None)
    _, tm = t(touch)
    def free():
        for p in ptrs: lib.cytvdn_free(p)
    _, tf = t(free)
    print(f"{n:3d} x {sz/GB:5.1f} GB: malloc {ta:8.1f} ms   memset {tm:8.1f} ms   free {tf:8.1f} ms")
