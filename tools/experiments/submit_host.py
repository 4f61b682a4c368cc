"""Host enqueue cost against total time per submit (device memspace, FLAG_ASYNC, 4096 mono x 160 frames)."""
import sys, time
sys.path.insert(0, ".")
import ctypes as C
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib
from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer
lib = _lib.load()
n, ch, call = 4096, 1, 160
b = FirBatch(n, ch, 16000, 48000, Latency(1), Attenuation.Db90, kernel=Kernel.EXACT)
bso = b.buffer_size_output()
d_in = DeviceBuffer(0, n * call * ch); d_out = DeviceBuffer(0, n * bso)
ins = (C.c_void_p * n)(*[d_in.ptr + 4 * s * call * ch for s in range(n)])
outs = (C.c_void_p * n)(*[d_out.ptr + 4 * s * bso for s in range(n)])
il = (C.c_size_t * n)(*([call * ch] * n)); ol = (C.c_size_t * n)(*([bso] * n))
for reps in (50, 500, 500):
    b.sync(); t0 = time.perf_counter()
    for _ in range(reps):
        b.submit_ptrs(ins, il, outs, ol, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    t1 = time.perf_counter(); b.sync(); t2 = time.perf_counter()
    print(f"reps {reps}: host enqueue per submit {(t1 - t0) / reps * 1e6:.1f} us, total per submit {(t2 - t0) / reps * 1e6:.1f} us")
