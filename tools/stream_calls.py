"""Small-chunk streaming throughput (BASELINE config 3 pattern): every stream gets one resample() call per
submit.  usage: python tools/stream_calls.py [streams] [channels] [in_hz] [out_hz] [latency] [call_frames] [kernel]"""
import sys
import time
sys.path.insert(0, ".")
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib
from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer

a = sys.argv[1:]
n = int(a[0]) if len(a) > 0 else 4096
ch = int(a[1]) if len(a) > 1 else 1
in_hz = int(a[2]) if len(a) > 2 else 16000
out_hz = int(a[3]) if len(a) > 3 else 48000
lat = int(a[4]) if len(a) > 4 else 1
call = int(a[5]) if len(a) > 5 else 160
kern = Kernel[a[6].upper()] if len(a) > 6 else Kernel.AUTO
lib = _lib.load()
b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kern)
bso = b.buffer_size_output()
ostr = (bso + 3) & ~3          # 16-byte aligned rows: the tensor kernel's TMA stores need them
d_in = DeviceBuffer(0, n * call * ch)
d_out = DeviceBuffer(0, n * ostr)
assert lib.rsb_fill_synthetic(0, d_in.ptr, 0, n, call, ch, in_hz, 1) == 0
import ctypes as C
ins = (C.c_void_p * n)(*[d_in.ptr + 4 * s * call * ch for s in range(n)])
outs = (C.c_void_p * n)(*[d_out.ptr + 4 * s * ostr for s in range(n)])
in_lens = (C.c_size_t * n)(*([call * ch] * n))
out_lens = (C.c_size_t * n)(*([bso] * n))
for reps in (20, 200):
    b.sync()
    t0 = time.time()
    b.timer_start()
    for _ in range(reps):
        cons, prod = b.submit_ptrs(ins, in_lens, outs, out_lens, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    dev_ms = b.timer_stop()
    b.sync()
    dt = time.time() - t0
print(f"{n} streams x {ch} ch, {in_hz}->{out_hz}, {call}-frame calls, kernel {b.last_kernel().name}: "
      f"{dt / reps * 1e6:.1f} us per submit wall, {dev_ms / reps * 1e3:.1f} us device (events around the loop), "
      f"{n * prod[0] / (dt / reps) / 1e6:.1f} Msamples/s out (wall)")
