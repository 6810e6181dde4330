import sys; sys.path.insert(0,'.')
import ctypes as C, torch
from pygeostatistics_b200 import _native
t=C.c_double(); m=C.c_double()
torch.cuda.set_device(0)
_native.check(_native.lib().k3d_measure_fp64_peak(C.byref(t), C.byref(m), None))
print("peak", t.value)
