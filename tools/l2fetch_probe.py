"""Does cudaLimitMaxL2FetchGranularity change the random-access kernels?  (ncu: a random 4-byte read pulls ~120 B
from DRAM.)  Usage: python tools/l2fetch_probe.py <bytes: 32|64|128> <script> [args...]  -- sets the limit, then runs the script in-process."""
import ctypes
import runpy
import sys

import torch

torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
val = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
before = val.value
rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(sys.argv[1])))  # cudaLimitMaxL2FetchGranularity = 5
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print(f"[l2fetch] limit before {before}, set rc={rc}, now {val.value}", file=sys.stderr)
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
