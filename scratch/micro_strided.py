import sys, os, time
sys.path.insert(0, ".")
import torch
from fluidsim_b200.fft import Plan
from fluidsim_b200._lib import call, ptr, stream_ptr
n = int(sys.argv[1]); nf = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pl = Plan(3, (n, n, n), (6.283,)*3)
a = torch.randn((nf,) + pl.shapeK, dtype=torch.complex128, device="cuda")
b = torch.empty_like(a)
F = a[0].numel()*16
def run(axis, d, inplace):
    o = a if inplace else b
    for _ in range(2): call("b2_dev_strided_pass", pl.handle, axis, d, ptr(a), ptr(o), nf, stream_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    R = 5
    for _ in range(R): call("b2_dev_strided_pass", pl.handle, axis, d, ptr(a), ptr(o), nf, stream_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/R
    return ms, 2*nf*F/ms/1e6
for axis in (0, 1):
    for inplace in (False, True):
        ms, gbs = run(axis, -1, inplace)
        print(f"SVAR={os.environ.get('B2_SVAR','0')} n={n} nf={nf} axis={axis} inplace={inplace}: {ms:.3f} ms  {gbs:.0f} GB/s  ({gbs/6546.6:.3f})", flush=True)
