import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import scipy.fft as sfft
from fluidsim_b200.fft import FFT3DWithB200, FFT2DWithB200
torch.manual_seed(0)
def t3(shape):
    o = FFT3DWithB200(*shape)
    x = torch.rand(shape, dtype=torch.float64, device="cuda") - 0.5
    k = o.fft(x)
    kr = sfft.rfftn(x.cpu().numpy()) / np.prod(shape)
    e1 = np.abs(k.cpu().numpy() - kr).max() / np.abs(kr).max()
    x2 = o.ifft(k)
    e2 = float((x2 - x).abs().max())
    kk = k.clone()
    x3 = o.create_arrayX(); o.ifft_as_arg_destroy(kk, x3)
    e3 = float((x3 - x).abs().max())
    print(shape, "fast" if o.plan.is_fast else "generic", "fwd err %.2e  roundtrip %.2e destroy %.2e" % (e1, e2, e3), flush=True)
def t2(shape):
    o = FFT2DWithB200(*shape)
    x = torch.rand(shape, dtype=torch.float64, device="cuda") - 0.5
    k = o.fft(x)
    kr = sfft.rfft2(x.cpu().numpy()) / np.prod(shape)
    e1 = np.abs(k.cpu().numpy() - kr).max() / np.abs(kr).max()
    x2 = o.ifft(k)
    e2 = float((x2 - x).abs().max())
    print(shape, "fast" if o.plan.is_fast else "generic", "fwd err %.2e  roundtrip %.2e" % (e1, e2), flush=True)
for s in [(8,8,8),(16,16,16),(8,16,32),(32,32,32),(64,64,64),(128,128,128),(256,256,256),(4,11,16),(8,12,16),(10,15,20),(6,6,6),(16,8,10)]:
    try: t3(s)
    except Exception as ex: print(s, "FAILED", ex, flush=True)
for s in [(8,8),(64,32),(256,256),(512,1024),(2048,2048),(24,32),(11,16),(15,9)]:
    try: t2(s)
    except Exception as ex: print(s, "FAILED", ex, flush=True)
for n in (512,):
    o = FFT3DWithB200(n,n,n)
    x = torch.rand((n,n,n), dtype=torch.float64, device="cuda") - 0.5
    k = o.fft(x); x2 = o.ifft(k); print(n, "roundtrip", float((x2-x).abs().max()))
    torch.cuda.synchronize(); t0=time.time()
    for i in range(5): o.fft_as_arg(x, k)
    torch.cuda.synchronize(); dt=(time.time()-t0)/5
    print("512^3 r2c: %.2f ms  (3-pass traffic %.1f GB -> %.0f GB/s)" % (dt*1e3, 6*k.numel()*16/1e9, 6*k.numel()*16/dt/1e9))
