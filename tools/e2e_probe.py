# where does the end-to-end time of one host-buffer call go?  (config 2, 4096 theta)
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
case = synth.make_case(2); B = 4096
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
th_host = case.draw_theta(B, seed=1)
th_pin = torch.from_numpy(th_host).pin_memory(); out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
th_dev = th_pin.cuda(); out_dev = torch.empty(B, dtype=torch.float64, device='cuda')
def timeit(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
def a(): m.log_likelihood_batch(th_pin.numpy(), out=out_pin.numpy())
def a2(): m.log_likelihood_batch(th_host)
def b():
    th_dev.copy_(th_pin, non_blocking=True); m.log_likelihood_device(th_dev, out=out_dev); out_pin.copy_(out_dev, non_blocking=True); torch.cuda.synchronize()
def c(): m.log_likelihood_device(th_dev, out=out_dev); torch.cuda.synchronize()
def d(): th_dev.copy_(th_pin, non_blocking=True); torch.cuda.synchronize()
def e(): out_pin.copy_(out_dev, non_blocking=True); torch.cuda.synchronize()
def f(): torch.cuda.synchronize()
print("rvl_loglike pinned host buffers : %.1f us" % timeit(a))
print("rvl_loglike pageable numpy      : %.1f us" % timeit(a2))
print("torch H2D + dev call + D2H+sync : %.1f us" % timeit(b))
print("dev call + sync                 : %.1f us" % timeit(c))
print("H2D 491 KB + sync               : %.1f us" % timeit(d))
print("D2H 32 KB + sync                : %.1f us" % timeit(e))
print("sync only                       : %.1f us" % timeit(f))
m.set_option("zero_copy", 0)
print("rvl_loglike pinned, zero-copy off: %.1f us" % timeit(a))
m.set_option("zero_copy", 1); m.set_option("timing", 1)
print("rvl_loglike pinned, timing on    : %.1f us" % timeit(a))
print("kernel ms", m.last_kernel_ms())
m.set_option("timing", 0)
for zc in (1, 0):
    m.set_option("zero_copy", zc)
    th_np, out_np = th_pin.numpy(), out_pin.numpy()
    for _ in range(10): m.log_likelihood_batch(th_np, out=out_np)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(1000): m.log_likelihood_batch(th_np, out=out_np)
    torch.cuda.synchronize()
    print("bench-style loop, zero_copy=%d: %.1f us per call" % (zc, (time.perf_counter() - t) / 1000 * 1e6))
m.set_option("zero_copy", 1)
import ctypes
dp = ctypes.POINTER(ctypes.c_double)
pa, pb = th_np.ctypes.data_as(dp), out_np.ctypes.data_as(dp)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(1000): m._lib.rvl_loglike(m._h, pa, B, pb)
print("raw ctypes rvl_loglike           : %.1f us per call" % ((time.perf_counter() - t) / 1000 * 1e6))
# the sampler's real calling pattern: pageable numpy arrays, transform then likelihood
m.set_priors(case.priordict)
U = case.draw_unit(B, seed=3)
def g(): th = m.prior_transform_batch(U); return m.log_likelihood_batch(th)
def g2(): return m.transform_loglike_batch(U)
for zc in (1, 0):
    m.set_option("zero_copy", zc)
    print("zero_copy=%d  pageable: loglike %.1f us | transform + loglike %.1f us | fused transform_loglike %.1f us" % (
        zc, timeit(a2), timeit(g), timeit(g2)))
m.set_option("zero_copy", 1)
