# Per-warp time stamps of one likelihood launch (option "trace"): where does a small batch lose
# time -- start-up, drain (tail), or imbalance between SMs?
#   python tools/warp_trace.py [cfg] [B] [opt=value ...]
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
case = synth.make_case(cfg)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    m.set_option(k, int(v))
m.set_option("trace", 1)
m.set_option("timing", 1)
th = torch.from_numpy(case.draw_theta(B, seed=1000)).cuda()
out = torch.empty(B, dtype=torch.float64, device='cuda')
for _ in range(5):
    m.log_likelihood_device(th, out=out)
torch.cuda.synchronize()
kms = m.last_kernel_ms()
buf = np.zeros((148 * 32, 4), dtype=np.uint64)
rows = ctypes.c_int32(0)
rc = m._lib.rvl_read_trace(m._h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(buf), ctypes.byref(rows))
assert rc == 0
tr = buf[:rows.value].astype(np.int64)
t0 = tr[:, 0].min()
enter, ready, done, items = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3]
span = done.max()
W = rows.value // 148
print(f"cfg {cfg} B {B} opts {sys.argv[3:]}: kernel {kms*1e3:.1f} us (events), trace span {span/1e3:.1f} us, {rows.value} warps ({W}/block)")
print(f"  block entry spread: max t_enter {enter.max()/1e3:.1f} us; epoch data resident at {np.median(ready)/1e3:.1f} us (median), {ready.max()/1e3:.1f} (max)")
print(f"  items per warp: mean {items.mean():.2f} min {items.min()} max {items.max()}")
q = np.percentile(done, [1, 10, 25, 50, 75, 90, 99, 100]) / 1e3
print("  warp finish time percentiles (us) 1/10/25/50/75/90/99/100: " + " ".join(f"{x:.1f}" for x in q))
busy = (done - ready).sum() / (rows.value * span)
print(f"  warp-time utilisation (sum of ready->done over warps / warps x span): {busy:.3f}")
sm_last = done.reshape(148, W).max(1) / 1e3
sm_med = np.median(done.reshape(148, W), 1) / 1e3
print(f"  per-SM last finish: min {sm_last.min():.1f} median {np.median(sm_last):.1f} max {sm_last.max():.1f} us; per-SM median finish: min {sm_med.min():.1f} max {sm_med.max():.1f}")
# active warps over time
ts = np.linspace(0, span, 21)
act = [(np.sum((ready <= t) & (done > t))) / rows.value for t in ts]
print("  active-warp fraction at 0..100% of the span: " + " ".join(f"{a:.2f}" for a in act))
