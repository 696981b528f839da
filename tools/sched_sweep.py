# Launch-plan sweep on the bench workload (config 2, 4096 theta): device time per step (CUDA
# events, L2 flushed between steps) and end-to-end time of the host-buffer call, per option set.
#   python tools/sched_sweep.py [cfg] [B] [steps]
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
case = synth.make_case(cfg)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
th_host = case.draw_theta(B, seed=1000)
th = torch.from_numpy(th_host).cuda()
out = torch.empty(B, dtype=torch.float64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
th_pin = torch.from_numpy(th_host).pin_memory()
out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
ref = None

DEFAULTS = dict(sched=1, phase_items=200, max_split=8, prepare=0, warps=0, slices=0, zero_copy=1,
                items_per_warp=4, min_chunks=8)


def run(**opts):
    global ref
    o = dict(DEFAULTS); o.update(opts)
    for k, v in o.items():
        m.set_option(k, v)
    m.set_option("timing", 1)
    for _ in range(10):
        m.log_likelihood_device(th, out=out)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kms = []
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        m.log_likelihood_device(th, out=out)
        ev[k][1].record()
        kms.append(m.last_kernel_ms())
    torch.cuda.synchronize()
    dev = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    m.set_option("timing", 0)
    got = out.cpu().numpy()
    if ref is None:
        ref = got.copy()
    err = float(np.max(np.abs(got - ref)))
    # end to end through the host-buffer call
    for _ in range(10):
        m.log_likelihood_batch(th_pin.numpy(), out=out_pin.numpy())
    t0 = time.perf_counter()
    for _ in range(steps):
        m.log_likelihood_batch(th_pin.numpy(), out=out_pin.numpy())
    e2e = (time.perf_counter() - t0) / steps * 1e6
    err2 = float(np.max(np.abs(out_pin.numpy() - ref)))
    print("%-60s step %7.1f us (median %7.1f)  kernel %7.1f us  e2e %7.1f us   dlnL %.1e %.1e" % (
        " ".join(f"{k}={v}" for k, v in opts.items()) or "defaults",
        dev.mean(), np.median(dev), np.mean(kms[1:]) * 1e3, e2e, err, err2), flush=True)


run()
run(sched=0)
run(sched=0, prepare=1)
for pi in (50, 75, 150, 200, 300):
    run(phase_items=pi)
for ms in (4, 8, 32):
    run(max_split=ms)
run(max_split=32, phase_items=50)
run(max_split=8, phase_items=200)
run(prepare=1)
run(zero_copy=2)
run(zero_copy=0)
run(warps=24)
run(warps=32)
run(warps=24, phase_items=150)
