python - <<'PY'
import sys; sys.argv=['x','2','4096','300']
exec(open('tools/sched_sweep.py').read().split("run()\nrun(sched=0)")[0])
run(); run(warps=32); run(warps=32, phase_items=150); run(warps=32, phase_items=250); run(warps=32, max_split=16); run(warps=30); run()
PY
for w in 28 32; do python tools/prof_sweep.py 5 32768 2 warps=$w | tail -1 | cut -c1-140; python tools/prof_sweep.py 2 65536 2 warps=$w | tail -1 | cut -c1-140; python tools/prof_sweep.py 1 65536 2 warps=$w | tail -1 | cut -c1-140; done
