python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/warp_trace.py 2 65536 2>/dev/null | head -1
for o in "sched=0" "sched=0 setup_items=0" "sched=0 prepare=1" "" "max_split=8 phase_items=200" "max_split=8 phase_items=150" "max_split=16 phase_items=150" "max_split=16 phase_items=200" "max_split=8 phase_items=300" "max_split=4 phase_items=300"; do python tools/warp_trace.py 2 4096 $o | egrep "^cfg|utilisation|active-warp"; done
python tools/warp_trace.py 2 65536 2>/dev/null | head -1
python tools/warp_trace.py 3 65536 2>/dev/null | head -1
python tools/warp_trace.py 5 32768 2>/dev/null | head -1
