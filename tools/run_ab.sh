python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
python tools/prof_sweep.py 3 65536 | tail -1
python tools/prof_sweep.py 2 65536 | tail -1
done
cp evidence_b200/librvlnl.so /tmp/keep.so; cp evidence_b200/alt_nr2.so evidence_b200/librvlnl.so
for rep in 1 2; do
python tools/prof_sweep.py 3 65536 | tail -1
python tools/prof_sweep.py 2 65536 | tail -1
done
cp /tmp/keep.so evidence_b200/librvlnl.so
python tools/prof_sweep.py 3 65536 | tail -1
