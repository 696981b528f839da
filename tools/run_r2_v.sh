# round 2, scratch A/B list of the last GPU calls (U = 2 against the automatic choice)
set -x
python tools/prof_sweep.py 3 524288 2 | tail -1
python tools/prof_sweep.py 3 524288 0 | tail -1
python tools/prof_sweep.py 5 131072 0 | tail -1
