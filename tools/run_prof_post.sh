python tools/fip_bench.py 200000 3 > gpurun_out/fip_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:fip_ranges -s 1 -c 1 -f -o gpurun_out/prof_r1c_fip python tools/fip_bench.py 200000 3 > gpurun_out/ncu_fip.log 2>&1
python - <<'PY' > gpurun_out/order_plain.log 2>&1
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from evidence_b200 import fip
from test_fip import _posterior
names, s = _posterior(1, 400000, 4)
for _ in range(3): fip.order_planets(s, names, 4)
print(fip.order_planets.last_kernel_ms)
PY
cat > /tmp/order_drv.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from evidence_b200 import fip
from test_fip import _posterior
names, s = _posterior(1, 400000, 4)
for _ in range(3): fip.order_planets(s, names, 4)
PY
ncu --set full --clock-control none -k regex:order_planets -s 1 -c 1 -f -o gpurun_out/prof_r1c_order python /tmp/order_drv.py > gpurun_out/ncu_order.log 2>&1
tail -2 gpurun_out/fip_plain.log; cat gpurun_out/order_plain.log
