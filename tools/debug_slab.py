import sys, time
import numpy as np
sys.path.insert(0, ".")
import frictionqpotspringblock_b200 as F
from frictionqpotspringblock_b200.slab import SlabSystem
from tests.test_gpu_slab import params, protocol
case = sys.argv[1]
module, cls, kw = params(case)
t0 = time.time()
ref = getattr(getattr(F, module), cls)(kernel=2, **kw)
ref.u_frame = 0.5
print("ref minimise", ref.minimise(), ref.inc, ref.step_count, time.time() - t0, flush=True)
s = SlabSystem(module, cls, halo=8, **kw)
s.u_frame = 0.5
from frictionqpotspringblock_b200 import slab as sl
ring = sl.StopList(10)
for b in range(40):
    log = s._logged(8)
    print(b, log[:, :2].tolist(), flush=True)
    st = sl.first_stop(log, ring, 1e-5)
    s.exchange()
    if st:
        print("stop", b, st); break
