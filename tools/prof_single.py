"""Single-frame latency driver (config C1 shape): one ORBextractor call on a 640x480 frame + one SearchForInitialization, through
the Python mirror of the C-ABI.  `ncu --profile-from-start off` captures exactly one extraction + one match; prints wall times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher

h, w = 480, 640
a, b = synth.frame_pair_in_time(h, w, 1001)
ex = ORBextractor(1000, 1.2, 8, 15, 5)
m = ORBmatcher(0.9, True)
for _ in range(5):
    ka, da = ex(a); kb, db = ex(b)
F1, F2 = Frame.front(ka, da, w, h), Frame.front(kb, db, w, h)
pm = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1), np.float32)
for _ in range(3):
    m.SearchForInitialization(F1, F2, pm.copy(), 100)
ts = []
for _ in range(30):
    t = time.perf_counter(); ex(a); ts.append(time.perf_counter() - t)
tm = []
for _ in range(30):
    p = pm.copy(); t = time.perf_counter(); m.SearchForInitialization(F1, F2, p, 100); tm.append(time.perf_counter() - t)
print(f"extract {np.median(ts) * 1e6:.1f} us, SearchForInitialization {np.median(tm) * 1e6:.1f} us (python mirror, wall)")
torch.cuda.cudart().cudaProfilerStart()
ex(a)
m.SearchForInitialization(F1, F2, pm.copy(), 100)
torch.cuda.cudart().cudaProfilerStop()
