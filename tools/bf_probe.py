import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fishbirdeyevisualslam_b200.matcher import ORBmatcher
rng = np.random.default_rng(5)
q = rng.integers(0, 256, (8000, 32), dtype=np.uint8); t = rng.integers(0, 256, (8000, 32), dtype=np.uint8)
m = ORBmatcher(0.9, True)
for _ in range(5): m.BruteForceTop2(q, t)
print("ok")
