import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
img = synth.frame(200, 300, 5)
g = ORBextractor(500, 1.2, 5, 15, 5)
k, d = g(img)
print("n", len(k))
