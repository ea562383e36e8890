import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bird_scenes as S
from fishbirdeyevisualslam_b200.bird_orb import BirdORB
img, mask, contour = S.bird_image(1), S.bird_mask(1), S.contour_image(1)
orb = BirdORB(2000, 384, 384)
for _ in range(4):
    orb.features(img, mask, contour)
