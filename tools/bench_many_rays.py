import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import invertsy_b200 as ib, isy_oracle as o, c_oracle
w = ib.Seville2009()
pitch, yaw = o.fibonacci_eye(100000)
omm = o.eye_rotations(pitch, yaw)
ori = o.view_rotations(33.3, omm)
pos = np.array([[5.1, 4.9, 0.01]])
for _ in range(3): c = w(pos, ori)
t0 = time.perf_counter()
for _ in range(10): c = w(pos, ori)
print("world(pos, ori) with 100000 rays: %.2f ms per call" % ((time.perf_counter() - t0) * 100))
hit, _ = w.hits(pos, ori)
yw, pt, _ = ori.as_euler('ZYX').T
ref = c_oracle.world_hits(w.polygons, 2.0, pos[0], pt, yw)[0]
print("mismatches vs oracle:", int((hit != ref).sum()), "of", ref.size)
