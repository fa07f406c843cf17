"""Writes tests/golden/search_pairs.npz: one synthetic frame pair and what the Python restatement of
ORBmatcher::SearchForInitialization / SearchByProjection(Frame, Frame) (tests/search_cases.py) returns for it.
Run from the repo root:  python tests/golden/make_search_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import search_cases as sc  # noqa: E402

k1, d1, k2, d2 = sc.frame_pair(1234, 500, 520, dup=0.1)
gp = sc.grid_params()
prev = np.stack([k1["x"], k1["y"]], 1)
n, m12, pm = sc.py_search_for_initialization(k1, d1, k2, d2, gp, prev, 100, 0.9, True)
proj, flags, _ = sc.projection_inputs(1234, k1, d1)
rng = np.random.default_rng(99)
ur = np.where(rng.random(len(k2)) < 0.5, k2["x"] - rng.uniform(0, 30, len(k2)), -1).astype(np.float32)
occ = (rng.random(len(k2)) < 0.1).astype(np.uint8)
sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
bounds = np.array([0, 640, 0, 480], np.float32)
pn, asg = sc.py_search_by_projection(k1, k1, proj, flags, d1, k2, d2, ur, occ, gp, sf, bounds, 15.0, 40.0, 0, True)
tri = sc.triangulation_case(77, 160, 170, dup=0.2)
tn, tm12 = sc.py_search_for_triangulation(*tri, True)
np.savez_compressed(os.path.join(HERE, "matcher", "triangulation_pair.npz"), k1=tri[0].view(np.uint8), d1=tri[1], mp1=tri[2], ur1=tri[3],
                    k2=tri[4].view(np.uint8), d2=tri[5], mp2=tri[6], ur2=tri[7], F12=tri[8], epipole=np.array([tri[9], tri[10]], np.float32),
                    sf=tri[11], sigma2=tri[12], n=tn, m12=tm12)
print("triangulation matches", tn)
np.savez_compressed(os.path.join(HERE, "matcher", "search_pairs.npz"), k1=k1.view(np.uint8), k2=k2.view(np.uint8), d1=d1, d2=d2,
                    gp=np.array(gp, np.float32), prev=prev, init_n=n, init_m12=m12, init_prev=pm, proj=proj, flags=flags, ur=ur,
                    occ=occ, sf=sf, bounds=bounds, proj_n=pn, proj_assigned=asg)
print("init matches", n, "projection matches", pn)
