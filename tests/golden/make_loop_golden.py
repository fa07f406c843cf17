"""Writes tests/golden/matcher/local_map_loop.npz: synthetic keyframes / map points and what the Python restatements
(tests/search_cases.py) of the local-mapping and loop-closing matchers return for them -- ORBmatcher::SearchByProjection(Frame,
vpMapPoints, th), SearchByPoints, the keypoint search of both Fuse overloads, SearchBySim3 and SearchByProjection(KeyFrame, Scw,
vpPoints, vpMatched, th) (/root/reference/src/ORBmatcher.cc:43-119, 146-254, 535-586, 682-708, 734-944, 1209-1304).
Run from the repo root:  python tests/golden/make_loop_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import search_cases as sc  # noqa: E402

gp = sc.grid_params()
sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
inv = (np.float32(1) / (sf * sf)).astype(np.float32)
out = dict(gp=np.array(gp, np.float32), sf=sf, inv_sigma2=inv)

# keyframe + map points shared by the local-map search, both Fuse searches and the Sim3 projection search
_, _, kf, df = sc.frame_pair(4321, 10, 420, dup=0.15, level0=0.3)
proj, lvl, fl, ur = sc.fuse_inputs(4321, kf, 380, stereo=True)
dmp = sc.fuse_descriptors(4321, df, len(kf), 380)
rng = np.random.default_rng(4321)
occ = (rng.random(len(kf)) < 0.12).astype(np.uint8)
out.update(kf=kf.view(np.uint8), df=df, proj=proj, level=lvl, flags=fl, ur=ur, dmp=dmp, occ=occ)
bi, bd = sc.py_fuse_search(proj, lvl, fl, dmp, kf, df, ur, gp, sf, inv, 3.0)
out.update(fuse_idx=bi, fuse_dist=bd)
bi2, bd2 = sc.py_fuse_search(proj, lvl, fl, dmp, kf, df, None, gp, sf, None, 4.0, False, sc.TH_LOW)
out.update(fuse_sim3_idx=bi2, fuse_sim3_dist=bd2)
n, asg = sc.py_search_by_projection_sim3(proj, lvl, fl, dmp, kf, df, occ, gp, sf, 10)
out.update(proj_sim3_n=n, proj_sim3_assigned=asg)
print("fuse", int((bi >= 0).sum()), "fuse sim3", int((bi2 >= 0).sum()), "projection sim3", n)

mp = sc.map_point_inputs(4321, kf, df, 380)
flags_lm = mp[3]
n, asg = sc.py_search_map_points(mp[0], mp[1], mp[2], flags_lm, mp[4], kf, df, ur, occ, gp, sf, 3.0, 0.8)
out.update(lm_proj=mp[0], lm_view_cos=mp[1], lm_level=mp[2], lm_flags=flags_lm, lm_dmp=mp[4], lm_n=n, lm_assigned=asg)
print("local map", n)

s1, s2 = sc.sim3_case(4321, 400, 420, dup=0.1)
n, m12, m1, m2 = sc.py_search_by_sim3(s1, s2, gp, sf, 7.5)
for tag, s in (("a", s1), ("b", s2)):
    out.update({"sim3_proj_" + tag: s[0], "sim3_level_" + tag: s[1], "sim3_flags_" + tag: s[2], "sim3_dmp_" + tag: s[3],
                "sim3_k_" + tag: s[4].view(np.uint8), "sim3_d_" + tag: s[5]})
out.update(sim3_n=n, sim3_m12=m12, sim3_m1=m1, sim3_m2=m2)
print("sim3", n)

k1, d1, k2, d2 = sc.frame_pair(4322, 300, 320, dup=0.3, flips=25)
v1, v2 = (rng.random(len(k1)) < 0.8).astype(np.uint8), (rng.random(len(k2)) < 0.8).astype(np.uint8)
n, m12 = sc.py_search_by_points(k1, d1, v1, k2, d2, v2, 0.75, True)
out.update(bp_k1=k1.view(np.uint8), bp_d1=d1, bp_v1=v1, bp_k2=k2.view(np.uint8), bp_d2=d2, bp_v2=v2, bp_n=n, bp_m12=m12)
print("by points", n)
np.savez_compressed(os.path.join(HERE, "matcher", "local_map_loop.npz"), **out)
