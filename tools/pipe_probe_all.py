"""Prints the measured rate of every pipe probe of libsdorb.so (warp-instructions / clk / SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api

NAMES = {0: "POPC", 1: "VIMNMX3.U16x2 (ALU)", 2: "PRMT (ALU)", 3: "HFMA2.RELU (FMA)", 4: "HADD2 (FMA)",
         5: "8 VIMNMX3 + 8 HFMA2.RELU", 6: "8 VIMNMX3 + 8 HFMA2.RELU + 4 LDS (+ address ALU)", 7: "IMAD.HI.U32"}
ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=1)
for p in sorted(NAMES):
    best = max(ex.pipe_probe(p)[1] for _ in range(3))
    print("pipe %d  %-50s %.3f warp-instr/clk/SM" % (p, NAMES[p], best))
