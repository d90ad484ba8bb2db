"""Debug (build with DBH_EXTRA_NVCC_FLAGS=-DDBH_TRACE): print the exact stage's steps for one candidate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import drake_b200 as db
from drake_b200 import workloads as W
from oracle import ho
ctx = db.Context(0)
wl = W.config4(3)
batch = wl.register(ctx)
batch.set_verify(True)
batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
