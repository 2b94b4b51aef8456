import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
from vfm_registration_b200 import synth
s = synth.make_pair(4, 200_000, 20_000, 768, sigma_f=0.02)
a, b = torch.from_numpy(s["scan_feat"]).cuda(), torch.from_numpy(s["map_feat"]).cuda()
m = v.match_nn(a, b)
