import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
n, m, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (10000, 50000, 384)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn(n, d, device="cuda", generator=g); b = torch.randn(m, d, device="cuda", generator=g)
for _ in range(4):
    v.match_nn(a, b, algo="tc")
torch.cuda.synchronize()
