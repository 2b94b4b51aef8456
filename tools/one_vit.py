import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v
model = sys.argv[1] if len(sys.argv) > 1 else "vitl14"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 48
f = v.ViTFeaturizer(model, seed=1, random_init=True)
imgs = torch.randint(0, 255, (b, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    f.forward(imgs)
torch.cuda.synchronize()
