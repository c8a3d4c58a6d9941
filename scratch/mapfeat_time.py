import sys, torch
sys.path.insert(0, ".")
from ivlnce_b200.map_encoder import MapFeatures
dev = torch.device("cuda:0")
occ = torch.randint(0, 2, (16, 128, 128), dtype=torch.uint8, device=dev); sem = torch.randint(0, 13, (16, 128, 128), dtype=torch.uint8, device=dev)
mf = MapFeatures(13)
for _ in range(30): out = mf({"occupancy_map": occ, "semantic_map": sem})
torch.cuda.synchronize()
