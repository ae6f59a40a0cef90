"""Per-CTA phase timing of the last tn_gemm launch of a TimeNet forward (bring-up).  gpurun -- python tools/tn_stamps.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dimo_b200 import _lib
from dimo_b200.deform import TimeNet
torch.manual_seed(0)
G, M, L = int(os.environ.get("TN_G", "8")), 512, 32
net = TimeNet(latent_code_dim=L).cuda()
pts = torch.rand(M, 3, device="cuda") - 0.5; times = torch.rand(G, device="cuda"); lat = torch.randn(G, L, device="cuda")
buf = torch.zeros(32 * 128 * 8, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net.forward_batched(pts, times, lat)
    torch.cuda.synchronize()
    _lib.lib().dimo_timenet_debug_stamps(buf.data_ptr())
    net.forward_batched(pts, times, lat)
    torch.cuda.synchronize()
    _lib.lib().dimo_timenet_debug_stamps(None)
T = buf.view(32, 128, 8).cpu().double()
names = ["start", "setup", "stage0", "mma issued", "prefetch", "acc ready", "epi done", "stores read"]
base = T[0, :, 0].min()
for l in range(10):
    t = T[l]
    rel = (t - t[:, 0].min()) / 1000.0
    print(f"gemm {l}: starts at {(t[:, 0].min() - base) / 1000.0:7.2f} us | " + "  ".join(f"{n} {rel[:, i].median():5.2f}" for i, n in enumerate(names)) + f" | last CTA ends {rel[:, 7].max():5.2f}")
