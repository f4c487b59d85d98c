"""ncu target: a few ADP training steps (default precision bf16x3) on a smaller batch so that `--set full` stays short.
usage: python scripts/probes/one_step.py [precision] [crystals] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import cartnet_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["adp_train"]
hb = bench.make_host_batch(bench.rank_structures("adp", count, 2, 0, 1, dev), 2, dev)
b = bench.shallow(hb).to(dev)
torch.manual_seed(0)
model = cartnet_b200.CartNet(256, 64, 4, precision=prec).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    pred, true = model(bench.shallow(b))
    loss = cartnet_b200.compute_loss(pred, true)[0]
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
