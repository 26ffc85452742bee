import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_models as T
from spgemm_gnn_b200 import models
from spgemm_gnn_b200.train import train_epochs, train_epochs_graphed
import maxk_kernels as mk
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "gcn"
g, x, y, mask = T._task(n=3000, avg_deg=int(sys.argv[2]) if len(sys.argv) > 2 else 120)
gc, xc, yc, mc = g.to("cuda"), x.cuda(), y.cuda(), mask.cuda()
torch.manual_seed(3)
m0 = models.MODELS[name](64, 256, 3, 7, maxk=32, feat_drop=0.0, norm=True).cuda()
l1, _ = train_epochs(copy.deepcopy(m0), gc, xc, yc, mc, 8, lr=0.01)
print("eager      ", ["%.5f" % v for v in l1])
for wu in (8, 3, 1):
    l2, _ = train_epochs_graphed(copy.deepcopy(m0), gc, xc, yc, mc, 8, lr=0.01, warmup=wu)
    print(f"graph wu={wu} ", ["%.5f" % v for v in l2])
mk.set_banked(False)
l2, _ = train_epochs_graphed(copy.deepcopy(m0), gc, xc, yc, mc, 8, lr=0.01, warmup=3)
print("graph wu=3 nobank", ["%.5f" % v for v in l2])
