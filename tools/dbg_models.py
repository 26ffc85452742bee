import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, torch.nn.functional as F
import test_gpu_models as T
from oracle import ref_torch
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "gin"
g, x, y, mask = T._task(n=2000, avg_deg=20)
ref, ours = T._pair(name, 64, 256, 3, 7, 32, norm=True)
adj = ref_torch.csr_matrix(g.indptr, g.indices, g.edge_weights(T.KIND[name]).double(), g.num_src)
opt = torch.optim.Adam(ref.parameters(), lr=0.01)
gc = g.to("cuda"); xc = x.cuda(); yc = y.cuda(); mc = mask.cuda()
def sync():
    sd = {}
    for key, v in ref.state_dict().items():
        key = key.replace("gcn_bias.", "gcnlayers.").replace("eps.", "gcnlayers.")
        if name == "gcn" and key.startswith("gcnlayers.") and key.count(".") == 1: key += ".bias"
        if name == "gin" and key.startswith("gcnlayers.") and key.count(".") == 1: key += ".eps"
        sd[key] = v.float().cuda()
    ours.load_state_dict(sd)
for ep in range(30):
    loss = F.cross_entropy(ref(adj, x.double())[mask], y[mask])
    opt.zero_grad(); loss.backward()
    sync(); ours.zero_grad()
    lo = F.cross_entropy(ours(gc, xc)[mc], yc[mc]); lo.backward()
    gerr = 0
    errs = []
    rg = dict(ref.named_parameters())
    for pn, p in ours.named_parameters():
        rn = pn
        if pn.startswith("gcnlayers."):
            rn = ("gcn_bias." if name == "gcn" else "eps.") + pn.split(".")[1]
        rp = rg[rn]
        e = float((p.grad.cpu().double() - rp.grad).abs().max() / (rp.grad.abs().max() + 1e-30))
        errs.append((e, pn, float(rp.grad.abs().max())))
        gerr = max(gerr, e)
    errs.sort(reverse=True)
    if gerr > 5e-4: print("   worst:", errs[:4])
    print(ep, float(loss), float(lo), abs(float(lo) - float(loss)) / float(loss), "max rel grad err", gerr, "eps", [float(e) for e in ref.eps] if name == "gin" else "")
    opt.step()
