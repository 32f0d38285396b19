import torch, sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
wl = dict(bench.WORKLOADS["tiny"])
outs = []
for d in (0, 1):
    dev = torch.device("cuda", d)
    torch.cuda.set_device(dev)
    keys, vals, parent, codes = bench.make_database(torch, dev, bench.DATABASES["tiny"], seed=43)
    b, o = bench.make_reads(torch, dev, wl, codes, 100000, seed=5343)
    outs.append((keys.cpu(), vals.cpu(), codes.cpu(), b.cpu()))
for i, nm in enumerate(("keys", "vals", "codes", "bases")):
    print(nm, torch.equal(outs[0][i], outs[1][i]))
# same device twice
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
b1, _ = bench.make_reads(torch, dev, wl, outs[0][2].to(dev), 100000, seed=5343)
print("repeat on cuda:0", torch.equal(b1.cpu(), outs[0][3]))
