import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from vectorindex_b200 import kernels as vk
from vectorindex_b200._lib import KMeansCfg, PQTrainCfg
from vectorindex_b200.index import IVFPQIndex
dev = torch.device("cuda", 0)
for shape, metric in (("embed", "dotProduct"), ("embed", "euclidean"), ("deep", "dotProduct")):
    cfg = dict(n=500_000, d=768, nlist=1024, nprobe=32, m=64, nq=1000, k=10, shape=shape, clusters=1024, metric=metric)
    bench.CHUNK = 500_000
    synth = bench.Synth(cfg, dev)
    x = synth.rows(0, cfg["n"])
    q = synth.queries(cfg["nq"])
    idx = IVFPQIndex(cfg["d"], metric, nlist=cfg["nlist"], nprobe=cfg["nprobe"], m=cfg["m"])
    idx.optimize(x[:131072].contiguous(), KMeansCfg(1024, 6, 1e-4, 42, 0, False, 1), PQTrainCfg(0, 8, 1e-4, 1024, 65536, 42, 0, 0, 1))
    idx.batch_insert(x)
    mi = 1 if metric == "dotProduct" else 0
    gd, gi = vk.flat_search_f32(q, x, 10, mi)
    d, i = idx.batch_search(q, 10)
    d100, i100 = idx.batch_search(q, 100)
    gi, i, i100 = gi.cpu().numpy(), i.cpu().numpy(), i100.cpu().numpy()
    rec = np.mean([len(set(gi[r]) & set(i[r])) / 10 for r in range(len(gi))])
    rec1 = np.mean([gi[r][0] in i[r] for r in range(len(gi))])
    rec100 = np.mean([len(set(gi[r]) & set(i100[r])) / 10 for r in range(len(gi))])
    asg = vk.ivf_assign_f32(x, torch.from_numpy(idx.get_coarse()).cuda()) if mi == 0 else vk.ivf_assign_metric_f32(x, torch.from_numpy(idx.get_coarse()).cuda(), 1, None)
    sizes = torch.bincount(asg.to(torch.int64), minlength=cfg["nlist"]).float()
    print(shape, metric, "recall@10", round(rec, 3), "top1 found", round(rec1, 3), "recall 10@100", round(rec100, 3), "gt d", gd[0].cpu().numpy()[:4], "list size max/mean", sizes.max().item(), sizes.mean().item(), flush=True)
