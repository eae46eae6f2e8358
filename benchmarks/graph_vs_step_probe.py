import sys, torch
sys.path.insert(0, "/root/repo")
from tests import helpers
from graphslim_b200 import data as gdata
from graphslim_b200.reduction import create_reducer
name = sys.argv[1] if len(sys.argv) > 1 else "mini_sgc2_arxiv"
def run(graphs, **kw):
    args = helpers.case_args(name, device="cuda", save_init=False, progress=False, gemm_precision=1, cuda_graphs=graphs)
    args.epochs = 2
    raw = helpers.case_graph(name)
    helpers.seed_everything(args.seed)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    for k, v in kw.items(): setattr(agent.K, k, v)
    agent.reduce(data, verbose=False)
    torch.cuda.synchronize()
    return data.adj_syn.cpu().clone(), torch.cat([p.reshape(-1) for p in agent.pge.parameters()]).cpu()
runs = {"s1": run(False), "s2": run(False), "s3": run(False), "g1": run(True), "g2": run(True), "g3": run(True),
        "s_nomn": run(False, grouped_mn=False), "g_nomn": run(True, grouped_mn=False),
        "s_nofuse": run(False, pge_fused=False), "g_nofuse": run(True, pge_fused=False)}
ks = list(runs)
for i in range(len(ks)):
    for j in range(i + 1, len(ks)):
        a, b = runs[ks[i]], runs[ks[j]]
        print(f"{ks[i]:9s} {ks[j]:9s} adj {float((a[0]-b[0]).abs().max()):.2e} pge {float((a[1]-b[1]).abs().max()):.2e}")
