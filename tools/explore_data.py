"""Build-side exploration (torch only): how hard is a synthetic SIFT1M-shape mixture for IVF probing?
Prints distance-rank IVF recall@10 vs nprobe and list-size statistics for a few generator settings."""
import sys, time
import torch
sys.path.insert(0, ".")
from lira_ann_search_b200.utils import Kmeans

dev = "cuda:0"
N, d, Q, B, k = 1_000_000, 128, 10_000, 1024, 10


def gen(kind, sigma, ncomp, seed=43):
    g = torch.Generator(device=dev).manual_seed(seed)
    centres = torch.randn(ncomp, d, generator=g, device=dev)
    w = torch.exp(0.5 * torch.randn(ncomp, generator=g, device=dev))
    if kind == "lowrank":
        basis = torch.randn(ncomp, 12, d, generator=g, device=dev) / 12 ** 0.5

    def draw(m):
        c = torch.multinomial(w, m, replacement=True, generator=g)
        if kind == "iso":
            x = centres[c] + sigma * torch.randn(m, d, generator=g, device=dev)
        else:
            z = torch.randn(m, 12, generator=g, device=dev)
            x = centres[c] + sigma * torch.einsum("mr,mrd->md", z, basis[c]) + 0.1 * torch.randn(m, d, generator=g, device=dev)
        return torch.clamp(torch.round(16 * x + 100), 0, 255)
    return draw(N), draw(Q)


def knn(x_d, qs, kk):
    out = torch.empty(qs.shape[0], kk, dtype=torch.int64, device=dev)
    bn = (x_d * x_d).sum(1)
    for a in range(0, qs.shape[0], 2048):
        out[a:a + 2048] = (bn[None, :] - 2.0 * qs[a:a + 2048] @ x_d.T).topk(kk, largest=False).indices
    return out


for kind, sigma, ncomp in [("iso", 0.8, 4096), ("iso", 0.5, 4096), ("iso", 0.35, 4096), ("lowrank", 1.5, 4096), ("lowrank", 2.5, 1024), ("iso", 1.0, 256)]:
    t0 = time.time()
    x_d, x_q = gen(kind, sigma, ncomp)
    gt = knn(x_d, x_q, k)
    km = Kmeans(d, B, niter=20, device=dev).train(x_d.cpu().numpy())
    cent = torch.as_tensor(km.centroids, device=dev)
    assign = Kmeans.assign(x_d, cent)
    sizes = torch.bincount(assign, minlength=B).float()
    rank = torch.cdist(x_q, cent).argsort(1)
    gt_b = assign[gt]  # [Q,k]
    line = f"{kind} sigma={sigma} ncomp={ncomp}: sizes min/med/mean/max={int(sizes.min())}/{int(sizes.median())}/{int(sizes.mean())}/{int(sizes.max())} |"
    for npb in (1, 2, 4, 8, 16, 32, 64):
        probed = rank[:, :npb]
        hit = (gt_b[:, :, None] == probed[:, None, :]).any(-1).float().mean().item()
        cmp_ = sizes[probed].sum(1).mean().item()
        line += f" np{npb}: {hit:.3f} ({cmp_:.0f})"
    print(line, f"[{time.time() - t0:.1f}s]", flush=True)
