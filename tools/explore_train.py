"""Build-side exploration (torch only): how much training does the probing MLP need on the bench workload
to beat distance-rank IVF? Prints recall@10 vs model top-n and vs threshold for a few training budgets."""
import sys, time
import torch
sys.path.insert(0, ".")
from lira_ann_search_b200.utils import Kmeans
from lira_ann_search_b200.model_probing import MLP_2_Input

dev = "cuda:0"
N, d, Q, B, k = 1_000_000, 128, 10_000, 1024, 10
g = torch.Generator(device=dev).manual_seed(43 * 1_000_003)
centres = torch.randn(4096, d, generator=g, device=dev)
w = torch.exp(0.5 * torch.randn(4096, generator=g, device=dev))


def draw(m):
    c = torch.multinomial(w, m, replacement=True, generator=g)
    return torch.clamp(torch.round(16 * (centres[c] + 0.8 * torch.randn(m, d, generator=g, device=dev)) + 100), 0, 255)


x_d, x_q = draw(N), draw(Q)
bn = (x_d * x_d).sum(1)


def knn(qs, kk):
    out = torch.empty(qs.shape[0], kk, dtype=torch.int64, device=dev)
    for a in range(0, qs.shape[0], 2048):
        out[a:a + 2048] = (bn[None, :] - 2.0 * qs[a:a + 2048] @ x_d.T).topk(kk, largest=False).indices
    return out


gt = knn(x_q, k)
km = Kmeans(d, B, niter=20, device=dev).train(x_d.cpu().numpy())
cent = torch.as_tensor(km.centroids, device=dev)
assign = Kmeans.assign(x_d, cent)
sizes = torch.bincount(assign, minlength=B).float()
f = torch.cdist(x_d[:200000], cent).double()
mean, scale = f.mean(0).float(), f.std(0).float()
gt_b = assign[gt]
fq = (torch.cdist(x_q, cent) - mean) / scale


def report(scores, tag):
    line = tag
    rank = scores.argsort(1, descending=True)
    for npb in (1, 2, 4, 8, 16, 32):
        probed = rank[:, :npb]
        line += f" top{npb}:{(gt_b[:, :, None] == probed[:, None, :]).any(-1).float().mean().item():.3f}"
    for thr in (0.5, 0.3, 0.2, 0.1, 0.05, 0.02):
        m = scores > thr
        hit = torch.gather(m, 1, gt_b).float().mean().item()
        line += f" | t{thr}: np{m.sum(1).float().mean().item():.1f} r{hit:.3f} cmp{(m.float() @ sizes).mean().item():.0f}"
    print(line, flush=True)


report(-fq, "distance-rank")
for n_tr, epochs, bs, lr in [(100_000, 8, 512, 1e-3), (300_000, 20, 1024, 2e-3), (300_000, 40, 512, 1e-3), (1_000_000, 10, 1024, 2e-3)]:
    t0 = time.time()
    tr = torch.randperm(N, device=dev)[:n_tr]
    knn_tr = knn(x_d[tr], k + 1)[:, 1:]
    labels = torch.zeros(n_tr, B, device=dev)
    labels.scatter_(1, assign[knn_tr], 1.0)
    xf, xv = (torch.cdist(x_d[tr], cent) - mean) / scale, x_d[tr]
    torch.manual_seed(43)
    model = MLP_2_Input(B, d, B).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    crit = torch.nn.BCELoss()
    for ep in range(epochs):
        perm = torch.randperm(n_tr, device=dev)
        for a in range(0, n_tr, bs):
            idx = perm[a:a + bs]
            opt.zero_grad()
            loss = crit(model(xf[idx], xv[idx]), labels[idx])
            loss.backward()
            opt.step()
    model.eval()
    with torch.no_grad():
        s = model(fq, x_q)
    report(s, f"n_tr={n_tr} ep={epochs} bs={bs} lr={lr} loss={loss.item():.4f} [{time.time() - t0:.0f}s]")
