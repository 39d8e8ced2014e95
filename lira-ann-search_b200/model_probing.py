"""Probing model: same architecture, parameter names and call shapes as the reference's
model_probing.py (MLP_2_Input :5-39, model_train :41-54, model_evaluate :86-132, model_infer
:135-156), so a state_dict / TorchScript file moves between the two unchanged.

PyTorch is used here for exactly what the north star allows: the MLP forward (and, on the build
side, its training). The query-phase forward used by the search engine is the library's own
(engine.LiraModel -> probe_kernels.cuh); `to_device_model` converts between the two.
"""
from __future__ import annotations

import torch
import torch.nn as nn

SIGMA = 0.5  # probing threshold of model_evaluate / model_infer (model_probing.py:93, 141)


def _two_layer(n_in, n_hidden, n_out, last: nn.Module):
    return nn.Sequential(nn.Linear(n_in, n_hidden), nn.ReLU(), nn.Linear(n_hidden, n_out), last)


class MLP_2_Input(nn.Module):
    """Two-tower MLP: standardised centroid distances [n, B] and the raw vector [n, d] ->
    per-partition probing probability [n, B]."""

    def __init__(self, input_dim1, input_dim2, output_dim):
        super().__init__()
        self.distance_net = _two_layer(input_dim1, 128, 64, nn.ReLU())
        self.vector_net = _two_layer(input_dim2, 128, 64, nn.ReLU())
        self.fc = _two_layer(64 + 64, 128, output_dim, nn.Sigmoid())

    def forward(self, x_dist, x_vec):
        towers = torch.cat((self.distance_net(x_dist), self.vector_net(x_vec)), dim=1)
        return self.fc(towers)


def model_train(model, train_loader, device, optimizer, criterion):
    """One epoch; returns the mean batch loss (model_probing.py:41-54)."""
    model.train()
    running = 0.0
    for x_dist, x_vec, y in train_loader:
        x_dist, x_vec, y = x_dist.to(device), x_vec.to(device), y.to(device)
        optimizer.zero_grad()
        loss = criterion(model(x_dist, x_vec), y)
        loss.backward()
        optimizer.step()
        running += loss.item()
    return running / len(train_loader)


@torch.no_grad()
def model_evaluate(model, test_loader, criterion, device):
    """-> (all_targets, all_predicts, mean loss, all_outputs), CPU tensors (model_probing.py:86-132)."""
    model.eval()
    outs, preds, tgts, running = [], [], [], 0.0
    for x_dist, x_vec, y in test_loader:
        o = model(x_dist.to(device, non_blocking=True), x_vec.to(device, non_blocking=True))
        running += criterion(o, y.to(device, non_blocking=True)).item()
        o = o.detach().cpu()
        outs.append(o)
        preds.append(o > SIGMA)
        tgts.append(y.detach().cpu())
    return torch.cat(tgts), torch.cat(preds), running / len(test_loader), torch.cat(outs)


@torch.no_grad()
def model_infer(model, test_loader, device):
    """-> (all_predicts, all_outputs), CPU tensors (model_probing.py:135-156)."""
    model.eval()
    outs = [model(x_dist.to(device), x_vec.to(device)) for x_dist, x_vec in test_loader]
    all_outputs = torch.cat(outs).cpu()
    return all_outputs > SIGMA, all_outputs


def to_device_model(model, centroids, scaler_mean, scaler_scale, device=0):
    """torch MLP_2_Input (+ centroids and scaler) -> engine.LiraModel resident on `device`."""
    from .engine import LiraModel
    return LiraModel.from_torch(model, centroids, scaler_mean, scaler_scale, device)
