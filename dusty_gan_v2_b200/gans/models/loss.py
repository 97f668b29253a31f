"""Drop-in for gans/models/loss.py: GANLoss (reference 22-88).  Thin scalar math on [B,1]
logits -- stays in PyTorch."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _rel(a, b):
    """a - mean(b) over the batch (relativistic average)."""
    return a - b.mean(0, keepdim=True)


class GANLoss(nn.Module):
    METRICS = ("nsgan", "wgan", "lsgan", "hinge", "ragan", "rahinge", "ralsgan")

    def __init__(self, metric: str, smoothing: float = 1.0):
        super().__init__()
        self.register_buffer("label_real", torch.tensor(1.0))
        self.register_buffer("label_fake", torch.tensor(0.0))
        self.metric = metric
        self.smoothing = smoothing

    def forward(self, pred_real, pred_fake, mode):
        if mode == "G":
            return self.loss_G(pred_real, pred_fake)
        if mode == "D":
            return self.loss_D(pred_real, pred_fake)
        raise ValueError(mode)

    def loss_D(self, r, f):
        m = self.metric
        if m == "nsgan":
            return F.softplus(-r).mean() + F.softplus(f).mean()
        if m == "wgan":
            return -r.mean() + f.mean()
        if m == "lsgan":
            return (F.mse_loss(r, self.label_real.expand_as(r) * self.smoothing)
                    + F.mse_loss(f, self.label_fake.expand_as(f)))
        if m == "hinge":
            return F.relu(1 - r).mean() + F.relu(1 + f).mean()
        if m == "ragan":
            return F.softplus(-_rel(r, f)).mean() + F.softplus(_rel(f, r)).mean()
        if m == "rahinge":
            return F.relu(1 - _rel(r, f)).mean() + F.relu(1 + _rel(f, r)).mean()
        if m == "ralsgan":
            return ((_rel(r, f) - 1.0) ** 2).mean() + ((_rel(f, r) + 1.0) ** 2).mean()
        raise NotImplementedError(m)

    def loss_G(self, r, f):
        m = self.metric
        if m == "nsgan":
            return F.softplus(-f).mean()
        if m in ("wgan", "hinge"):
            return -f.mean()
        if m == "lsgan":
            return F.mse_loss(f, self.label_real.expand_as(f))
        if m == "ragan":
            return F.softplus(_rel(r, f)).mean() + F.softplus(-_rel(f, r)).mean()
        if m == "rahinge":
            return F.relu(1 + _rel(r, f)).mean() + F.relu(1 - _rel(f, r)).mean()
        if m == "ralsgan":
            return ((_rel(r, f) + 1.0) ** 2).mean() + ((_rel(f, r) - 1.0) ** 2).mean()
        raise NotImplementedError(m)
