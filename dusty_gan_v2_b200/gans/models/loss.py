"""Drop-in for gans/models/loss.py: `GANLoss(metric, smoothing)(pred_real, pred_fake, mode)`
(reference 22-88) for the seven objectives the reference implements.  Scalar math on [B, 1]
logits: stays in PyTorch.  Each objective is a pair (discriminator loss, generator loss) of
functions of (real logits, fake logits); the relativistic-average ones see each side relative to
the batch mean of the other (reference 7-18).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _rel(a, b):
    """a - mean(b) over the batch (relativistic average)."""
    return a - b.mean(0, keepdim=True)


def _sq(x):
    return (x ** 2).mean()


def _objectives(crit):
    one = lambda t: crit.label_real.expand_as(t)            # noqa: E731
    zero = lambda t: crit.label_fake.expand_as(t)           # noqa: E731
    sp, relu = F.softplus, F.relu
    return {
        "nsgan": (lambda r, f: sp(-r).mean() + sp(f).mean(),
                  lambda r, f: sp(-f).mean()),
        "wgan": (lambda r, f: -r.mean() + f.mean(),
                 lambda r, f: -f.mean()),
        "lsgan": (lambda r, f: F.mse_loss(r, one(r) * crit.smoothing) + F.mse_loss(f, zero(f)),
                  lambda r, f: F.mse_loss(f, one(f))),
        "hinge": (lambda r, f: relu(1 - r).mean() + relu(1 + f).mean(),
                  lambda r, f: -f.mean()),
        "ragan": (lambda r, f: sp(-_rel(r, f)).mean() + sp(_rel(f, r)).mean(),
                  lambda r, f: sp(_rel(r, f)).mean() + sp(-_rel(f, r)).mean()),
        "rahinge": (lambda r, f: relu(1 - _rel(r, f)).mean() + relu(1 + _rel(f, r)).mean(),
                    lambda r, f: relu(1 + _rel(r, f)).mean() + relu(1 - _rel(f, r)).mean()),
        "ralsgan": (lambda r, f: _sq(_rel(r, f) - 1.0) + _sq(_rel(f, r) + 1.0),
                    lambda r, f: _sq(_rel(r, f) + 1.0) + _sq(_rel(f, r) - 1.0)),
    }


class GANLoss(nn.Module):
    METRICS = ("nsgan", "wgan", "lsgan", "hinge", "ragan", "rahinge", "ralsgan")

    def __init__(self, metric: str, smoothing: float = 1.0):
        super().__init__()
        self.register_buffer("label_real", torch.tensor(1.0))
        self.register_buffer("label_fake", torch.tensor(0.0))
        self.metric = metric
        self.smoothing = smoothing

    def _pair(self):
        table = _objectives(self)
        if self.metric not in table:
            raise NotImplementedError(self.metric)
        return table[self.metric]

    def loss_D(self, pred_real, pred_fake):
        return self._pair()[0](pred_real, pred_fake)

    def loss_G(self, pred_real, pred_fake):
        return self._pair()[1](pred_real, pred_fake)

    def forward(self, pred_real, pred_fake, mode):
        if mode not in ("G", "D"):
            raise ValueError(mode)
        return (self.loss_G if mode == "G" else self.loss_D)(pred_real, pred_fake)
