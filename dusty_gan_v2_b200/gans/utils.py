"""Drop-in for the host helpers of gans/utils.py that the training / inversion path imports
(reference 21-31, 85-139, 212-275): seeding, requires_grad toggles, the [-1,1] <-> [0,1] maps,
the infinite windowed-shuffle sampler.  The visualisation helpers of that file (colorize,
save_video, normal maps, spectra) are out of the hot path and raise.
"""
import os
import random

import numpy as np
import torch

from .inversion import SphericalOptimizer, tanh_to_sigmoid  # noqa: F401  (same objects, one home)


def init_random_seed(random_seed=0, rank=0):
    """reference 21-30: every RNG the path draws from gets `random_seed + rank`; cuDNN autotuning on."""
    seed = random_seed + rank
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.enabled = True
    torch.backends.cudnn.benchmark = True


def set_requires_grad(net, requires_grad: bool = True):
    """reference 85-87 (also accepts a plain iterable of parameters)."""
    params = net.parameters() if hasattr(net, "parameters") else net
    for p in params:
        p.requires_grad = requires_grad


def zero_grad(optim):
    """reference 90-93: drop the gradients (set to None) of everything the optimiser owns."""
    for group in optim.param_groups:
        for p in group["params"]:
            p.grad = None


def sigmoid_to_tanh(x):
    """[0,1] -> [-1,+1] (reference 96-99)."""
    return x * 2.0 - 1.0


def noise(tensor, std: float = 0.1):
    """reference 119-121."""
    return tensor + tensor.clone().normal_(0, std)


def cycle(iterable):
    """reference 136-138."""
    while True:
        yield from iterable


_DISTANCES = {"l1": torch.nn.functional.l1_loss, "l2": torch.nn.functional.mse_loss}


def masked_loss(img_ref, img_gen, mask, distance="l1"):
    """reference 225-235 (the older helper kept in utils.py; gans/inversion.py has the one in use):
    per-sample mean of the masked point-wise distance."""
    if distance not in _DISTANCES:
        raise NotImplementedError(distance)
    per_pixel = _DISTANCES[distance](img_ref, img_gen, reduction="none") * mask
    return per_pixel.sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3))


class InfiniteSampler(torch.utils.data.Sampler):
    """reference 238-275: endless index stream for one rank.  A seeded permutation that keeps
    being locally re-shuffled (each visited slot is swapped with a random slot at most `window`
    positions back), every `num_replicas`-th draw going to this rank.  Same RandomState call
    sequence as the reference, hence the same order for the same (seed, rank, replicas)."""

    def __init__(self, dataset, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
        if len(dataset) <= 0 or num_replicas <= 0 or not (0 <= rank < num_replicas):
            raise AssertionError("bad sampler geometry")
        if not (0 <= window_size <= 1):
            raise AssertionError("window_size must be in [0, 1]")
        super().__init__()
        self.dataset, self.rank, self.num_replicas = dataset, rank, num_replicas
        self.shuffle, self.seed, self.window_size = shuffle, seed, window_size

    def __iter__(self):
        n = len(self.dataset)
        order = np.arange(n)
        rnd, window = None, 0
        if self.shuffle:
            rnd = np.random.RandomState(self.seed)
            rnd.shuffle(order)
            window = int(np.rint(n * self.window_size))
        step = 0
        while True:
            slot = step % n
            if step % self.num_replicas == self.rank:
                yield order[slot]
            if window >= 2:
                other = (slot - rnd.randint(window)) % n
                order[slot], order[other] = order[other], order[slot]
            step += 1


def _visualisation_only(name):
    def fn(*_a, **_k):
        raise NotImplementedError(f"gans.utils.{name} is a visualisation helper, outside the B200 hot path")
    fn.__name__ = name
    return fn


colorize = _visualisation_only("colorize")
save_video = _visualisation_only("save_video")
points_to_normal_2d = _visualisation_only("points_to_normal_2d")
power_spectrum_2d = _visualisation_only("power_spectrum_2d")
