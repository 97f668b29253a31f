"""Which .contiguous() / .to() calls of the R1 iteration really copy a large tensor, and from where."""
import os, sys, traceback, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, dusty_gan_v2_b200 as pkg
from dusty_gan_v2_b200.gans.trainer import Trainer
from dusty_gan_v2_b200.presets import preset

dev = torch.device("cuda", 0)
pkg.set_precision("bf16")
cfg = preset("dusty_v2", batch_size=64)
tr = Trainer(cfg, bench.cycle(bench.synthetic_batches(4, 64, seed=2, device=dev)), device=dev,
             angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"))
for i in range(3):
    tr.step(i)
torch.cuda.synchronize()
log = collections.Counter()
orig_contig = torch.Tensor.contiguous
orig_to = torch.Tensor.to
orig_clone = torch.Tensor.clone

def where():
    fr = [f for f in traceback.extract_stack()[:-2] if "dusty_gan_v2_b200" in f.filename]
    return " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in fr[-3:])

def producer(t):
    """Name of the autograd node that made t (set under create_graph) and of its inputs' nodes."""
    fn = t.grad_fn
    if fn is None:
        return "leaf/no-graph"
    ups = ",".join(type(u[0]).__name__ if u[0] is not None else "None" for u in fn.next_functions[:3])
    return f"{type(fn).__name__}({ups})"


def contiguous(self, *a, **k):
    out = orig_contig(self, *a, **k)
    if self.is_cuda and self.numel() >= (1 << 22) and out.data_ptr() != self.data_ptr():
        log[("contiguous", tuple(self.shape), tuple(self.stride()), str(self.dtype), where(), producer(self))] += 1
    return out

def to(self, *a, **k):
    out = orig_to(self, *a, **k)
    if self.is_cuda and self.numel() >= (1 << 22) and out.data_ptr() != self.data_ptr():
        log[("to", tuple(self.shape), tuple(self.stride()), f"{self.dtype}->{out.dtype}", where())] += 1
    return out

torch.Tensor.contiguous = contiguous
torch.Tensor.to = to
tr.step(16)
torch.cuda.synchronize()
torch.Tensor.contiguous = orig_contig
torch.Tensor.to = orig_to
for k, v in sorted(log.items(), key=lambda kv: -kv[1] * (1 if True else 0))[:40]:
    print(v, k)
