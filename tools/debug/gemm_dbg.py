import torch, sys
sys.path.insert(0, "/root/repo")
import dusty_gan_v2_b200.functional as DF
torch.manual_seed(0)
def rel(a, b): return float((a.double().cpu() - b).norm() / b.norm())
for (M, N, K) in [(128, 256, 64), (128, 256, 32), (64, 512, 512), (128, 64, 256)]:
    x = torch.randn(M, K); w = torch.randn(N, K)
    ref = x.double() @ w.double().t()
    xd, wd = x.cuda(), w.cuda()
    wt = wd.t().contiguous()          # [K, N] memory -> wt.t() is an MN-major [N, K] operand
    xt = xd.t().contiguous()
    print(M, N, K, "kk", rel(DF.matmul_nt(xd, wd), ref), " b_mn", rel(DF.matmul_nt(xd, wt.t()), ref),
          " a_mn", rel(DF.matmul_nt(xt.t(), wd), ref), " both", rel(DF.matmul_nt(xt.t(), wt.t()), ref))
