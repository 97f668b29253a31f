"""Event timeline of CTA 0 of the halo-resident convolution (instrumented library)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DUSTY_LIB", os.path.join(ROOT, "build", "libdusty_b200_prof.so"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dusty_gan_v2_b200.functional as DF  # noqa: E402
from dusty_gan_v2_b200 import _cabi as K  # noqa: E402

lib = K.load()
lib.dusty_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
C, O, H, W = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (32, 32, 66, 514)))
x = torch.randn(64, C, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(O, C, 3, 3, device="cuda") / (C * 9) ** 0.5).to(torch.bfloat16)
buf = (ctypes.c_ulonglong * 8192)()
for _ in range(2):
    DF.conv2d_fprop_tc(x, w, (1, 1))
    torch.cuda.synchronize()
    n = lib.dusty_conv_trace(ctypes.addressof(buf), 8192)
ev = sorted(((buf[i] & 0xffffffffff, buf[i] >> 56, (buf[i] >> 40) & 0xffff) for i in range(n) if buf[i]))
names = {1: "TMA issue tile", 2: "MMA operands ready tile", 3: "MMA issue start blk", 4: "MMA issue end blk",
         5: "EPI acc_full blk", 6: "EPI tmem loaded blk", 7: "EPI stored blk"}
t0 = ev[0][0]
lo, hi = int(os.environ.get("EV_LO", 0)), int(os.environ.get("EV_HI", 260))
for t, tag, idx in ev[lo:hi]:
    print(f"{t - t0:9d}  {'  ' * (0 if tag == 1 else 1 if tag <= 4 else 2)}{names.get(tag, tag)} {idx}")
print("events", n)
