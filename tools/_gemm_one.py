import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.runtime import Context, random_w4
dev = torch.device("cuda", 0)
ctx = Context(0)
oc, ic, M = 15360, 5120, 2048
w, z, s = random_w4(oc, ic, dev, 5, random_zeros=True)
x = torch.randn((M, ic), device=dev).to(torch.float16)
y = torch.empty((M, oc), dtype=torch.float16, device=dev)
for _ in range(4):
    ctx.w4a16_gemv(x, w, z, s, out=y, gemm=True)
torch.cuda.synchronize()
