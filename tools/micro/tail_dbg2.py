import sys, torch
sys.path.insert(0, "/root/repo")
from drone_yolo_b200 import kernels as K
import torch.nn.functional as F
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
B, H, W = 1, 16, 8
x = torch.randn(B, 64, H, W, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w1 = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(dev); b1 = torch.randn(64, generator=g).to(dev)
w1p, b1p = K.pack_conv_weight(w1, b1)
w2 = torch.eye(64, device=dev)
w2p, b2p = K.pack_conv_weight(w2.view(64, 64, 1, 1), torch.zeros(64, device=dev))
ref = F.silu(F.conv2d(x.float(), w1.to(torch.bfloat16).float(), b1, padding=1))[0].permute(1, 2, 0).reshape(-1, 64)
torch.cuda.synchronize()
outs = []
for it in range(3):
    out2 = K.empty_nhwc(B, 64, H, W, dev, torch.float32)
    out2.fill_(-7.0)
    torch.cuda.synchronize()
    K.conv2d(x, w1p, b1p, 64, 3, 1, True, tail=(w2p, b2p, 64, out2))
    torch.cuda.synchronize()
    o = out2[0].permute(1, 2, 0).reshape(-1, 64).clone()
    outs.append(o)
    good = ((o - ref).abs() < 3e-2)
    print("run", it, "rows fully ok:", int(good.all(1).sum()), "/128; elements ok:", int(good.sum()), "/8192; untouched(-7):", int((o == -7.0).sum()))
    print("   per-16B-chunk ok count by chunk index:", [int(good[:, 4 * j:4 * j + 4].all(1).sum()) for j in range(16)])
print("run0 == run1:", bool(torch.equal(outs[0], outs[1])), " run1 == run2:", bool(torch.equal(outs[1], outs[2])))
