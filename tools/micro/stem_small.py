import sys, torch
sys.path.insert(0, "/root/repo")
from drone_yolo_b200 import kernels as K
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "u8"
x = torch.rand(1, 3, 64, 64, device=dev)
if which == "u8":
    x = (x * 255).to(torch.uint8)
w = torch.randn(32, 27, device=dev) * 0.3
b = torch.randn(32, device=dev)
out = K.stem_conv(x, w, b)
torch.cuda.synchronize()
ref = torch.nn.functional.silu(torch.nn.functional.conv2d(x.float() / (255 if which == "u8" else 1), w.view(32, 3, 3, 3), b, stride=2, padding=1))
print("ok", which, (out.float() - ref).abs().max().item())
