"""Quick device timing of the ViT embedding forward (not the bench contract).  argv: model batch iters."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visiondk_b200.vit import ViTWrapper, VIT_ARCHS

name = sys.argv[1] if len(sys.argv) > 1 else "vit_base_patch16_224"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
size = 336 if name.endswith("336") else 224
m = ViTWrapper(name, 512, size, pretrained=False).cuda().eval()
x = torch.randn(B, 3, size, size, device="cuda")
for _ in range(2):
    y = m.embed(x, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = m.embed(x, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
patch, dim, depth, heads = VIT_ARCHS[name]
T = (size // patch) ** 2 + 1
flops = depth * (2.0 * T * dim * dim * 12 + 4.0 * T * T * dim) + 2.0 * (T - 1) * 3 * patch * patch * dim + 2.0 * T * dim * 512
print(json.dumps({"model": name, "batch": B, "ms": ms, "img_per_s": B / ms * 1e3, "tflops": B * flops / ms / 1e9, "gflop_per_img": flops / 1e9}))
