"""Quick device timing of the ConvNeXt-B embedding forward (not the bench contract).  argv[1] = batch sizes csv."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visiondk_b200.backbone import TimmWrapper

sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64,128,256").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = TimmWrapper("convnext_base", 512, 224, pretrained=False).cuda().eval()
for B in sizes:
    x = torch.randn(B, 3, 224, 224, device="cuda")
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
    print(json.dumps({"batch": B, "ms": ms, "img_per_s": B / ms * 1e3, "tflops": B * 30.76e9 / ms / 1e9}))
