"""Quick device timing of the ConvNeXt-B embedding forward at a few batch sizes (not the bench contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visiondk_b200.backbone import TimmWrapper

m = TimmWrapper("convnext_base", 512, 224, pretrained=False).cuda().eval()
for B in (64, 128, 256):
    x = torch.randn(B, 3, 224, 224, device="cuda")
    for _ in range(3):
        y = m.embed(x, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        y = m.embed(x, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t0 = time.perf_counter()
    for _ in range(n):
        y = m.embed(x, True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    print(json.dumps({"batch": B, "ms": ms, "wall_ms": wall, "img_per_s": B / ms * 1e3, "tflops": B * 30.76e9 / ms / 1e9}))
