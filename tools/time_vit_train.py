"""Quick device timing of the ViT-B/16 + CircleLoss train step (BASELINE config 3; not the bench contract).  argv: batch iters."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visiondk_b200.train import FaceTrainingModel, FaceTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = {"backbone": {"timm-vit_base_patch16_224": {"pretrained": False, "image_size": 224, "feat_dim": 512}},
       "head": {"circleloss": {"feat_dim": 512, "num_class": 1000, "margin": 0.25, "gamma": 256}}}
torch.manual_seed(0)
model = FaceTrainingModel(cfg).cuda()
trainer = FaceTrainer(model, lr0=0.01, momentum=0.937, weight_decay=5e-4, label_smooth=0.1, layer_wise=True, warm_steps=0,
                      total_steps=100000, use_ema=True)
x = [torch.randn(B, 3, 224, 224, device="cuda") for _ in range(2)]
y = [torch.randint(0, 1000, (B,), device="cuda") for _ in range(2)]
for i in range(3):
    loss = trainer.step(x[i & 1], y[i & 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    loss = trainer.step(x[i & 1], y[i & 1])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
gflop = 3 * 35.28
print(json.dumps({"model": "vit_base_patch16_224 + CircleLoss(C=1000)", "batch": B, "ms_per_step": ms, "img_per_s": B / ms * 1e3,
                  "tflops": B * gflop / ms, "loss": float(loss)}))
