"""GPU smoke of the kept entry surface: engine.vision_engine.CenterProcessor.run_embedding (train -> EMA eval -> checkpoint
with the reference's Epoch_N.pt key set) and validate.py-style reload through the backbone factory."""
import copy

import pytest
import torch
import yaml

from engine.vision_engine import CenterProcessor, increment_path, yaml_load
from visiondk_b200.backbone import BackboneFactory

pytestmark = pytest.mark.gpu

CFG = """
model:
  task: cbir
  image_size: &imgsz 64
  load_from: null
  backbone:
    timm-convnext_pico:
      pretrained: False
      image_size: *imgsz
      feat_dim: &featd 64
  head:
    circleloss:
      feat_dim: *featd
      num_class: 8
      margin: 0.25
      gamma: 64
data:
  root: synthetic://cbir?ids=8&per_id=16&queries=32
  nw: 0
  train: {bs: 16, base_aug: null, class_aug: null, augment: [], aug_epoch: 0}
  val: {bs: 16, metrics: {metrics: [mrr], cutoffs: [1, 5]}, augment: []}
hyp:
  epochs: 2
  lr0: 0.01
  lrf_ratio: null
  momentum: 0.937
  weight_decay: 0.0005
  warmup_momentum: 0.8
  warm_ep: 1
  loss: {ce: True}
  label_smooth: 0.1
  optimizer: [sgd, True]
  scheduler: cosine_with_warm
"""


def test_run_embedding_trains_evaluates_and_checkpoints(lib, tmp_path):
    cfg_file = tmp_path / "cbir.yaml"
    cfg_file.write_text(CFG)
    cfgs = yaml_load(str(cfg_file))
    out = increment_path(tmp_path / "run" / "exp")
    cp = CenterProcessor(cfgs, rank=-1, project=str(out))
    cp.run_embedding(None)
    ckpt = torch.load(out / "Epoch_2.pt", map_location="cpu", weights_only=False)
    assert {"epoch", "batch_id", "fitness", "state_dict", "ema", "updates", "optimizer", "scheduler"} <= set(ckpt)
    assert ckpt["epoch"] == 1 and ckpt["updates"] == 16
    model = BackboneFactory(cfgs["model"]["backbone"]).get_backbone()
    model.load_state_dict(ckpt["ema"], strict=True)  # validate.py --ema path (face_model.py:73-86)
    emb = model.cuda().eval().embed(torch.randn(4, 3, 64, 64, device="cuda"), l2_normalize=True)
    assert torch.isfinite(emb).all() and emb.shape == (4, 64)
    # the EMA differs from the raw weights and both differ from initialisation (training happened)
    assert any((ckpt["ema"][k].float() - ckpt["state_dict"][k].float()).abs().max() > 0 for k in ckpt["ema"] if "weight" in k)
