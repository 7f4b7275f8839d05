#!/usr/bin/env python
"""validate.py — eval entry point (reference flags --cfgs --weight --ema --eval_topk): loads a checkpoint written by the
trainer (keys 'state_dict' / 'ema', models/faceX/face_model.py:73-86) into the B200 backbone and runs the CBIR eval path
(extract -> index -> search) on the data root of the config.  New code over visiondk_b200.cbir."""
import argparse
import os

import torch

from engine.vision_engine import SyntheticFaceData, yaml_load
from visiondk_b200.backbone import BackboneFactory
from visiondk_b200.cbir import FeatureExtractor, index, search


def parse_opt():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfgs", default="configs/faceX/cbir_convnext_b200.yaml")
    ap.add_argument("--weight", default=None)
    ap.add_argument("--ema", action="store_true")
    ap.add_argument("--eval_topk", type=int, default=5)
    return ap.parse_args()


def main(opt):
    cfgs = yaml_load(opt.cfgs)
    if cfgs["model"]["task"] not in ("face", "cbir"):
        raise ValueError("only the faceX / cbir embedding tasks are built for B200")
    device = torch.device("cuda", int(os.getenv("LOCAL_RANK", 0)))
    model = BackboneFactory(cfgs["model"]["backbone"]).get_backbone()
    if opt.weight:
        ckpt = torch.load(opt.weight, map_location="cpu", weights_only=False)
        model.load_state_dict(ckpt["ema"] if opt.ema else ckpt["state_dict"], strict=True)
    data = SyntheticFaceData(str(cfgs["data"]["root"]), cfgs["model"]["image_size"], cfgs["data"]["val"]["bs"], device)
    ext = FeatureExtractor(model)
    k = cfgs["data"]["val"]["metrics"]["cutoffs"][-1]
    idx = index(ext, data.images(data.ids * data.per_id, 11), device)
    scores, ids = search(ext, data.images(data.queries, 12), idx, device, k=k)
    print({"gallery": idx.ntotal, "queries": scores.shape[0], "k": k, "top1_score_mean": float(scores[:, 0].mean())})


if __name__ == "__main__":
    main(parse_opt())
