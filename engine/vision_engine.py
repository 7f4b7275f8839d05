"""Orchestrator of the faceX / CBIR embedding path on B200 — the surface of the reference's
engine/vision_engine.py (yaml_load :35-38, increment_path :41-57, CenterProcessor.__init__ :67-167,
run_embedding :438-560) over the visiondk_b200 kernels.  Written from scratch: model / step / eval come from
visiondk_b200.{train,cbir}; datasets are out of the hot-path scope, so `data.root` must be a synthetic:// URL (device-
resident random images with labels) or the caller passes its own iterables of (images[B,3,S,S], labels[B]).
"""
from __future__ import annotations

import glob
import os
import re
import time
from pathlib import Path
from typing import Iterable, Optional
from urllib.parse import parse_qs, urlparse

import torch
import yaml

from visiondk_b200.cbir import FeatureExtractor, index, search
from visiondk_b200.train import FaceTrainer, FaceTrainingModel


def yaml_load(file="data.yaml", error_ok=False):
    with open(file, errors="ignore") as f:
        return yaml.safe_load(f)


def increment_path(path, exist_ok=False, sep="", mkdir=False):
    """run/exp -> run/exp2, run/exp3 ... (first free), like the reference's run-directory naming."""
    path = Path(path)
    if path.exists() and not exist_ok:
        base, suffix = (path.with_suffix(""), path.suffix) if path.is_file() else (path, "")
        taken = [int(m.group(1)) for d in glob.glob(f"{base}{sep}*")
                 if (m := re.search(rf"{re.escape(base.name)}{re.escape(sep)}(\d+)", d))]
        path = Path(f"{base}{sep}{max(taken) + 1 if taken else 2}{suffix}")
    if mkdir:
        path.mkdir(parents=True, exist_ok=True)
    return path


class SyntheticFaceData:
    """`synthetic://cbir?ids=1000&per_id=100&queries=1000`: seeded random images generated on the device."""

    def __init__(self, url: str, image_size: int, batch: int, device, rank: int = 0, world: int = 1):
        q = parse_qs(urlparse(url).query)
        self.ids, self.per_id = int(q.get("ids", ["1000"])[0]), int(q.get("per_id", ["100"])[0])
        self.queries = int(q.get("queries", ["1000"])[0])
        self.size, self.batch, self.device, self.rank, self.world = image_size, batch, device, rank, world

    def __len__(self):  # batches per epoch and rank (DistributedSampler semantics, drop_last)
        return (self.ids * self.per_id) // (self.batch * self.world)

    def train_batches(self, epoch: int) -> Iterable:
        gen = torch.Generator(device=self.device).manual_seed(1000 * epoch + self.rank)  # sampler.set_epoch equivalent
        for _ in range(len(self)):
            yield (torch.randn(self.batch, 3, self.size, self.size, device=self.device, generator=gen),
                   torch.randint(0, self.ids, (self.batch,), device=self.device, generator=gen))

    def images(self, n: int, seed: int) -> Iterable:
        gen = torch.Generator(device=self.device).manual_seed(seed)
        for a in range(0, n, self.batch):
            yield torch.randn(min(self.batch, n - a), 3, self.size, self.size, device=self.device, generator=gen)


class CenterProcessor:
    def __init__(self, cfgs: dict, rank: int, project: Optional[str] = None, train: bool = True, opt=None):
        self.cfgs, self.rank, self.project, self.opt = cfgs, rank, project, opt
        self.model_cfg, self.data_cfg, self.hyp_cfg = cfgs["model"], cfgs["data"], cfgs["hyp"]
        self.task = self.model_cfg["task"]
        if self.task not in ("face", "cbir"):
            raise ValueError("only the faceX / cbir embedding tasks are built for B200")
        if not torch.cuda.is_available():
            raise RuntimeError("visiondk_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device("cuda", max(rank, 0))
        torch.cuda.set_device(self.device)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if rank != -1 else 1
        self.model = FaceTrainingModel(self.model_cfg).to(self.device)
        root = str(self.data_cfg["root"])
        if not root.startswith("synthetic://"):
            raise NotImplementedError("dataset loading is outside the B200 hot-path scope: use a synthetic:// root or drive "
                                      "FaceTrainer / visiondk_b200.cbir with your own (images, labels) iterables")
        self.data = SyntheticFaceData(root, self.model_cfg["image_size"], self.data_cfg["train"]["bs"], self.device,
                                      max(rank, 0), self.world)

    def log(self, msg: str):
        if self.rank in (-1, 0):
            print(msg, flush=True)

    def run_embedding(self, resume: Optional[str] = None):
        hyp = self.hyp_cfg
        steps_per_epoch = len(self.data)
        name, layer_wise = hyp["optimizer"][0], bool(hyp["optimizer"][1])
        if name != "sgd" or hyp["scheduler"] != "cosine_with_warm":
            raise NotImplementedError("the B200 step implements sgd + cosine_with_warm (the faceX / cbir configs)")
        trainer = FaceTrainer(self.model, lr0=hyp["lr0"], momentum=hyp["warmup_momentum"], weight_decay=hyp["weight_decay"],
                              label_smooth=hyp.get("label_smooth", 0.0), layer_wise=layer_wise,
                              warm_steps=hyp["warm_ep"] * steps_per_epoch, total_steps=hyp["epochs"] * steps_per_epoch,
                              lrf_ratio=hyp.get("lrf_ratio"), use_ema=self.rank in (-1, 0))
        start_epoch = 0
        if resume:
            ckpt = torch.load(resume, map_location="cpu", weights_only=False)
            self.model.trainingwrapper["backbone"].load_state_dict(ckpt["state_dict"], strict=True)
            start_epoch = ckpt["epoch"] + 1
        t0 = time.time()
        save_freq = getattr(self.opt, "save_freq", 1) if self.opt else 1
        print_freq = getattr(self.opt, "print_freq", 50) if self.opt else 50
        for epoch in range(start_epoch, hyp["epochs"]):
            if epoch == hyp["warm_ep"]:
                trainer.set_momentum(hyp["momentum"])
            running, n = torch.zeros((), device=self.device), 0
            for it, (images, labels) in enumerate(self.data.train_batches(epoch)):
                running += trainer.step(images, labels)
                n += 1
                if (it + 1) % print_freq == 0:  # one device->host read per print, not per step
                    self.log(f"Epoch {epoch + 1}, iter {it + 1}/{steps_per_epoch}, lr {trainer.opt.param_groups[0]['lr']:.6f}, "
                             f"loss {running.item() / n:.4f}")
                    running.zero_()
                    n = 0
            if self.rank in (-1, 0) and (epoch + 1) % save_freq == 0:
                self.save_and_eval(trainer, epoch, steps_per_epoch)
        self.log(f"Training complete ({(time.time() - t0) / 3600:.3f} hours)")

    def save_and_eval(self, trainer: FaceTrainer, epoch: int, steps_per_epoch: int):
        ema_backbone = trainer.ema.trainingwrapper["backbone"] if trainer.ema is not None else self.model.trainingwrapper["backbone"]
        ext = FeatureExtractor(ema_backbone)
        k = self.data_cfg["val"]["metrics"]["cutoffs"][-1]
        idx = index(ext, self.data.images(min(self.data.ids * self.data.per_id, 4096), 11), self.device)
        scores, ids = search(ext, self.data.images(min(self.data.queries, 256), 12), idx, self.device, k=k)
        fitness = {"fitness": {"top1_score_mean": float(scores[:, 0].mean())}, "checkpoint": f"Epoch_{epoch + 1}.pt"}
        out_dir = Path(self.project or "run/exp")
        out_dir.mkdir(parents=True, exist_ok=True)
        ckpt = {"epoch": epoch, "batch_id": steps_per_epoch - 1, "fitness": fitness,
                "state_dict": self.model.trainingwrapper["backbone"].state_dict(), "ema": ema_backbone.state_dict(),
                "updates": trainer.opt.updates, "optimizer": {"steps": trainer.opt.steps, "param_groups": trainer.opt.param_groups},
                "scheduler": {"step": trainer.sched_step}}
        torch.save(ckpt, out_dir / f"Epoch_{epoch + 1}.pt")
        self.log(str(fitness))
