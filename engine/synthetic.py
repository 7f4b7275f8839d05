"""Synthetic image source behind `data.root: synthetic://cbir?ids=1000&per_id=100&queries=1000[&noise=0.6]`.

The reference reads folders / HuggingFace datasets through dataset/basedataset.py (ImageDatasets :18-306, CBIRDatasets
:463-596) — CPU image IO that is outside the hot-path scope.  BASELINE.json's configs are quoted on
synthetic images of the named shapes, so the entry points need a data root that yields them: seeded, device-resident, with
an identity structure (every identity has a low-resolution prototype; an image is its up-sampled prototype plus noise), so
that training has something to learn and the CBIR metrics of the eval tail are meaningful.

Query / gallery layout follows CBIRDatasets: gallery item i has label `i % ids` (per_id items per identity), one query per
identity (the first `queries` identities); a query's positives are the gallery items of its identity."""
from __future__ import annotations

from typing import Iterable
from urllib.parse import parse_qs, urlparse

import torch
import torch.nn.functional as F


def is_synthetic(root) -> bool:
    return str(root).startswith("synthetic://")


class SyntheticFaceData:
    def __init__(self, url: str, image_size: int, batch: int, device, rank: int = 0, world: int = 1):
        q = parse_qs(urlparse(url).query)

        def arg(name, default, cast=int):
            return cast(q.get(name, [default])[0])

        self.ids, self.per_id = arg("ids", 1000), arg("per_id", 100)
        self.queries = min(arg("queries", 1000), self.ids)
        self.noise = arg("noise", 0.6, float)
        self.size, self.batch, self.device, self.rank, self.world = image_size, batch, torch.device(device), rank, world
        self._proto = None

    # ---- identity prototypes: [ids, 3, 8, 8], the same on every rank ----
    def prototypes(self) -> torch.Tensor:
        if self._proto is None:
            gen = torch.Generator(device=self.device).manual_seed(20240917)
            self._proto = torch.randn(self.ids, 3, 8, 8, device=self.device, generator=gen)
        return self._proto

    def render(self, labels: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
        base = F.interpolate(self.prototypes()[labels], size=(self.size, self.size), mode="nearest")
        return base + self.noise * torch.randn(base.shape, device=self.device, generator=gen)

    # ---- train split ----
    @property
    def num_classes(self) -> int:
        return self.ids

    def __len__(self):  # batches per epoch and rank (DistributedSampler + drop_last, engine/vision_engine.py:458-468)
        return (self.ids * self.per_id) // (self.batch * self.world)

    def train_batches(self, epoch: int) -> Iterable:
        gen = torch.Generator(device=self.device).manual_seed(1000 * epoch + self.rank)  # sampler.set_epoch equivalent
        for _ in range(len(self)):
            labels = torch.randint(0, self.ids, (self.batch,), device=self.device, generator=gen)
            yield self.render(labels, gen), labels

    # ---- query / gallery splits (shuffle=False: order defines the ids the search returns) ----
    def gallery_labels(self, limit: int | None = None) -> torch.Tensor:
        n = self.ids * self.per_id if limit is None else min(limit, self.ids * self.per_id)
        return torch.arange(n, device=self.device) % self.ids

    def query_labels(self, limit: int | None = None) -> torch.Tensor:
        n = self.queries if limit is None else min(limit, self.queries)
        return torch.arange(n, device=self.device)

    def _batches(self, labels: torch.Tensor, seed: int) -> Iterable:
        gen = torch.Generator(device=self.device).manual_seed(seed)
        for a in range(0, labels.numel(), self.batch):
            yield self.render(labels[a:a + self.batch], gen)

    def gallery_batches(self, limit: int | None = None) -> Iterable:
        return self._batches(self.gallery_labels(limit), 11)

    def query_batches(self, limit: int | None = None) -> Iterable:
        return self._batches(self.query_labels(limit), 12)
