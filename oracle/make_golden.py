"""Generates tests/golden/*.npz by executing the REFERENCE's own modules (authoring container only).

    python oracle/make_golden.py            # needs /root/reference; never runs on the GPU box

The reference package cannot be imported as a whole (timm / faiss / torchmetrics are not installed), so the
torch-only files are loaded by path: models/faceX/head/{arcface,circleloss}.py, models/losses/loss.py,
models/ema.py, engine/scheduler.py, engine/optimizer.py.  The committed .npz files are what the CPU tests
hold oracle/ to; the GPU tests then hold the CUDA kernels to oracle/.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def heads():
    arc = load("models/faceX/head/arcface.py", "ref_arcface")
    cir = load("models/faceX/head/circleloss.py", "ref_circleloss")
    loss = load("models/losses/loss.py", "ref_loss")
    for name, B, D, Cn, smooth in (("small", 8, 64, 40, 0.0), ("cfg", 24, 512, 160, 0.1)):
        torch.manual_seed(7)
        feats = torch.randn(B, D) * 3.0
        labels = torch.randint(0, Cn, (B,))
        out = {"feats": feats.numpy(), "labels": labels.numpy(), "label_smooth": np.float32(smooth)}
        for kind, mod, kwargs in (("arcface", arc.ArcFace, dict(margin_arc=0.35, margin_am=0.0, scale=32)),
                                  ("arcface_am", arc.ArcFace, dict(margin_arc=0.5, margin_am=0.2, scale=64)),
                                  ("circleloss", cir.CircleLoss, dict(margin=0.25, gamma=256))):
            torch.manual_seed(11)
            head = mod(D, Cn, **kwargs)
            if kind == "arcface_am":
                # make a few rows hit the theta+m > pi fallback branch (cos <= cos(pi - m))
                # (not exactly cos = -1: the reference's sqrt(1 - cos^2) has a NaN gradient there)
                with torch.no_grad():
                    gen = torch.Generator().manual_seed(5)
                    for r in range(0, B, 3):
                        noise = torch.randn(D, generator=gen) * feats[r].norm() / D ** 0.5
                        head.weight[:, labels[r]] = -feats[r] + 0.3 * noise
            f = feats.clone().requires_grad_(True)
            logits = head(f, labels)
            crit = loss.create_Lossfn("ce")(label_smooth=smooth)
            l = crit(logits, labels)
            l.backward()
            out[f"{kind}_weight"] = head.weight.detach().numpy().copy()
            out[f"{kind}_logits"] = logits.detach().numpy()
            out[f"{kind}_loss"] = l.detach().numpy()
            out[f"{kind}_dfeats"] = f.grad.numpy().copy()
            out[f"{kind}_dweight"] = head.weight.grad.numpy().copy()
        np.savez_compressed(os.path.join(OUT, f"heads_{name}.npz"), **out)


def heads_mv():
    mv = load("models/faceX/head/mv_softmax.py", "ref_mv_softmax")
    loss = load("models/losses/loss.py", "ref_loss")
    B, D, Cn, smooth = 16, 64, 40, 0.1
    torch.manual_seed(17)
    feats = torch.randn(B, D) * 2.0
    labels = torch.randint(0, Cn, (B,))
    out = {"feats": feats.numpy(), "labels": labels.numpy(), "label_smooth": np.float32(smooth)}
    for kind, kwargs in (("mv_arc", dict(is_am=False, margin=0.35, mv_weight=1.12, scale=32)),
                         ("mv_am", dict(is_am=True, margin=0.35, mv_weight=1.12, scale=32))):
        torch.manual_seed(19)
        head = mv.MV_Softmax(D, Cn, **kwargs)
        with torch.no_grad():  # pull some class centres towards their samples so that both gt > 0 / gt > margin branches occur
            for r in range(0, B, 2):
                head.weight[:, labels[r]] = feats[r] / feats[r].norm() + 0.4 * head.weight[:, labels[r]]
        f = feats.clone().requires_grad_(True)
        logits = head(f, labels)
        l = loss.create_Lossfn("ce")(label_smooth=smooth)(logits, labels)
        l.backward()
        out[f"{kind}_weight"] = head.weight.detach().numpy().copy()
        out[f"{kind}_logits"] = logits.detach().numpy()
        out[f"{kind}_loss"] = l.detach().numpy()
        out[f"{kind}_dfeats"] = f.grad.numpy().copy()
        out[f"{kind}_dweight"] = head.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "heads_mv.npz"), **out)


def ema_sgd_sched():
    ema_mod = load("models/ema.py", "ref_ema")
    sched = load("engine/scheduler.py", "ref_sched")
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    ema = ema_mod.ModelEMA(model)
    opt = torch.optim.SGD([{"params": model[0].parameters(), "lr": 0.01},
                           {"params": list(model[1].parameters()) + list(model[2].parameters()), "lr": 0.1}],
                          lr=0.01, momentum=0.8, weight_decay=5e-4)
    sc = sched.create_Scheduler("cosine_with_warm", opt, warm_ep=4, epochs=20, lr0=0.01, lrf_ratio=0.1)
    rec = {"init": np.concatenate([p.detach().numpy().ravel() for p in model.state_dict().values()])}
    xs, ys, lrs, params, emas, gnorms = [], [], [], [], [], []
    for step in range(12):
        x = torch.randn(16, 6)
        y = torch.randint(0, 3, (16,))
        lrs.append([g["lr"] for g in opt.param_groups])  # learning rates in effect during this step
        loss = torch.nn.functional.cross_entropy(model(x) * 30, y)
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)  # train.py:209
        opt.step()
        opt.zero_grad()
        ema.update(model)  # train.py:214-215
        sc.step()          # train.py:230 (per batch)
        xs.append(x.numpy()); ys.append(y.numpy()); gnorms.append(float(gn))
        params.append(np.concatenate([p.detach().numpy().ravel() for p in model.state_dict().values()]))
        emas.append(np.concatenate([p.detach().numpy().ravel().astype(np.float32) for p in ema.ema.state_dict().values()]))
    np.savez_compressed(os.path.join(OUT, "step_sgd_ema.npz"), init=rec["init"], x=np.stack(xs), y=np.stack(ys),
                        lrs=np.array(lrs), params=np.stack(params), emas=np.stack(emas), gnorms=np.array(gnorms))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (authoring container only)")
    os.makedirs(OUT, exist_ok=True)
    heads()
    heads_mv()
    ema_sgd_sched()
    print("golden vectors written to", OUT)
