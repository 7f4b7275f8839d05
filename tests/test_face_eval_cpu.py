"""engine/faceX/evaluation.py (ours) — host logic of the face task's pair verification: pair-file parsing and the Evaluator,
held to numbers produced by the REFERENCE's own Evaluator.test_one_model (tests/golden/face_verification.npz)."""
import os

import numpy as np
import pytest
import torch

from engine.faceX.evaluation import Evaluator, process_pairtxt

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_pairs():
    z = np.load(os.path.join(GOLD, "face_verification.npz"))
    feats, pairs = z["feats"], z["pairs"]
    names = [f"id{i // 4:03d}/img{i % 4}.jpg" for i in range(len(feats))]  # as oracle/make_golden.py::face_verification named them
    pair_list = [[names[a], names[b], str(l)] for a, b, l in pairs]
    return z, names, feats, pair_list


def test_evaluator_reproduces_the_reference_evaluator():
    z, names, feats, pair_list = golden_pairs()
    name2feat = {os.path.normpath(n): f for n, f in zip(names, feats)}
    mean, std = Evaluator(None).test_one_model(pair_list, name2feat, device="cpu")
    assert abs(mean - float(z["mean"])) < 1e-12 and abs(std - float(z["std"])) < 1e-12
    # un-normalised features are normalised first (is_normalize=False, evaluation.py:60-62)
    scaled = {k: v * np.float32(3.0) for k, v in name2feat.items()}
    mean2, std2 = Evaluator(None).test_one_model(pair_list, scaled, is_normalize=False, device="cpu")
    assert abs(mean2 - mean) < 2e-3
    Evaluator.check_nps(pair_list)
    with pytest.raises(AssertionError, match="multiple of 10"):
        Evaluator.check_nps(pair_list[:-1])


def test_pair_file_parsing_and_end_to_end_from_a_file(tmp_path):
    z, names, feats, pair_list = golden_pairs()
    txt = tmp_path / "pairs.txt"
    txt.write_text("\n".join(" ".join(p) for p in pair_list) + "\n")
    paths, parsed = process_pairtxt(str(txt), "/data/faces")
    assert parsed == pair_list
    used = sorted({n for p in pair_list for n in p[:2]})
    assert paths == [os.path.join("/data/faces", "val", n) for n in used]  # distinct images, sorted (np.unique), under <root>/val

    class Extractor:  # extract_face's contract: {"<parent dir>/<file>": feature} for every image of the loader
        def extract_face(self, dataloader, device):
            row = {n: i for i, n in enumerate(names)}
            out = {}
            for _, _, file_paths in dataloader:
                for p in file_paths:
                    key = os.path.join(os.path.basename(os.path.dirname(p)), os.path.basename(p))
                    out[key] = feats[row[key]]
            return out

    loader = [(None, None, paths[a:a + 64]) for a in range(0, len(paths), 64)]
    mean, std = Evaluator(Extractor()).test(parsed, loader, "cpu")
    assert abs(mean - float(z["mean"])) < 1e-12 and abs(std - float(z["std"])) < 1e-12
    with pytest.raises(AssertionError, match="please check the path"):
        process_pairtxt(str(tmp_path / "missing.txt"), "/data/faces")
