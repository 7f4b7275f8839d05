"""Cross-checks oracle/convnext.py (timm 0.9.16 restatement, parity unpinned at the timm boundary) against the
architecture-identical torchvision ConvNeXt, and checks the timm state_dict surface."""
import torch
import torchvision

from oracle.convnext import CONVNEXT_ARCHS, ConvNeXt, TimmWrapperOracle, randomize_


def map_torchvision_to_timm(tv_sd, depths):
    """torchvision convnext `features.*` keys -> timm ConvNeXt keys."""
    out = {}
    out["stem.0.weight"], out["stem.0.bias"] = tv_sd["features.0.0.weight"], tv_sd["features.0.0.bias"]
    out["stem.1.weight"], out["stem.1.bias"] = tv_sd["features.0.1.weight"], tv_sd["features.0.1.bias"]
    for i, d in enumerate(depths):
        f = 1 + 2 * i
        if i > 0:
            for a, b in (("0", "0"), ("1", "1")):
                for p in ("weight", "bias"):
                    out[f"stages.{i}.downsample.{b}.{p}"] = tv_sd[f"features.{f - 1}.{a}.{p}"]
        for j in range(d):
            src, dst = f"features.{f}.{j}", f"stages.{i}.blocks.{j}"
            out[f"{dst}.conv_dw.weight"], out[f"{dst}.conv_dw.bias"] = tv_sd[f"{src}.block.0.weight"], tv_sd[f"{src}.block.0.bias"]
            out[f"{dst}.norm.weight"], out[f"{dst}.norm.bias"] = tv_sd[f"{src}.block.2.weight"], tv_sd[f"{src}.block.2.bias"]
            out[f"{dst}.mlp.fc1.weight"], out[f"{dst}.mlp.fc1.bias"] = tv_sd[f"{src}.block.3.weight"], tv_sd[f"{src}.block.3.bias"]
            out[f"{dst}.mlp.fc2.weight"], out[f"{dst}.mlp.fc2.bias"] = tv_sd[f"{src}.block.5.weight"], tv_sd[f"{src}.block.5.bias"]
            out[f"{dst}.gamma"] = tv_sd[f"{src}.layer_scale"].reshape(-1)
    return out


def test_oracle_matches_torchvision_convnext_tiny_features():
    torch.manual_seed(0)
    tv = torchvision.models.convnext_tiny(weights=None).eval()
    with torch.no_grad():  # make layer-scale visible
        for n, p in tv.named_parameters():
            if n.endswith("layer_scale"):
                p.fill_(0.3)
    depths, dims = CONVNEXT_ARCHS["convnext_tiny"]
    ours = ConvNeXt(depths, dims).eval()
    sd = map_torchvision_to_timm(tv.state_dict(), depths)
    sd["head.norm.weight"], sd["head.norm.bias"] = torch.ones(dims[-1]), torch.zeros(dims[-1])
    ours.load_state_dict(sd, strict=True)
    x = torch.randn(2, 3, 96, 96)
    with torch.no_grad():
        ref = tv.features(x)
        got = ours.stages(ours.stem(x))  # torchvision's features stop before timm's head norm
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)


def test_timm_state_dict_surface_and_shapes():
    m = TimmWrapperOracle("convnext_base", 512, 224)
    sd = m.state_dict()
    assert sd["model.stem.0.weight"].shape == (128, 3, 4, 4)
    assert sd["model.stages.2.blocks.26.mlp.fc1.weight"].shape == (2048, 512)
    assert sd["model.stages.3.downsample.1.weight"].shape == (1024, 512, 2, 2)
    assert sd["model.head.norm.weight"].shape == (1024,)
    assert sd["output_layer.2.weight"].shape == (512, 50176)
    assert sd["output_layer.3.running_var"].shape == (512,)
    assert "model.stages.0.downsample.0.weight" not in sd
    n_backbone = sum(p.numel() for n, p in m.named_parameters() if n.startswith("model."))
    assert abs(n_backbone - 87.56e6) < 0.02e6  # 87.56 M


def test_small_wrapper_runs_train_and_eval():
    m = randomize_(TimmWrapperOracle("x", 64, 64, depths=(1, 1, 2, 1), dims=(16, 32, 64, 128)), seed=1)
    x = torch.randn(4, 3, 64, 64)
    m.train()
    y = m(x)
    assert y.shape == (4, 64)
    y.square().mean().backward()
    m.eval()
    with torch.no_grad():
        assert m(x).shape == (4, 64)


def test_oracle_matches_hf_convnext_model():
    """A second independent implementation of the architecture: HF transformers' ConvNextModel (weights mapped across).  Its
    `last_hidden_state` is the feature map BEFORE the final LayerNorm (HF normalises the pooled vector instead), i.e. the
    input of timm's head.norm."""
    import pytest
    transformers = pytest.importorskip("transformers")
    depths, dims = (1, 1, 2, 1), (32, 64, 96, 128)
    ours = randomize_(ConvNeXt(depths, dims), seed=2).eval()
    cfg = transformers.ConvNextConfig(num_channels=3, patch_size=4, num_stages=4, hidden_sizes=list(dims), depths=list(depths),
                                      hidden_act="gelu", layer_norm_eps=1e-6, layer_scale_init_value=1e-6, drop_path_rate=0.0,
                                      image_size=64)
    hf = transformers.ConvNextModel(cfg).eval()
    sd = {}
    o = ours.state_dict()
    sd["embeddings.patch_embeddings.weight"], sd["embeddings.patch_embeddings.bias"] = o["stem.0.weight"], o["stem.0.bias"]
    sd["embeddings.layernorm.weight"], sd["embeddings.layernorm.bias"] = o["stem.1.weight"], o["stem.1.bias"]
    for i, d in enumerate(depths):
        if i > 0:
            sd[f"encoder.stages.{i}.downsampling_layer.0.weight"] = o[f"stages.{i}.downsample.0.weight"]
            sd[f"encoder.stages.{i}.downsampling_layer.0.bias"] = o[f"stages.{i}.downsample.0.bias"]
            sd[f"encoder.stages.{i}.downsampling_layer.1.weight"] = o[f"stages.{i}.downsample.1.weight"]
            sd[f"encoder.stages.{i}.downsampling_layer.1.bias"] = o[f"stages.{i}.downsample.1.bias"]
        for j in range(d):
            src, dst = f"stages.{i}.blocks.{j}", f"encoder.stages.{i}.layers.{j}"
            sd[f"{dst}.dwconv.weight"], sd[f"{dst}.dwconv.bias"] = o[f"{src}.conv_dw.weight"], o[f"{src}.conv_dw.bias"]
            sd[f"{dst}.layernorm.weight"], sd[f"{dst}.layernorm.bias"] = o[f"{src}.norm.weight"], o[f"{src}.norm.bias"]
            sd[f"{dst}.pwconv1.weight"], sd[f"{dst}.pwconv1.bias"] = o[f"{src}.mlp.fc1.weight"], o[f"{src}.mlp.fc1.bias"]
            sd[f"{dst}.pwconv2.weight"], sd[f"{dst}.pwconv2.bias"] = o[f"{src}.mlp.fc2.weight"], o[f"{src}.mlp.fc2.bias"]
            sd[f"{dst}.layer_scale_parameter"] = o[f"{src}.gamma"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("layernorm.") for k in missing), (missing, unexpected)
    torch.manual_seed(0)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state
        got = ours.stages(ours.stem(x))  # before head.norm
    assert got.shape == ref.shape == (2, dims[-1], 2, 2)
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-5
