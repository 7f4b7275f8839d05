"""The ViT oracle (oracle/vit.py restates timm 0.9.16's VisionTransformer; timm itself is not installed: parity unpinned)
cross-checked against an independent implementation of the same architecture that IS installed: HF transformers' ViTModel."""
import pytest
import torch

from oracle.vit import ViTWrapperOracle, randomize_


def test_vit_oracle_matches_hf_vit_model():
    transformers = pytest.importorskip("transformers")
    dim, depth, heads, patch, size = 96, 3, 4, 16, 64
    o = randomize_(ViTWrapperOracle("x", 32, size, patch=patch, dim=dim, depth=depth, heads=heads), seed=1).eval()
    cfg = transformers.ViTConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * dim,
                                 image_size=size, patch_size=patch, num_channels=3, qkv_bias=True, layer_norm_eps=1e-6,
                                 hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = transformers.ViTModel(cfg, add_pooling_layer=False).eval()
    sd = {}
    m = o.model
    sd["embeddings.cls_token"] = m.cls_token.data
    sd["embeddings.position_embeddings"] = m.pos_embed.data
    sd["embeddings.patch_embeddings.projection.weight"] = m.patch_embed.proj.weight.data
    sd["embeddings.patch_embeddings.projection.bias"] = m.patch_embed.proj.bias.data
    for i, b in enumerate(m.blocks):
        pre = f"encoder.layer.{i}."
        qw, kw, vw = b.attn.qkv.weight.data.chunk(3, dim=0)
        qb, kb, vb = b.attn.qkv.bias.data.chunk(3, dim=0)
        for name, w, bb in (("query", qw, qb), ("key", kw, kb), ("value", vw, vb)):
            sd[pre + f"attention.attention.{name}.weight"] = w
            sd[pre + f"attention.attention.{name}.bias"] = bb
        sd[pre + "attention.output.dense.weight"] = b.attn.proj.weight.data
        sd[pre + "attention.output.dense.bias"] = b.attn.proj.bias.data
        sd[pre + "layernorm_before.weight"] = b.norm1.weight.data
        sd[pre + "layernorm_before.bias"] = b.norm1.bias.data
        sd[pre + "layernorm_after.weight"] = b.norm2.weight.data
        sd[pre + "layernorm_after.bias"] = b.norm2.bias.data
        sd[pre + "intermediate.dense.weight"] = b.mlp.fc1.weight.data
        sd[pre + "intermediate.dense.bias"] = b.mlp.fc1.bias.data
        sd[pre + "output.dense.weight"] = b.mlp.fc2.weight.data
        sd[pre + "output.dense.bias"] = b.mlp.fc2.bias.data
    sd["layernorm.weight"] = m.norm.weight.data
    sd["layernorm.bias"] = m.norm.bias.data
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if "pooler" not in k], (missing, unexpected)
    torch.manual_seed(0)
    x = torch.randn(2, 3, size, size)
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state
        got = m(x)
    assert got.shape == ref.shape == (2, (size // patch) ** 2 + 1, dim)
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-5


def test_vit_wrapper_neck_shapes_and_keys():
    o = ViTWrapperOracle("vit_tiny_patch16_224", 64, 224)
    keys = set(o.state_dict().keys())
    for k in ("model.cls_token", "model.pos_embed", "model.patch_embed.proj.weight", "model.blocks.0.attn.qkv.weight",
              "model.blocks.11.mlp.fc2.bias", "model.norm.weight", "output_layer.0.weight", "output_layer.2.weight",
              "output_layer.3.running_mean"):
        assert k in keys, k
    assert o.output_layer[2].weight.shape == (64, 197 * 192)
    with torch.no_grad():
        assert o.eval()(torch.randn(2, 3, 224, 224)).shape == (2, 64)
