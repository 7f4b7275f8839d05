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


def test_pre_norm_clip_tower_matches_hf_clip_vision_model():
    """timm's `vit_*_clip_*` entries (pre_norm=True, bias-free patch embedding, LayerNorm eps 1e-5: BASELINE config 5's ViT-L/14 at
    336^2) restated in oracle/vit.py against HF transformers' CLIPVisionModel with hidden_act='gelu' — the same tower, built
    independently.  HF applies post_layernorm to the pooled token only, so tokens are compared before the final norm, and the cls
    token after it."""
    transformers = pytest.importorskip("transformers")
    dim, depth, heads, patch, size = 128, 3, 2, 14, 56  # patch 14: the 3*14*14 = 588-wide patch rows of ViT-L/14
    o = randomize_(ViTWrapperOracle("x", 32, size, patch=patch, dim=dim, depth=depth, heads=heads, pre_norm=True), seed=2).eval()
    m = o.model
    assert m.patch_embed.proj.bias is None and isinstance(m.norm_pre, torch.nn.LayerNorm) and m.norm.eps == 1e-5
    keys = set(o.state_dict().keys())
    assert {"model.norm_pre.weight", "model.norm_pre.bias"} <= keys and "model.patch_embed.proj.bias" not in keys
    cfg = transformers.CLIPVisionConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * dim,
                                        image_size=size, patch_size=patch, num_channels=3, layer_norm_eps=1e-5, hidden_act="gelu",
                                        attention_dropout=0.0, projection_dim=16)
    hf = transformers.CLIPVisionModel(cfg).eval()
    sd = {"vision_model.embeddings.class_embedding": m.cls_token.data.reshape(-1),
          "vision_model.embeddings.position_embedding.weight": m.pos_embed.data[0],
          "vision_model.embeddings.patch_embedding.weight": m.patch_embed.proj.weight.data,
          "vision_model.pre_layrnorm.weight": m.norm_pre.weight.data, "vision_model.pre_layrnorm.bias": m.norm_pre.bias.data,
          "vision_model.post_layernorm.weight": m.norm.weight.data, "vision_model.post_layernorm.bias": m.norm.bias.data}
    for i, b in enumerate(m.blocks):
        pre = f"vision_model.encoder.layers.{i}."
        qw, kw, vw = b.attn.qkv.weight.data.chunk(3, dim=0)
        qb, kb, vb = b.attn.qkv.bias.data.chunk(3, dim=0)
        for name, w, bb in (("q_proj", qw, qb), ("k_proj", kw, kb), ("v_proj", vw, vb)):
            sd[pre + f"self_attn.{name}.weight"], sd[pre + f"self_attn.{name}.bias"] = w, bb
        sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"] = b.attn.proj.weight.data, b.attn.proj.bias.data
        sd[pre + "layer_norm1.weight"], sd[pre + "layer_norm1.bias"] = b.norm1.weight.data, b.norm1.bias.data
        sd[pre + "layer_norm2.weight"], sd[pre + "layer_norm2.bias"] = b.norm2.weight.data, b.norm2.bias.data
        sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"] = b.mlp.fc1.weight.data, b.mlp.fc1.bias.data
        sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"] = b.mlp.fc2.weight.data, b.mlp.fc2.bias.data
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if "position_ids" not in k], (missing, unexpected)
    torch.manual_seed(1)
    x = torch.randn(2, 3, size, size)
    with torch.no_grad():
        out = hf(pixel_values=x)
        tokens = m.forward_tokens(x)
        final = m(x)
    ref = out.last_hidden_state
    assert tokens.shape == ref.shape == (2, (size // patch) ** 2 + 1, dim)
    assert (tokens - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-5
    assert (final[:, 0] - out.pooler_output).abs().max().item() <= 2e-5 * out.pooler_output.abs().max().item() + 1e-5
