"""The Python host mirrors the reference's model / data API (src/models/custom.py, blocks.py,
src/data/prepare_data.py): same constructor, attributes, module tree, state_dict schema and error
behaviour.  CPU only - nothing here computes."""
import numpy as np
import pytest
import torch

import fsr_b200
from oracle import weights


def test_state_dict_schema_matches_reference():
    m = fsr_b200.FaceEnhanceNet(num_groups=6, blocks_per_group=10)
    schema = weights.state_dict_schema()
    sd = m.state_dict()
    assert list(sd.keys()) == list(schema.keys())
    assert [tuple(v.shape) for v in sd.values()] == list(schema.values())
    assert all(v.dtype == torch.float32 for v in sd.values())


def test_constructor_defaults_and_kwargs_override():
    m = fsr_b200.FaceEnhanceNet()
    assert (m.config.num_groups, m.config.blocks_per_group, m.config.num_channels) == (3, 4, 64)
    assert m.scale_factor == 4 and m.num_channels == 64 and m.config.res_scale == 0.2
    m2 = fsr_b200.FaceEnhanceNet(num_groups=2, blocks_per_group=1, not_a_field=5)  # unknown kwargs ignored
    assert len(m2.residual_groups) == 2 and len(m2.residual_groups[0].blocks) == 1
    cfg = fsr_b200.FaceEnhanceNetConfig(num_groups=1)
    m3 = fsr_b200.FaceEnhanceNet(cfg, blocks_per_group=2)
    assert m3.config is cfg and cfg.blocks_per_group == 2
    m4 = fsr_b200.create_face_enhance_net(num_rcab_blocks=4, num_groups=1)
    assert m4.config.num_rcab_blocks == 4 and len(m4.residual_groups) == 1


def test_literal_init_matches_reference_recipe():
    torch.manual_seed(0)
    m = fsr_b200.FaceEnhanceNet(num_groups=1, blocks_per_group=1)
    assert m.conv_last.weight.abs().max() == 0 and m.conv_last.bias.abs().max() == 0
    assert all(p.abs().max() == 0 for n, p in m.named_parameters() if n.endswith(".bias"))
    assert torch.all(m.residual_groups[0].blocks[0].prelu.weight == 0.25)
    w = m.residual_groups[0].blocks[0].conv1.weight
    assert abs(w.std().item() - (2.0 / (64 * 9)) ** 0.5) < 0.01  # Kaiming normal, fan_out
    assert m.residual_groups[0].blocks[0].channel_attention.fc[0].weight.shape == (16, 64)


def test_load_state_dict_strict_roundtrip():
    cfg = dict(num_groups=2, blocks_per_group=2)
    sd = weights.make_state_dict(1, "T1", **cfg)
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    back = m.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    with pytest.raises(RuntimeError):
        fsr_b200.FaceEnhanceNet(num_groups=1, blocks_per_group=2).load_state_dict(sd, strict=True)


def test_module_tree_names_used_by_reference_callers():
    m = fsr_b200.FaceEnhanceNet(num_groups=1, blocks_per_group=1)
    names = dict(m.named_modules())
    for n in ("residual_groups.0.blocks.0.channel_attention.global_pool",
              "residual_groups.0.blocks.0.channel_attention.fc.2", "upsample.stages.1.pixel_shuffle",
              "conv_after_body", "residual_groups.0.conv"):
        assert n in names
    assert m.residual_groups[0].blocks[0].res_scale == 0.2


def test_no_cpu_fallback_anywhere():
    m = fsr_b200.FaceEnhanceNet(num_groups=1, blocks_per_group=1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="expected input"):
        m(torch.rand(3, 64, 64))
    with pytest.raises(RuntimeError, match="parameter container"):
        m.residual_groups[0].blocks[0](torch.rand(1, 64, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.lr_from_hr(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))


def test_training_path_has_no_cpu_fallback_either():
    # train() mode with grad enabled takes the fen_forward_train / fen_backward path: CUDA only, like the rest
    m = fsr_b200.FaceEnhanceNet(num_groups=1, blocks_per_group=1).train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.Stage1Step(m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.l1_loss(torch.rand(2, 2), torch.rand(2, 2))
    m.mark_parameters_updated()          # harmless without a packed copy
    assert m._packed is None


def test_model_info_keys():
    info = fsr_b200.FaceEnhanceNet(num_groups=6, blocks_per_group=10).get_model_info()
    assert info["total_params"] == 5_115_651 and info["total_rcab_blocks"] == 60
    assert set(info) >= {"name", "trainable_params", "size_mb", "num_groups", "blocks_per_group",
                         "num_channels", "scale_factor", "input_size", "output_size"}


def test_lite_variant_constructs_but_is_not_runnable():
    lite = fsr_b200.FaceEnhanceNetLite()
    assert lite.config.num_channels == 32
    assert lite.conv_first.weight.shape == (32, 3, 3, 3)


def test_create_lr_image_argument_errors():
    img = np.zeros((256, 256, 3), np.uint8)
    with pytest.raises(ValueError, match="Unknown degradation method"):
        fsr_b200.create_lr_image(img, 64, "lanczos")
    with pytest.raises(NotImplementedError):
        fsr_b200.create_lr_image(img, 64, "bilinear")
    with pytest.raises(NotImplementedError):
        fsr_b200.create_lr_image(img, 32, "bicubic")
    with pytest.raises(TypeError):
        fsr_b200.create_lr_image(img.astype(np.float32), 64)
    with pytest.raises(TypeError):
        fsr_b200.lr_from_hr(torch.zeros(1, 8, 8, 3))


def test_to_tensor():
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    t = fsr_b200.to_tensor(img)
    assert t.shape == (3, 2, 3) and t.dtype == torch.float32
    assert torch.equal(t, torch.from_numpy(img.transpose(2, 0, 1)).float() / 255.0)
    assert fsr_b200.to_tensor(img, normalize=False).dtype == torch.uint8
