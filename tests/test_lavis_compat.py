"""CPU: the LAVIS checkpoint bridge (pnp_ovss_b200/lavis_compat.py) against a LAVIS-shaped test double
(tests/lavis_shaped_model.py): key map, position-embedding resize (BASE:44-73), ITM logits after loading."""
import pytest
import torch
import torch.nn.functional as F

import synth
from lavis_shaped_model import LavisShapedBlipITM

DIMS = dict(vit_dim=32, vit_depth=2, vit_heads=2, hidden=24, layers=3, heads=2, inter=48, max_pos=64)


def _pair(img_size_ckpt=32, img_size_model=32):
    from pnp_ovss_b200.blip_itm import BlipITM
    tok = synth.SyntheticWordPieceTokenizer()
    torch.manual_seed(5)
    lavis = LavisShapedBlipITM(tok, img_size=img_size_ckpt, **DIMS).eval()
    native = BlipITM(img_size=img_size_model, tokenizer=tok, vocab=30524, **DIMS).eval()
    return tok, lavis, native


def test_checkpoint_loads_and_itm_logits_match():
    tok, lavis, native = _pair()
    ignored = native.load_lavis_checkpoint({"model": lavis.state_dict()})
    assert sorted(ignored) == ["text_encoder.embeddings.position_ids", "text_proj.bias", "text_proj.weight",
                               "vision_proj.bias", "vision_proj.weight"]
    caps = ["A picture of cat aeroplane", "A picture of dog"]
    imgs = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = lavis({"image": imgs, "text_input": caps}, match_head="itm")
        got = native(imgs, caps)
        got_dict_form = native({"image": imgs, "text_input": caps}, match_head="itm")
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got, want)
    assert torch.equal(got, got_dict_form)


def test_export_is_the_inverse_of_load():
    from pnp_ovss_b200 import lavis_compat as L
    tok, lavis, native = _pair()
    sd = {k: v for k, v in lavis.state_dict().items() if not k.startswith(("vision_proj", "text_proj")) and "position_ids" not in k}
    L.load_lavis_state_dict(native, sd)
    back = L.export_lavis_state_dict(native)
    assert sorted(back) == sorted(sd)
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    # every native parameter is covered by the key map
    assert sorted(L.native_to_lavis_keys(native)) == sorted(native.state_dict())


def test_position_embedding_is_resized_like_the_reference_loader():
    from pnp_ovss_b200 import lavis_compat as L
    tok, lavis, native = _pair(img_size_ckpt=64, img_size_model=48)        # 4x4 grid in the checkpoint, 3x3 in the model
    sd = lavis.state_dict()
    native.load_lavis_checkpoint(sd)
    pe = sd["visual_encoder.pos_embed"]
    grid = pe[:, 1:].reshape(1, 4, 4, -1).permute(0, 3, 1, 2)
    want = F.interpolate(grid, size=(3, 3), mode="bicubic", align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
    got = native.visual_encoder.pos_embed.detach()
    assert got.shape == (1, 10, 32)
    assert torch.equal(got[:, :1], pe[:, :1]) and torch.allclose(got[:, 1:], want)
    assert L.resize_pos_embed(pe, 17) is pe
    with pytest.raises(ValueError):
        L.resize_pos_embed(pe, 12)


def test_loader_rejects_missing_and_misshapen_keys_and_warns_on_unknown_ones():
    from pnp_ovss_b200 import lavis_compat as L
    tok, lavis, native = _pair()
    sd = dict(lavis.state_dict())
    missing = dict(sd)
    del missing["itm_head.weight"]
    with pytest.raises(KeyError):
        L.load_lavis_state_dict(native, missing)
    with pytest.warns(UserWarning, match="does not use"):      # strict=False like BASE:112
        assert "text_decoder.cls.bias" in L.load_lavis_state_dict(native, dict(sd, **{"text_decoder.cls.bias": torch.zeros(1)}))
    with pytest.raises(ValueError):
        L.load_lavis_state_dict(native, dict(sd, **{"itm_head.weight": torch.zeros(3, 24)}))


def test_reference_attribute_path_on_the_native_model():
    tok, lavis, native = _pair()
    for m in (lavis, native):
        layers = m.text_encoder.base_model.base_model.encoder.layer
        assert len(layers) == 3 and layers[1].crossattention.self.save_attention is False
    from pnp_ovss_b200.lavis_compat import cross_attention_modules
    assert cross_attention_modules(native)[2] is native.layer[2].crossattention.self
    assert cross_attention_modules(lavis)[2] is lavis.text_encoder.encoder.layer[2].crossattention.self


def test_original_blip_retrieval_checkpoint_extras_are_tolerated():
    """model_large_retrieval_flickr.pth (YAML:10) is the original BLIP retrieval checkpoint: it carries momentum copies, the
    queues and their pointer under the name `ptr_queue`; the reference loads it with strict=False (BASE:112).  A key nobody
    has heard of is ignored with a warning; a missing ITM-path key or a wrong shape still raises."""
    import warnings
    tok, lavis, native = _pair()
    sd = dict(lavis.state_dict())
    sd.update({"ptr_queue": torch.zeros(1, dtype=torch.long), "idx_queue": torch.full((1, 8), -100), "image_queue": torch.randn(4, 8),
               "text_queue": torch.randn(4, 8), "temp": torch.tensor(0.07),
               "visual_encoder_m.cls_token": torch.zeros(1, 1, 32), "text_encoder_m.embeddings.word_embeddings.weight": torch.zeros(4, 24),
               "vision_proj_m.weight": torch.zeros(4, 32), "text_proj_m.weight": torch.zeros(4, 24)})
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ignored = native.load_lavis_checkpoint({"model": sd})
    assert "ptr_queue" in ignored and "visual_encoder_m.cls_token" in ignored
    sd["some.future.buffer"] = torch.zeros(3)
    with pytest.warns(UserWarning, match="does not use"):
        ignored = native.load_lavis_checkpoint(sd)
    assert "some.future.buffer" in ignored
    broken = {k: v for k, v in sd.items() if k != "itm_head.weight"}
    with pytest.raises(KeyError):
        native.load_lavis_checkpoint(broken)
    sd["itm_head.weight"] = torch.zeros(3, 24)
    with pytest.raises(ValueError):
        native.load_lavis_checkpoint(sd)
