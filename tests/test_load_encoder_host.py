"""CPU: `load_encoder` (map_detector_core.py:150-179) -- state-dict and full-module checkpoints, hydra-style YAML nodes."""
import functools
import sys
import types

import pytest
import torch
import yaml

from amphibian_vae_latent_detector_b200 import reference_api as api


def _tiny(c=4):
    return torch.nn.Sequential(torch.nn.Conv2d(1, c, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(c * 6 * 4, 8))


def test_instantiate_nested_targets_partial_and_args():
    mod = types.ModuleType("avld_fake_factory")
    mod.make = lambda width, act=None, *rest: ("made", width, act, rest)
    mod.act = lambda kind="relu": ("act", kind)
    sys.modules["avld_fake_factory"] = mod
    try:
        cfg = {"_target_": "avld_fake_factory.make", "width": 7, "act": {"_target_": "avld_fake_factory.act", "kind": "gelu"}}
        assert api._instantiate(cfg) == ("made", 7, ("act", "gelu"), ())
        p = api._instantiate({"_target_": "avld_fake_factory.make", "_partial_": True, "width": 3})
        assert isinstance(p, functools.partial) and p() == ("made", 3, None, ())
        a = api._instantiate({"_target_": "avld_fake_factory.make", "_args_": [1, {"_target_": "avld_fake_factory.act"}, 9]})
        assert a == ("made", 1, ("act", "relu"), (9,))
        assert api._instantiate({"plain": [1, {"_target_": "avld_fake_factory.act"}]}) == {"plain": [1, ("act", "relu")]}
    finally:
        del sys.modules["avld_fake_factory"]


def test_load_encoder_state_dict_with_factory_yaml(tmp_path):
    mod = types.ModuleType("avld_fake_models")
    mod.tiny_factory = lambda c=4: (lambda: _tiny(c))          # the YAML names a factory of factories, as `_BirdNet` does
    sys.modules["avld_fake_models"] = mod
    try:
        net = _tiny(4)
        torch.save({"state_dict": net.state_dict()}, tmp_path / "enc.pt")
        (tmp_path / "enc.yaml").write_text(yaml.safe_dump({"encoder": {"_target_": "avld_fake_models.tiny_factory", "c": 4}}))
        got = api.load_encoder(tmp_path / "enc.pt", tmp_path / "enc.yaml", tmp_path)
        assert not got.training
        for a, b in zip(got.state_dict().values(), net.state_dict().values()):
            assert torch.equal(a, b)
    finally:
        del sys.modules["avld_fake_models"]


def test_full_module_checkpoint_needs_trust(tmp_path, monkeypatch):
    net = _tiny(4)
    torch.save(net, tmp_path / "full.pt")                        # the pickled module itself (core:160-165 accepts it)
    monkeypatch.delenv("AVLD_TRUST_CHECKPOINTS", raising=False)
    with pytest.raises(RuntimeError, match="trust_checkpoint"):
        api.load_encoder(tmp_path / "full.pt", tmp_path / "none.yaml", tmp_path)
    got = api.load_encoder(tmp_path / "full.pt", tmp_path / "none.yaml", tmp_path, trust_checkpoint=True)
    assert isinstance(got, torch.nn.Sequential) and not got.training
    monkeypatch.setenv("AVLD_TRUST_CHECKPOINTS", "1")
    assert isinstance(api.load_encoder(tmp_path / "full.pt", tmp_path / "none.yaml", tmp_path), torch.nn.Sequential)


def test_engine_cache_fingerprint_changes_with_weights():
    net = _tiny(4)
    f0 = api._weights_fingerprint(net)
    assert f0 == api._weights_fingerprint(net)
    with torch.no_grad():
        net[0].weight.mul_(1.5)                                  # in-place update of the same object
    assert api._weights_fingerprint(net) != f0
