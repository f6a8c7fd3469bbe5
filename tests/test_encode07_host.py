"""CPU: the output-unpacking / guarded-forward / shape-probe helpers of 07_encode_wav_to_latent.py (07:195-199, :264-352)
against explicit cases and, when the reference tree is present (build container), against the reference's own functions."""
import pytest
import torch

from amphibian_vae_latent_detector_b200 import reference_api as api

OUTPUTS = {
    "tensor": lambda: torch.arange(12.0).view(2, 6),
    "tuple_mu_logvar": lambda: (torch.ones(2, 5), torch.zeros(2, 5)),
    "list_with_junk": lambda: [None, "x", torch.full((1, 3), 2.0)],
    "dict_mu": lambda: {"logvar": torch.zeros(1, 4), "mu": torch.ones(1, 4)},
    "dict_enc": lambda: {"other": torch.zeros(1, 2), "enc": torch.ones(1, 7)},
    "dict_z_first": lambda: {"mu": torch.zeros(1, 4), "z": torch.ones(1, 4)},
    "dict_fallback": lambda: {"a": 3, "b": torch.full((1, 2), 5.0)},
    "btc": lambda: torch.arange(24.0).view(2, 3, 4),
    "btc_empty_t": lambda: torch.zeros(2, 0, 4),
    "four_d": lambda: torch.arange(48.0).view(2, 2, 3, 4),
}
BAD = {"tuple_no_tensor": lambda: (1, "a"), "dict_no_tensor": lambda: {"a": 1}, "string": lambda: "nope"}


def _ref07():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    return ref_import.load("07")


@pytest.mark.parametrize("name", sorted(OUTPUTS))
def test_extract_vector_cases(name):
    got = api.extract_vector(OUTPUTS[name]())
    assert got.ndim == 2 or name == "btc_empty_t"      # 07:289-290 returns the empty [B, 0, C] slice as is
    if name == "tuple_mu_logvar":
        assert torch.equal(got, torch.ones(2, 5))
    if name == "dict_z_first":
        assert torch.equal(got, torch.ones(1, 4))
    if name == "btc":
        assert torch.equal(got, torch.arange(24.0).view(2, 3, 4).mean(dim=1))
    if name == "btc_empty_t":
        assert got.shape == (2, 0, 4) and got.numel() == 0
    if name == "four_d":
        assert got.shape == (2, 24)


@pytest.mark.parametrize("name", sorted(BAD))
def test_extract_vector_rejects(name):
    with pytest.raises(ValueError):
        api.extract_vector(BAD[name]())


@pytest.mark.parametrize("name", sorted(OUTPUTS) + sorted(BAD))
def test_extract_vector_equals_reference(name):
    ref = _ref07()
    make = {**OUTPUTS, **BAD}[name]
    try:
        want = ref.extract_vector(make())
    except ValueError as e:
        with pytest.raises(ValueError) as ours:
            api.extract_vector(make())
        assert str(ours.value) == str(e)
        return
    assert torch.equal(api.extract_vector(make()), want)


class _Net(torch.nn.Module):
    def __init__(self, fail=False, empty=False):
        super().__init__()
        self.conv = torch.nn.Conv2d(1, 2, 3, padding=1)
        self.pool = torch.nn.MaxPool2d(4)
        self.fc = torch.nn.Linear(2 * 4 * 4, 3)
        self.fail, self.empty = fail, empty

    def forward(self, x):
        if self.fail:
            raise RuntimeError("boom")
        h = self.pool(self.conv(x)).flatten(1)
        if self.empty:
            return h[:, :0]
        return self.fc(h), h


def test_try_forward_and_probe():
    net = _Net().eval()
    x = torch.randn(1, 1, 16, 16)
    ok, vec, err = api.try_forward(net, x)
    assert ok and err is None and vec.shape == (1, 3)
    assert api.try_forward(_Net(fail=True), x) == (False, None, "boom")
    ok, vec, err = api.try_forward(_Net(empty=True), x)
    assert (ok, vec) == (False, None) and "vac" in err
    assert api.find_first_linear(net) is net.fc
    with pytest.raises(RuntimeError):
        api.find_first_linear(torch.nn.Sequential(torch.nn.ReLU()))
    ok, shp, err = api.probe_linear_input_shape(net, net.fc, x)
    assert (ok, shp, err) == (True, (1, 32), None)
    ok, shp, err = api.probe_linear_input_shape(net, net.fc, torch.randn(1, 1, 32, 16))     # wrong width: forward fails
    assert ok and shp == (1, 64) and err                                                     # ... after the hook fired
    ok, shp, err = api.probe_linear_input_shape(_Net(fail=True), net.fc, x)
    assert (ok, shp, err) == (False, None, "boom")


def test_try_forward_and_probe_equal_reference():
    ref = _ref07()
    x = torch.randn(1, 1, 16, 16)
    for net in (_Net().eval(), _Net(fail=True), _Net(empty=True)):
        a, b = api.try_forward(net, x), ref.try_forward(net, x)
        assert a[0] == b[0] and a[2] == b[2] and (a[1] is None) == (b[1] is None)
        if a[1] is not None:
            assert torch.equal(a[1], b[1])
    net = _Net().eval()
    for xin in (x, torch.randn(1, 1, 32, 16)):
        a, b = api.probe_linear_input_shape(net, net.fc, xin), ref.probe_linear_input_shape(net, net.fc, xin)
        assert a[:2] == b[:2] and (a[2] is None) == (b[2] is None)
