"""CPU: the script / flag surface ("the reference's CLI flags") -- every flag, type and default of
the reference's argparse parsers is accepted by ours; launcher files carry the reference's names; project-root search and
the missing-file messages behave as 08:69-102.  The comparison with the reference's own parsers runs only where
/root/reference exists (this container); the expected tables below were read from it and are checked everywhere."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

from amphibian_vae_latent_detector_b200 import cli

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/latent_space_exploration")
MEL = {"sr": 48000, "n_mels": 64, "target_frames": 192, "fmin": 150.0, "fmax": 15000.0, "hop_length": 384, "n_fft": 2048}
EXPECTED = {
    "00": (cli.parser_00, [], {"base_dir": "latent_space_exploration", "sr": 48000}),
    "07": (cli.parser_07, ["--wav", "x.wav"],
           {**MEL, "wav": "x.wav", "encoder": None, "encoder_config": None, "device": "cpu", "duration": 3.0, "auto_frames": False,
            "auto_max_frames": 512, "auto_step": 8, "jsonl": False, "precision": 6}),
    "08": (cli.parser_08, ["--root", "train_chunks"],
           {**MEL, "config": "config.json", "root": "train_chunks", "q_in": 0.95, "q_out": 0.01, "device": "cpu", "encoder_pt": None,
            "encoder_yaml": None, "max_per_class": 0, "seed": 123, "cache": False}),
    "09": (cli.parser_09, ["--wav", "x.wav"],
           {**MEL, "wav": "x.wav", "config": None, "encoder_pt": None, "encoder_yaml": None, "device": "cpu"}),
    "10": (cli.parser_10, [], {**MEL, "root": None, "config": None, "encoder_pt": None, "encoder_yaml": None, "device": "cpu"}),
}
SCRIPTS = {"00": "00_normalize_dataset_rms.py", "07": "07_encode_wav_to_latent.py", "08": "08_fit_radial_detector.py",
           "09": "09_evaluate_wav_detection.py", "10": "10_benchmark_folder_detection.py"}


@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_parser_defaults(key):
    make, argv, want = EXPECTED[key]
    assert vars(make().parse_args(argv)) == want


def _reference_parser_actions(key: str):
    """The (dest, option strings, type, default, nargs/const kind) of every add_argument call in a reference script,
    collected by running its parse_args()/main() with argparse patched to capture the parser instead of parsing."""
    import argparse
    sys.path.insert(0, str(REPO))
    from oracle import ref_import
    captured = {}

    class Stop(Exception):
        pass

    def fake_parse(self, *a, **k):
        captured["parser"] = self
        raise Stop

    orig = argparse.ArgumentParser.parse_args
    argparse.ArgumentParser.parse_args = fake_parse
    try:
        mod = ref_import.load(key)
        for fn in ("parse_args", "main"):
            if "parser" not in captured and hasattr(mod, fn):
                try:
                    getattr(mod, fn)()
                except Stop:
                    pass
    finally:
        argparse.ArgumentParser.parse_args = orig
    return {a.dest: (tuple(a.option_strings), a.type, a.default, type(a).__name__, a.required)
            for a in captured["parser"]._actions if a.dest != "help"}


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_parser_equals_reference_parser(key):
    make = EXPECTED[key][0]
    ours = {a.dest: (tuple(a.option_strings), a.type, a.default, type(a).__name__, a.required)
            for a in make()._actions if a.dest != "help"}
    theirs = _reference_parser_actions(key)
    assert ours == theirs


def test_launchers_exist_under_reference_names():
    for name in [*SCRIPTS.values(), "run_qout_grid.py", "map_detector_core.py"]:
        assert (REPO / "latent_space_exploration" / name).is_file(), name
    if REF.exists():
        for name in SCRIPTS.values():
            assert (REF / name).is_file()


def test_find_project_root(tmp_path):
    proj = tmp_path / "p"
    (proj / "downloaded_models").mkdir(parents=True)
    deep = proj / "latent_space_exploration" / "a" / "b"
    deep.mkdir(parents=True)
    assert cli.find_project_root(deep) == proj.resolve()
    lone = tmp_path / "elsewhere" / "x"
    lone.mkdir(parents=True)
    assert cli.find_project_root(lone) == lone.resolve()


def _project(tmp_path, with_encoder=True):
    proj = tmp_path / "proj"
    (proj / "latent_space_exploration" / "train_chunks").mkdir(parents=True)
    enc = proj / "downloaded_models" / "bird_net_vae_audio_splitted_encoder_v0"
    enc.mkdir(parents=True)
    if with_encoder:
        (enc / "model.pt").write_bytes(b"")
        (enc / "bird_net_vae_audio_splitted.yaml").write_text("encoder: {}\n")
    (proj / "config.json").write_text(json.dumps({"species": ["A", "B"]}))
    return proj


def test_08_argument_and_file_errors(tmp_path, monkeypatch):
    proj = _project(tmp_path, with_encoder=False)
    here = proj / "latent_space_exploration"
    monkeypatch.chdir(proj)
    with pytest.raises(SystemExit, match="--q-in debe estar"):
        cli.main_08(["--root", "train_chunks", "--q-in", "1.5"], here=here)
    with pytest.raises(SystemExit, match="--q-out debe estar"):
        cli.main_08(["--root", "train_chunks", "--q-out", "0"], here=here)
    with pytest.raises(SystemExit, match="No existe config.json"):
        cli.main_08(["--root", "train_chunks", "--config", "nope.json"], here=here)
    with pytest.raises(SystemExit, match="No existe chunks_dir"):
        cli.main_08(["--root", "no_such_dir"], here=here)
    with pytest.raises(SystemExit, match="No encontré encoder .pt"):
        cli.main_08(["--root", "train_chunks"], here=here)           # relative root found under latent_space_exploration/
    (proj / "config.json").write_text(json.dumps({"species": "A"}))
    with pytest.raises(SystemExit, match="campo 'species'"):
        cli.main_08(["--root", "train_chunks"], here=here)


def test_10_missing_root_and_dispatcher(tmp_path):
    proj = _project(tmp_path)
    with pytest.raises(FileNotFoundError, match="No existe root"):
        cli.main_10(["--root", str(proj / "nope")], here=proj / "latent_space_exploration")
    with pytest.raises(SystemExit, match="usage"):
        cli.main(["frobnicate"])


def test_map_detector_core_module_surface(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_our_map_detector_core", REPO / "latent_space_exploration" / "map_detector_core.py")
    core = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(core)
    for name in ("find_project_root", "resolve_default_config", "resolve_default_encoder_pt", "resolve_default_encoder_yaml",
                 "load_json", "save_json", "summarize_1d", "load_yaml_cfg", "pick_encoder_cfg", "split_model_and_state",
                 "build_nn_module", "load_encoder", "crop_or_pad_time", "wav_to_mel", "encode_wav_to_latent", "inv_and_logdet",
                 "gaussian_logpdf_from_precision", "get_priors_from_map_meta", "get_chunk_seconds_for_map",
                 "read_map_detector_params"):
        assert callable(getattr(core, name)), name
    x = np.arange(101, dtype=np.float64)
    assert core.summarize_1d(x) == {"min": 0.0, "p05": 5.0, "p50": 50.0, "p95": 95.0, "max": 100.0}
    assert all(np.isnan(v) for v in core.summarize_1d(np.zeros(0)).values())
    core.save_json(tmp_path / "o.json", {"ñ": 1})
    assert core.load_json(tmp_path / "o.json") == {"ñ": 1} and "ñ" in (tmp_path / "o.json").read_text(encoding="utf-8")
    with pytest.raises(FileNotFoundError):
        core.resolve_default_encoder_pt(tmp_path)
