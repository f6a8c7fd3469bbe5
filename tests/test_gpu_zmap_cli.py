"""GPU, row N1 at file level: ``08b_fit_map_detector.py`` -> ``10b_benchmark_folder_detection_map.py`` ->
``09n_evaluate_wav_detection.py`` through ``cli.main_08b / main_10b / main_09n`` on the real engine, against the artefacts the
reference's own ``main()``s wrote for the same seeded WAV tree (tests/golden/pipeline_map, oracle/make_golden_pipeline_map.py).
The comparison is the one of tests/test_map_cli_host.py (which runs the same glue on the numpy oracle, tightly); here the
latents come from the CUDA path, so numbers are held to the north-star tolerance (latents within 1e-3 of max|z|) propagated
through the fit: means 2e-3 of their largest entry, covariances 1e-2 (error ~ 2 delta / sigma), the precision matrices
(n = 10 per class in D = 128, kept invertible by the shrinkage) 3e-2, scores / log-determinants / tau 2e-3 relative or 2.0
absolute; at most 3 of 24 files may change class, and only then is ``summary.txt`` allowed to differ.  These limits were
checked on the CPU by perturbing the oracle's latents with noise of 1e-4 ... 1e-3 of max|z|: all pass."""
import json
from pathlib import Path

import pytest

from test_map_cli_host import GOLD, check_case_against_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def project(tmp_path_factory):
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import synth
    pm = json.loads((GOLD.parent / "pipeline" / "meta.json").read_text())
    root = tmp_path_factory.mktemp("mapproj_gpu")
    lse = root / "latent_space_exploration"
    (root / "downloaded_models").mkdir()
    mdir = root / "models" / "bird_net_vae_audio_splitted_encoder_v0"          # map_detector_core.py:64-77
    mdir.mkdir(parents=True)
    (mdir / "model.pt").write_bytes(b"")
    (mdir / "bird_net_vae_audio_splitted.yaml").write_text("encoder: {}\n")
    synth.write_wav_tree(lse / "raw" / "train_chunks", pm["species"], pm["n_train"], pm["length"], seed=pm["seed_train"],
                         special_every=17)
    synth.write_wav_tree(lse / "raw" / "val_chunks", pm["species"], pm["n_val"], pm["length"], seed=pm["seed_val"],
                         special_every=11)
    for split in ("train_chunks", "val_chunks"):
        api.process_folder(lse / "raw" / split, lse / split, sr=48000)         # bit-identical files (test_gpu_pipeline)
    return root, lse


@pytest.mark.parametrize("case", ["lda_diag_tau", "qda_full_uniform"])
def test_map_scripts_match_reference_artifacts(project, standin_encoder, monkeypatch, capsys, case):
    from amphibian_vae_latent_detector_b200 import reference_api as api
    root, lse = project
    monkeypatch.setattr(api, "load_encoder", lambda *a, **k: standin_encoder)  # the thesis checkpoint is not public
    cfg = check_case_against_reference(root, lse, case, capsys, tol=2e-3, cov_tol=1e-2, prec_tol=3e-2, score_tol=2.0)
    assert Path(cfg["map_detector"]["meta_fit"]["chunks_dir"]).name == "train_chunks"


def test_stream_map_window_equals_chunk_file_workflow(tmp_path, standin_encoder):
    """BASELINE configs[4], 09n / 10b style: window i of ``stream.detect_long_wav_map`` == ``process_folder`` (00) on a chunk file
    holding the same samples, then ``MapDetectorSession.predict_many`` (10b) -- same kernels, only the batch differs."""
    import wave

    import numpy as np
    import torch

    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import stream, synth
    params = np.load(GOLD / "lda_diag_tau" / "params.npz")
    ref = json.loads((GOLD / "lda_diag_tau" / "config_used.json").read_text(encoding="utf-8"))
    md = ref["map_detector"]
    names = list(md["means"])
    for key in ("means", "precision"):                                           # the reference-made MAP parameters
        md[key] = {sp: params[key][i].tolist() for i, sp in enumerate(names)}
    md.pop("cov")
    (tmp_path / "config.json").write_text(json.dumps(ref))
    L, n_full, tail = 144000, 7, 61000
    x, _ = synth.make_chunks(n_full + 1, L, seed=91, special_every=4)
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).numpy()

    def write(path, samples):
        path.parent.mkdir(parents=True, exist_ok=True)
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(48000)
            w.writeframes(samples.astype("<i2").tobytes())

    write(tmp_path / "long.wav", np.concatenate([pcm[:n_full].reshape(-1), pcm[n_full, :tail]]))
    for i in range(n_full):
        write(tmp_path / "raw" / "sp" / f"w{i:02d}.wav", pcm[i])
    pad = np.zeros(L, np.int16)
    pad[:tail] = pcm[n_full, :tail]
    write(tmp_path / "raw" / "sp" / f"w{n_full:02d}.wav", pad)
    api.process_folder(tmp_path / "raw", tmp_path / "norm")
    sess = api.MapDetectorSession(tmp_path, tmp_path / "config.json", tmp_path / "x.pt", tmp_path / "x.yaml", "cuda")
    sess.set_params(api.load_json(tmp_path / "config.json"))
    sess.encoder = standin_encoder
    by_file = sess.predict_many(sorted((tmp_path / "norm" / "sp").glob("*.wav")))
    res = stream.detect_long_wav_map(tmp_path / "long.wav", config_path=tmp_path / "config.json", encoder=standin_encoder,
                                     slab_windows=3)
    assert [w.start_s for w in res] == [3.0 * i for i in range(n_full + 1)]
    for w, (det, sp, best) in zip(res, by_file):
        assert (w.detected, w.species) == (det, sp)
        assert abs(w.best_score - best) <= 1e-5 * abs(best)
