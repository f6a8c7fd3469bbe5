#!/usr/bin/env python
"""Drop-in for the reference's ``09n_evaluate_wav_detection.py`` (same flags, messages and exit codes); the work is done by
``amphibian_vae_latent_detector_b200.cli.main_09n`` on the GPU library."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from amphibian_vae_latent_detector_b200.cli import main_09n  # noqa: E402

if __name__ == "__main__":
    main_09n(here=Path(__file__).resolve().parent)
