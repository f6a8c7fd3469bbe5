#!/usr/bin/env python
"""Drop-in for the reference's ``08b_fit_map_detector.py`` (same flags, messages and exit codes); the work is done by
``amphibian_vae_latent_detector_b200.cli.main_08b`` on the GPU library."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from amphibian_vae_latent_detector_b200.cli import main_08b  # noqa: E402

if __name__ == "__main__":
    main_08b(here=Path(__file__).resolve().parent)
