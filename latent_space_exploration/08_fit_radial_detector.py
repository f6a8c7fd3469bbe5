#!/usr/bin/env python
"""Drop-in for the reference's ``08_fit_radial_detector.py`` (same flags, messages and exit codes); the work is done by
``amphibian_vae_latent_detector_b200.cli.main_08`` on the GPU library."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from amphibian_vae_latent_detector_b200.cli import main_08  # noqa: E402

if __name__ == "__main__":
    main_08(here=Path(__file__).resolve().parent)
