#!/usr/bin/env python
"""Drop-in for the reference's ``run_qout_grid.py`` (same flags, messages and exit codes); the work is done by
``amphibian_vae_latent_detector_b200.cli.main_grid`` on the GPU library."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from amphibian_vae_latent_detector_b200.cli import main_grid  # noqa: E402

if __name__ == "__main__":
    main_grid(here=Path(__file__).resolve().parent)
