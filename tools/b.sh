#!/usr/bin/env bash
# rebuild libavld.so from anywhere
cd "$(dirname "$0")/.." && python -m amphibian_vae_latent_detector_b200.build "$@"
