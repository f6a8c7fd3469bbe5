"""Drop-in script / module surface of the reference's ``latent_space_exploration`` package, backed by the B200 library
(``amphibian_vae_latent_detector_b200``).  Same file names, flags and function names; see INTEGRATION.md."""
