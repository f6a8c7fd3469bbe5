"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the encode+detect hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or the thing timed as the CPU
baseline), never as an implementation the product path falls back to.

Contents
--------
``librosa_port``  restatement of the third-party arithmetic the reference calls but does
                  not vendor: ``librosa==0.9.2`` (``load``, ``feature.melspectrogram``,
                  ``power_to_db``, ``filters.mel``) and ``soundfile==0.13.1`` (``write`` of a
                  float array to ``.wav`` = PCM_16).  Pinned against torchaudio in
                  ``tests/test_oracle_pinning.py``.
``shims``         installs those restatements (plus tiny ``omegaconf``/``hydra`` stand-ins)
                  into ``sys.modules`` so the *unmodified* reference files import.
``ref_import``    loads the reference's own modules by path from ``/root/reference``
                  (build container only; the GPU box has no reference).
``hotpath``       numpy restatement of the reference's own functions on the path, each
                  citing the reference file:line it follows.  This is what travels to the
                  GPU box.  Pinned against the real reference code by
                  ``tests/golden/*.npz`` (made by ``oracle/make_golden.py``).

Parity status: the reference ships no golden vectors / tests, so
the oracle is pinned against outputs of the reference's own code executed in the build
container (fixtures + generating script committed).  The third-party pieces (librosa,
soundfile) and the encoder architecture are absent from the reference tree:
**for those, parity is unpinned against the thesis artefacts** and pinned only against
torchaudio (mel/STFT) and our stand-in encoder.
"""
