"""Bring-up probe: does a UMMA smem descriptor whose start is shifted by s rows inside a 128B-swizzled tile read the
right data (a) with base_offset = 0, (b) with base_offset = s?  Decides the design of the halo-reuse convolution."""
import os, sys, subprocess
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    import torch
    from amphibian_vae_latent_detector_b200.engine import Engine
    s = int(os.environ.get("AVLD_DBG_SHIFT", "0"))
    eng = Engine(0, chunk_len=144000, max_batch=8)
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 128, 64, 64
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    C = eng.dbg_gemm(A, B, 1)
    ref = (A.double() @ B.double().T).float()
    ok_rows = M - s
    err = (C[:ok_rows] - ref[s:s + ok_rows]).abs().max().item() / ref.abs().max().item()
    print(f"shift {s} baseoff {os.environ.get('AVLD_DBG_BASEOFF','0')}: rel err of rows [0,{ok_rows}) vs reference rows shifted by {s}: {err:.3e}")
else:
    for s in (0, 1, 2, 3, 5):
        for bo in sorted({0, s}):
            env = dict(os.environ, AVLD_DBG_SHIFT=str(s), AVLD_DBG_BASEOFF=str(bo),
                       AVLD_LIB_PATH=os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'amphibian_vae_latent_detector_b200', 'libavld_bringup.so'))   # the probes live in the bring-up build only
            r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True, timeout=120)
            print((r.stdout.strip() or r.stderr.strip()[-300:]))
