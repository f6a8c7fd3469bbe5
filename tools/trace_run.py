import torch, sys
sys.path.insert(0, '/root/repo')
from amphibian_vae_latent_detector_b200.engine import Engine
from amphibian_vae_latent_detector_b200 import synth
eng = Engine(0, chunk_len=144000, max_batch=1024)
x, _ = synth.make_chunks(1024, 144000, seed=1, device='cuda')
for _ in range(3):
    f = eng.logmel(x)
torch.cuda.synchronize()
