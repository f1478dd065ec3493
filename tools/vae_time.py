"""Time N back-to-back c2 VAE decodes (and encodes) with CUDA events: A/B runs of kernel variants selected by env."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import candle_video_b200 as cv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
do_enc = len(sys.argv) > 2 and sys.argv[2] == "enc"
dev = torch.device("cuda:0")
vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig())
if do_enc:
    vae.enable_encoder()
vae.init_random(4321)
z = torch.randn(1, 128, 13, 16, 24, device=dev)
ts = torch.tensor([0.05], device=dev)
out = vae.decode(z, ts)
for _ in range(2):
    vae.decode(z, ts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    lib_out = vae.decode(z, ts)
e1.record()
torch.cuda.synchronize()
print(f"decode_ms {e0.elapsed_time(e1) / n:.3f}")
if do_enc:
    clip = torch.tanh(torch.randn(1, 3, 121, 512, 768, device=dev))
    vae.encode(clip)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(max(n // 2, 1)):
        vae.encode(clip)
    e1.record()
    torch.cuda.synchronize()
    print(f"encode_ms {e0.elapsed_time(e1) / max(n // 2, 1):.3f}")
