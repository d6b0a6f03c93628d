"""torch.profiler kernel table of one FixedSpectralBlock spectral-half fwd+bwd step (where the GPU time goes)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensor_cuda_fft_b200 import spectral_conv as sc
dev = torch.device("cuda:0")
B, T, C, K = 64, 1024, 512, 128
blk = sc.FixedSpectralBlock(C, seq_len=T, kernel_len=K, transition_bins=16, dropout=0.0).to(dev).eval()
x2 = torch.randn(B, T, C, device=dev); g2 = torch.randn(B, T, C, device=dev)
def step():
    for p in blk.parameters(): p.grad = None
    xr = x2.detach().requires_grad_(True)
    blk.spectral_half(xr).backward(g2)
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
