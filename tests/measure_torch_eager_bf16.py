"""Comparison point, not a test (no `test_` functions): the torch-eager bf16 restatement of the reference
dataflow — the oracle modules run on the GPU in bf16 through cuDNN / cuBLAS / SDPA, the library path the
reference reaches through diffusers — for one C2 UNet forward (batch 16, 64x64 latents). Lives under tests/
because it executes `oracle/`, which only test infrastructure may import.

    python tests/measure_torch_eager_bf16.py        # round 1: 38.8 ms (331 TFLOP/s) on B200
"""
import sys

import torch

sys.path.insert(0, ".")
from oracle.unet import UNet2DConditionOracle, UNetConfig, init_synthetic_  # noqa: E402

dev = torch.device("cuda:0")
FLOPS_PER_SAMPLE = 0.8034e12  # one 64x64-latent SD-1.5 UNet forward

om = init_synthetic_(UNet2DConditionOracle(UNetConfig.sd15(9))).to(dev).to(torch.bfloat16).eval()
x = torch.randn(16, 9, 64, 64, device=dev, dtype=torch.bfloat16)
ctx = torch.randn(16, 77, 768, device=dev, dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        om(x, 500, ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        om(x, 500, ctx)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"torch_eager_restatement (oracle modules, bf16, cuDNN/cuBLAS/SDPA) C2 UNet forward batch 16: {ms:.2f} ms "
      f"= {16 * FLOPS_PER_SAMPLE / (ms / 1e3) / 1e12:.0f} TFLOP/s", flush=True)
