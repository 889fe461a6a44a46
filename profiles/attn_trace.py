"""One launch of the d=40 self-attention at the C2 shape: the target of `ncu --set full -k regex:attn2`
captures (and of temporary clock64 phase traces while the kernel was being restructured)."""
import math, sys, torch
sys.path.insert(0, ".")
from powerpaint_b200 import ops
dev, BF = "cuda", torch.bfloat16
B, H, d, n = 16, 8, 40, 4096
C = H * d
q = torch.randn(B, n, C, device=dev).to(BF); k = torch.randn(B, n, C, device=dev).to(BF)
vt = torch.randn(B, C, n, device=dev).to(torch.float16); out = torch.empty(B, n, C, device=dev, dtype=BF)
dd = ops.attn_desc(q=q, k=k, vt=vt, out=out, batch=B, heads=H, d=d, nq=n, nk=n, q_ld=C, k_ld=C, vt_ld=n, o_ld=C,
                   q_batch_stride=n * C, k_batch_stride=n * C, scale=1 / math.sqrt(d))
ops.run(dd); torch.cuda.synchronize()
