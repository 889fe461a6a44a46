"""Role hand-over timeline of the persistent GEMM kernel (CTA 0, first tiles), from a -DGEMM_TRACE build:

    bash profiles/build_variant.sh trace -DGEMM_TRACE
    PP_B200_LIB=powerpaint_b200/_variants/libtrace.so python profiles/gemm_trace.py [linear|qk|conv|geglu]

Columns (cycles relative to the producer's first TMA issue of tile 0):
  P0/P1  producer: first / last TMA issue of the tile      M2 accumulator stage free   M3 first smem stage full
  M4 last k-iteration's stage full (its MMAs issue now)    E5 accumulator ready (epilogue)   E6 tile math done + staged
  E9 previous staged tile read by the TMA unit   B1 first barrier   STG staging stores issued   FNC proxy fence done
  E7 second barrier (all warps staged)   E8 TMA stores issued   E10 end of tile   SET next tile set up, about to wait for its accumulator
"""
import ctypes
import math
import sys

import torch

sys.path.insert(0, ".")
from powerpaint_b200 import _native as nat  # noqa: E402
from powerpaint_b200 import ops  # noqa: E402

dev = "cuda"
BF = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "linear"
if what == "conv":
    nb, h, w, cin, cout = 16, 64, 64, 320, 320
    x = torch.randn(nb, h * w, cin, device=dev, generator=g).to(BF)
    wt = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin))
    out = torch.empty(nb, h * w, cout, device=dev, dtype=BF)
    d = ops.gemm_desc(a0=x, w=wt, out=out, N_=cout, a_mode=nat.PP_A_CONV3X3, c0=cin, nb=nb, h=h, w_=w,
                      bias=torch.zeros(cout, device=dev))
else:
    M, K = 65536, 320
    N = {"linear": 320, "qk": 640, "geglu": 2560}[what]
    a = torch.randn(M, K, device=dev, generator=g).to(BF)
    if what == "geglu":
        wgt, b = ops.pack_geglu_weight(torch.randn(N, K, device=dev, generator=g) / math.sqrt(K), torch.zeros(N, device=dev), 128)
        out = torch.empty(M, N // 2, device=dev, dtype=BF)
        d = ops.gemm_desc(a0=a, w=wgt, out=out, N_=N, M=M, bias=b, epilogue=nat.PP_EPI_GEGLU, block_n=128)
    else:
        wgt = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).to(BF)
        out = torch.empty(M, N, device=dev, dtype=BF)
        d = ops.gemm_desc(a0=a, w=wgt, out=out, N_=N, M=M, bias=torch.zeros(N, device=dev))
for _ in range(3):
    ops.run(d)
torch.cuda.synchronize()
lib = ctypes.CDLL(str(nat.lib_path()))
n = 64 * 32
buf = (ctypes.c_longlong * n)()
rc = lib.pp_debug_gemm_trace(buf, n)
assert rc == 0, rc
t = [[buf[i * 32 + j] for j in range(19)] for i in range(10)]
t0 = t[0][0]
names = ["P0", "P1", "M2", "M3", "M4", "E5", "E6", "E7", "E8", "E9", "E10", "B1", "STG", "FNC", "SET", "TOP", "T16", "T17", "T18"]
print(what, " ".join(f"{n_:>7s}" for n_ in names))
for i, row in enumerate(t):
    if row[0] == 0 and i > 0:
        break
    print(f"tile {i}", " ".join(f"{(v - t0) if v else 0:7d}" for v in row))
