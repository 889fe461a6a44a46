"""Isolated launches of the hot kernels at C2 shapes, for `ncu --set full` captures:

    ncu --set full --clock-control none --import-source on -k regex:'gemm_conv|attn_fwd' -c 12 \
        -o gpurun_out/prof python profiles/prof_kernels.py
"""
import math
import sys

import torch

sys.path.insert(0, ".")
from powerpaint_b200 import _native as nat  # noqa: E402
from powerpaint_b200 import ops  # noqa: E402

dev = "cuda"
BF = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)


def conv(nb, h, w, cin, cout, bn=0):
    x = torch.randn(nb, h * w, cin, device=dev, generator=g).to(BF)
    wt = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin))
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(nb, h * w, cout, device=dev, dtype=BF)
    d = ops.gemm_desc(a0=x, w=wt, out=out, N_=cout, a_mode=nat.PP_A_CONV3X3, c0=cin, nb=nb, h=h, w_=w, bias=bias,
                      block_n=bn)
    return lambda: ops.run(d)


def conv_splitk(nb, h, w, cin, cout):
    x = torch.randn(nb, h * w, cin, device=dev, generator=g).to(BF)
    wt = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin))
    out = torch.empty(nb, h * w, cout, device=dev, dtype=BF)
    d = ops.gemm_desc(a0=x, w=wt, out=out, N_=cout, a_mode=nat.PP_A_CONV3X3, c0=cin, nb=nb, h=h, w_=w,
                      bias=torch.zeros(cout, device=dev))
    nbytes, tiles = ops.gemm_splitk_query(d)
    if nbytes > 0:
        ops.attach_splitk(d, torch.empty(nbytes // 4, device=dev), torch.zeros(tiles, dtype=torch.int32, device=dev))
    return lambda: ops.run(d)


def linear(M, K, N, geglu=False, bn=0):
    a = torch.randn(M, K, device=dev, generator=g).to(BF)
    if geglu:
        gbn = bn or 128
        w, b = ops.pack_geglu_weight(torch.randn(N, K, device=dev, generator=g) / math.sqrt(K),
                                     torch.zeros(N, device=dev), gbn)
        out = torch.empty(M, N // 2, device=dev, dtype=BF)
        d = ops.gemm_desc(a0=a, w=w, out=out, N_=N, M=M, bias=b, epilogue=nat.PP_EPI_GEGLU, block_n=gbn)
    else:
        w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).to(BF)
        out = torch.empty(M, N, device=dev, dtype=BF)
        d = ops.gemm_desc(a0=a, w=w, out=out, N_=N, M=M, bias=torch.zeros(N, device=dev), block_n=bn)
    return lambda: ops.run(d)


def attn(B, H, d, nq, nk):
    C = H * d
    q = torch.randn(B, nq, C, device=dev, generator=g).to(BF)
    k = torch.randn(B, nk, C, device=dev, generator=g).to(BF)
    ld = (nk + 7) // 8 * 8
    vt = torch.randn(B, C, ld, device=dev, generator=g).to(torch.float16)
    out = torch.empty(B, nq, C, device=dev, dtype=BF)
    dd = ops.attn_desc(q=q, k=k, vt=vt, out=out, batch=B, heads=H, d=d, nq=nq, nk=nk, q_ld=C, k_ld=C, vt_ld=ld,
                       o_ld=C, q_batch_stride=nq * C, k_batch_stride=nk * C, scale=1 / math.sqrt(d))
    return lambda: ops.run(dd)


def gnorm(nb, hw, c, silu=True):
    x = torch.randn(nb, hw, c, device=dev, generator=g).to(BF)
    y = torch.empty_like(x)
    gamma = torch.ones(c, device=dev); beta = torch.zeros(c, device=dev)
    d = ops.gn_desc(x0=x, x1=None, c0=c, c1=0, y=y, gamma=gamma, beta=beta, batch=nb, hw=hw, groups=32,
                    eps=1e-5, silu=silu)
    return lambda: ops.run(d)


def conv_gn_fused(nb, h, w, c):
    """conv emitting GroupNorm partial sums from its epilogue + the consumer GroupNorm (finalize + apply)"""
    x = torch.randn(nb, h * w, c, device=dev, generator=g).to(BF)
    wt = ops.pack_conv3x3_weight(torch.randn(c, c, 3, 3, device=dev, generator=g) / math.sqrt(9 * c))
    out = torch.empty(nb, h * w, c, device=dev, dtype=BF)
    d = ops.gemm_desc(a0=x, w=wt, out=out, N_=c, a_mode=nat.PP_A_CONV3X3, c0=c, nb=nb, h=h, w_=w,
                      bias=torch.zeros(c, device=dev))
    geo = ops.gemm_stats_geometry(d)
    part = torch.empty(int(geo.bytes) // 4, device=dev)
    ops.attach_chan_stats(d, part)
    y = torch.empty_like(out)
    gd = ops.gn_desc(x0=out, x1=None, c0=c, c1=0, y=y, gamma=torch.ones(c, device=dev), beta=torch.zeros(c, device=dev),
                     batch=nb, hw=h * w, groups=32, eps=1e-5, silu=True, part0=part, geom0=geo)

    def run():
        ops.run(d)
        ops.run(gd)
    return run


def cfg_ddim(B, hw):
    eps = torch.randn(2 * B, hw, 4, device=dev, generator=g)
    lat = torch.randn(B, hw, 4, device=dev, generator=g)
    coef = torch.tensor([[0.6, 0.8, 0.7, 0.71, 0.0, 7.5, 1.0, 1.0]], device=dev)
    nxt = torch.zeros(2 * B, hw, 16, device=dev, dtype=BF)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    d = ops.cfg_ddim_desc(eps=eps, eps_fp32=True, eps_ld=4, latents=lat, coef=coef, step_idx=step, advance_step=False,
                          noise=None, guidance_scale=7.5, do_cfg=True, batch=B, hw=hw, next_in=nxt, next_c=16,
                          n_copies=2, guidance_from_coef=True)
    return lambda: ops.run(d)


def lnorm(rows, c):
    x = torch.randn(rows, c, device=dev, generator=g).to(BF)
    y = torch.empty_like(x)
    gamma = torch.ones(c, device=dev); beta = torch.zeros(c, device=dev)
    return lambda: ops.layer_norm(x, y, gamma, beta, 1e-5)


def ln_block(M, C, fold, part=None):
    """norm1 -> q|k and V^T of one BasicTransformerBlock, either as LayerNorm kernel + two GEMMs or folded:
    the producer (proj_in-like GEMM) emits per-row records and the two consumers finish the normalisation"""
    a = torch.randn(M, C, device=dev, generator=g).to(BF)
    wp = (torch.randn(C, C, device=dev, generator=g) / math.sqrt(C)).to(BF)
    x = torch.empty(M, C, device=dev, dtype=BF)
    pd = ops.gemm_desc(a0=a, w=wp, out=x, N_=C, M=M, bias=torch.zeros(C, device=dev))
    wqk = (torch.randn(2 * C, C, device=dev, generator=g) / math.sqrt(C)).to(BF)
    wv = (torch.randn(C, C, device=dev, generator=g) / math.sqrt(C)).to(BF)
    qk = torch.empty(M, 2 * C, device=dev, dtype=BF)
    hw = 4096 if M % 4096 == 0 else M
    vt = torch.empty(M // hw, C, hw, device=dev, dtype=torch.float16)
    ones, zeros = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    if fold:
        nrec = ops.gemm_row_stats_records(pd)
        rec = torch.empty(M, 2, device=dev)  # {rstd, -rstd * mean} per row, left by the producer
        ops.attach_row_stats(pd, torch.empty(nrec, M, 4, device=dev), rec, torch.zeros(M // 128, dtype=torch.int32, device=dev), 1e-5)
        uqk, uv = wqk.float().sum(1).contiguous(), wv.float().sum(1).contiguous()
        d1 = ops.gemm_desc(a0=x, w=wqk, out=qk, N_=2 * C, M=M, bias=torch.zeros(2 * C, device=dev), ln=(rec, uqk, 1e-5))
        d2 = ops.gemm_desc(a0=x, w=wv, out=vt, N_=C, M=M, bias=zeros, epilogue=nat.PP_EPI_TRANSPOSED, t_rows=hw, t_ld=hw,
                           t_fp16=True, ln=(rec, uv, 1e-5))

        def run():
            ops.run(pd); ops.run(d1); ops.run(d2)
        if part == "producer":
            return lambda: ops.run(pd)
        if part == "qkv":
            ops.run(pd)
            w3 = torch.cat([wqk, wv], 0).contiguous()
            d3 = ops.gemm_desc(a0=x, w=w3, out=qk, N_=3 * C, M=M, bias=torch.zeros(3 * C, device=dev), block_n=160,
                               ln=(rec, w3.float().sum(1).contiguous(), 1e-5), epilogue=nat.PP_EPI_ROWS_THEN_TRANSPOSED,
                               out_t=vt, trans_from_col=2 * C, t_rows=hw, t_ld=hw, t_fp16=True)
            return lambda: ops.run(d3)
        if part == "geglu":
            ops.run(pd)
            wg, bg = ops.pack_geglu_weight(torch.randn(8 * C, C, device=dev, generator=g) / math.sqrt(C),
                                           torch.zeros(8 * C, device=dev), 128)
            ffh = torch.empty(M, 4 * C, device=dev, dtype=BF)
            dg = ops.gemm_desc(a0=x, w=wg, out=ffh, N_=8 * C, M=M, bias=bg, epilogue=nat.PP_EPI_GEGLU, block_n=128,
                               ln=(rec, wg.float().sum(1).contiguous(), 1e-5))
            return lambda: ops.run(dg)
        if part == "qk":
            ops.run(pd)
            return lambda: ops.run(d1)
        if part == "vt":
            ops.run(pd)
            return lambda: ops.run(d2)
    else:
        l1 = torch.empty_like(x)
        d1 = ops.gemm_desc(a0=l1, w=wqk, out=qk, N_=2 * C, M=M)
        d2 = ops.gemm_desc(a0=l1, w=wv, out=vt, N_=C, M=M, epilogue=nat.PP_EPI_TRANSPOSED, t_rows=hw, t_ld=hw, t_fp16=True)

        def run():
            ops.run(pd); ops.layer_norm(x, l1, ones, zeros, 1e-5); ops.run(d1); ops.run(d2)
    return run


def vt_only(M, C):
    x = torch.randn(M, C, device=dev, generator=g).to(BF)
    wv = (torch.randn(C, C, device=dev, generator=g) / math.sqrt(C)).to(BF)
    hw = 4096
    vt = torch.empty(M // hw, C, hw, device=dev, dtype=torch.float16)
    d2 = ops.gemm_desc(a0=x, w=wv, out=vt, N_=C, M=M, epilogue=nat.PP_EPI_TRANSPOSED, t_rows=hw, t_ld=hw, t_fp16=True)
    return lambda: ops.run(d2)


cases = [
    ("conv 320->320 @64x64 b16", conv(16, 64, 64, 320, 320), 2 * 16 * 4096 * 320 * 2880),
    ("conv 640->640 @32x32 b16", conv(16, 32, 32, 640, 640), 2 * 16 * 1024 * 640 * 5760),
    ("conv 1280->1280 @16x16 b16", conv(16, 16, 16, 1280, 1280), 2 * 16 * 256 * 1280 * 11520),
    ("conv 1280->1280 @8x8 b16", conv(16, 8, 8, 1280, 1280), 2 * 16 * 64 * 1280 * 11520),
    ("conv 1280->1280 @8x8 b16 split-K x2", conv_splitk(16, 8, 8, 1280, 1280), 2 * 16 * 64 * 1280 * 11520),
    ("conv 1280->1280 @16x16 b4 (C4/C5)", conv(4, 16, 16, 1280, 1280), 2 * 4 * 256 * 1280 * 11520),
    ("conv 1280->1280 @16x16 b4 split-K x2", conv_splitk(4, 16, 16, 1280, 1280), 2 * 4 * 256 * 1280 * 11520),
    ("conv 2560->1280 @16x16 b16", conv(16, 16, 16, 2560, 1280), 2 * 16 * 256 * 1280 * 23040),
    ("linear 320->320 M=65536", linear(65536, 320, 320), 2 * 65536 * 320 * 320),
    ("linear 320->640 M=65536 (q|k)", linear(65536, 320, 640), 2 * 65536 * 320 * 640),
    ("linear 320->320 transposed fp16 (V^T)", vt_only(65536, 320), 2 * 65536 * 320 * 320),
    ("proj + LayerNorm + q|k + V^T 320 M=65536 [unfused]", ln_block(65536, 320, False), 2 * 65536 * 320 * 320 * 4),
    ("proj + q|k + V^T 320 M=65536 [LayerNorm folded]", ln_block(65536, 320, True), 2 * 65536 * 320 * 320 * 4),
    ("  folded: producer 320->320 + row records (mode 3)", ln_block(65536, 320, True, "producer"), 2 * 65536 * 320 * 320),
    ("  folded: q|k 320->640 (mode 4)", ln_block(65536, 320, True, "qk"), 2 * 65536 * 320 * 640),
    ("  folded: V^T 320->320 transposed (mode 1 + LN)", ln_block(65536, 320, True, "vt"), 2 * 65536 * 320 * 320),
    ("  folded: q|k|v^T 320->960 one launch (mode 5 + LN)", ln_block(65536, 320, True, "qkv"), 2 * 65536 * 320 * 960),
    ("  folded: geglu 320->2560 (mode 2 + LN)", ln_block(65536, 320, True, "geglu"), 2 * 65536 * 320 * 2560),
    ("geglu 320->2560 M=65536", linear(65536, 320, 2560, geglu=True), 2 * 65536 * 320 * 2560),
    ("linear 1280->320 M=65536", linear(65536, 1280, 320), 2 * 65536 * 1280 * 320),
    ("geglu 640->5120 M=16384", linear(16384, 640, 5120, geglu=True), 2 * 16384 * 640 * 5120),
    ("geglu 320->2560 M=65536 tile 256", linear(65536, 320, 2560, geglu=True, bn=256), 2 * 65536 * 320 * 2560),
    ("geglu 640->5120 M=16384 tile 256", linear(16384, 640, 5120, geglu=True, bn=256), 2 * 16384 * 640 * 5120),
    ("geglu 1280->10240 M=4096", linear(4096, 1280, 10240, geglu=True), 2 * 4096 * 1280 * 10240),
    ("geglu 1280->10240 M=4096 tile 256", linear(4096, 1280, 10240, geglu=True, bn=256), 2 * 4096 * 1280 * 10240),
    ("self-attn d40 N=4096 b16", attn(16, 8, 40, 4096, 4096), 4 * 16 * 8 * 4096 * 4096 * 40),
    ("cross-attn d40 N=4096x77 b16", attn(16, 8, 40, 4096, 77), 4 * 16 * 8 * 4096 * 77 * 40),
    ("self-attn d80 N=1024 b16", attn(16, 8, 80, 1024, 1024), 4 * 16 * 8 * 1024 * 1024 * 80),
    # bandwidth kernels: "flops" = bytes moved (stats read + apply read + write), so the last column is TB/s
    ("groupnorm+silu 320ch @64x64 b16 [TB/s]", gnorm(16, 4096, 320), 3 * 16 * 4096 * 320 * 2),
    ("groupnorm+silu 640ch @32x32 b16 [TB/s]", gnorm(16, 1024, 640), 3 * 16 * 1024 * 640 * 2),
    ("layernorm 320ch M=65536 [TB/s]", lnorm(65536, 320), 2 * 65536 * 320 * 2),
    # conv (84 MB in + out) + one-pass GroupNorm (84 MB): the GroupNorm share is (this - plain conv)
    ("conv320+stats+groupnorm @64x64 b16 [fused]", conv_gn_fused(16, 64, 64, 320), 2 * 16 * 4096 * 320 * 2880),
    # 80 B per pixel (BASELINE / DESIGN): eps 32 + latents 32 + next input 16
    ("cfg_ddim C2 8x64x64 [TB/s]", cfg_ddim(8, 4096), 8 * 4096 * 80),
    ("conv 128->128 @512x512 b8 (VAE decoder)", conv(8, 512, 512, 128, 128), 2 * 8 * 262144 * 128 * 1152),
    ("conv 256->256 @512x512 b8 (VAE upsampler)", conv(8, 512, 512, 256, 256), 2 * 8 * 262144 * 256 * 2304),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, fn, flops in cases:
    if only and only not in name:
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:52s} {us:9.1f} us  {flops / us / 1e6:8.2f} {'TB/s' if 'TB/s' in name else 'TFLOP/s'}")
