"""AutoencoderKL (SD-1.5 VAE) on the repo's CUDA kernels — SURVEY.md §8f row 1.

Drop-in for the diffusers class the reference pipelines inject (`vae.encode(x).latent_dist.sample(generator)`,
`vae.decode(z, return_dict=False)[0]`, `vae.config.scaling_factor`; call sites `_encode_vae_image`
powerpaint/pipelines/pipeline_PowerPaint.py:657-669, decode :1051; pipeline_PowerPaint_Brushnet_CA.py:1338-1341,
:1476) with diffusers state-dict names, so `vae/diffusion_pytorch_model.safetensors` loads unchanged.

`encode` / `decode` replay a recorded `pp_program` (CUDA graph) of `VaeEngine`:

  reference (diffusers AutoencoderKL)                  here
  ---------------------------------------------------- ----------------------------------------------------
  ResnetBlock2D(temb=None): GN(eps 1e-6)+SiLU+conv x2   pp_group_norm (statistics from the producing conv's
                                                        epilogue) -> implicit-GEMM 3x3 conv, shortcut as res1
  Downsample2D(padding=0): F.pad(0,1,0,1) + conv s2     PP_A_CONV3X3_S2P0 (the pad is TMA's out-of-bounds zero)
  Upsample2D: nearest 2x + conv                         pp_upsample_nearest + conv
  mid-block Attention, 1 head x 512 channels            to_q|to_k GEMM + V^T GEMM, then per sample
                                                        S = QK^T (fp32) -> pp_softmax_rows -> O = PV, to_out(+res)
  conv_out -> quant_conv (1x1)                          folded on the host into ONE 3x3 conv (exact: both linear)
  post_quant_conv (1x1) -> conv_in                      1x1 GEMM, then conv (not folded: the bias meets zero padding)

There is no CPU path: parameters must live on a CUDA device (the torch restatement used as the checker is
`oracle/vae.py`).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops
from ..engine import BF16, NetEngine, Plan, _ceil


# --------------------------------------------------------------------------- parameter layout
def vae_param_shapes(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                     layers_per_block=2) -> "OrderedDict[str, Tuple[int, ...]]":
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = tuple(block_out_channels)

    def conv(name, cout, cin, k):
        d[name + ".weight"] = (cout, cin, k, k)
        d[name + ".bias"] = (cout,)

    def norm(name, c):
        d[name + ".weight"] = (c,)
        d[name + ".bias"] = (c,)

    def lin(name, cout, cin):
        d[name + ".weight"] = (cout, cin)
        d[name + ".bias"] = (cout,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def mid(p, c):
        a = p + ".attentions.0"
        norm(a + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{a}.{n}", c, c)
        lin(a + ".to_out.0", c, c)
        resnet(p + ".resnets.0", c, c)
        resnet(p + ".resnets.1", c, c)

    conv("encoder.conv_in", boc[0], in_channels, 3)
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        c = co
    mid("encoder.mid_block", c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * latent_channels, c, 3)
    rev = list(reversed(boc))
    conv("decoder.conv_in", rev[0], latent_channels, 3)
    mid("decoder.mid_block", rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        c = co
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", out_channels, c, 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    conv("post_quant_conv", latent_channels, latent_channels, 1)
    return d


# --------------------------------------------------------------------------- engine
class VaeEngine(NetEngine):
    """Records VAE encode / decode as `pp_program`s over the conv / GroupNorm / GEMM kernels of the denoising
    step (buffer recycling, epilogue GroupNorm statistics and weight packing are inherited from `NetEngine`)."""

    GN_EPS = 1e-6
    MAX_PLANS = 2

    def __init__(self, cfg: SimpleNamespace, state_dict: Dict[str, torch.Tensor], device):
        self.cfg = cfg  # needs .norm_num_groups for _gn
        self.kind = "vae"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VaeEngine needs a CUDA device: there is no CPU path")
        N.lib()
        self._sd = state_dict
        self._w: Dict[str, torch.Tensor] = {}
        self._plans: "OrderedDict[tuple, Plan]" = OrderedDict()

    # ---- blocks
    def _vae_resnet(self, plan, prog, name, x, nb, h, w, cout):
        hw = h * w
        cin = x.shape[-1]
        n1 = self._gn(plan, prog, x, None, nb, hw, name + ".norm1", self.GN_EPS, True)
        t1 = self._conv3(plan, prog, n1, nb, h, w, name + ".conv1", cout, stats=True)
        self._free(plan, n1)
        n2 = self._gn(plan, prog, t1, None, nb, hw, name + ".norm2", self.GN_EPS, True)
        self._free(plan, t1)
        sc, owned = x, None
        if cin != cout:
            sc = owned = self._linear(plan, prog, x, nb * hw, name + ".conv_shortcut", cout,
                                      bias=self.vec(name + ".conv_shortcut.bias"))
        out = self._conv3(plan, prog, n2, nb, h, w, name + ".conv2", cout, res1=sc, stats=True)
        self._free(plan, n2, owned)
        return out

    def _vae_attention(self, plan, prog, name, x, nb, h, w):
        """one head over all C channels: softmax(q k^T / sqrt(C)) v, then to_out + residual"""
        hw = h * w
        M = nb * hw
        C = x.shape[-1]
        hw_ld = _ceil(hw, 8)
        g = self._gn(plan, prog, x, None, nb, hw, name + ".group_norm", self.GN_EPS, False)
        wqk = self._cached("vqk:" + name, lambda: torch.cat(
            [self._raw(name + ".to_q.weight"), self._raw(name + ".to_k.weight")], 0).to(BF16).contiguous())
        bqk = self._cached("vqkb:" + name, lambda: torch.cat(
            [self._raw(name + ".to_q.bias"), self._raw(name + ".to_k.bias")], 0).contiguous())
        qk = self._linear(plan, prog, g, M, None, 2 * C, w=wqk, bias=bqk)
        # V^T [nb, C, hw_ld] and P [hw, hw_ld]: zero-initialised once, the kernels never write the pad columns
        vt = torch.zeros(nb, C, hw_ld, dtype=BF16, device=self.device)
        pm = torch.zeros(hw, hw_ld, dtype=BF16, device=self.device)
        sm = torch.empty(hw, hw, dtype=torch.float32, device=self.device)
        plan.bytes += vt.numel() * 2 + pm.numel() * 2 + sm.numel() * 4
        prog.add(ops.gemm_desc(a0=g, w=self.w_linear(name + ".to_v"), out=vt, N_=C, M=M, bias=self.vec(name + ".to_v.bias"),
                               epilogue=N.PP_EPI_TRANSPOSED, t_rows=hw, t_ld=hw_ld))
        self._free(plan, g)
        o = self._buf(plan, M, C)
        scale = 1.0 / math.sqrt(C)
        for b in range(nb):
            rows = qk[b * hw:(b + 1) * hw]
            prog.add(ops.gemm_desc(a0=rows, c0=C, lda0=2 * C, w=rows[:, C:], ldb=2 * C, out=sm, N_=hw, M=hw, alpha=scale,
                                   out_fp32=True))
            prog.add_softmax_rows(sm, pm, hw, hw, hw, hw_ld)
            prog.add(ops.gemm_desc(a0=pm, c0=hw_ld, lda0=hw_ld, w=vt[b], ldb=hw_ld, out=o[b * hw:(b + 1) * hw], N_=C, M=hw))
        self._free(plan, qk)
        out = self._linear(plan, prog, o, M, name + ".to_out.0", C, bias=self.vec(name + ".to_out.0.bias"), res1=x,
                           stats_hw=hw)
        self._free(plan, o)
        return out

    def _vae_mid(self, plan, prog, name, x, nb, h, w):
        C = x.shape[-1]
        a = self._vae_resnet(plan, prog, name + ".resnets.0", x, nb, h, w, C)
        self._free(plan, x)
        b = self._vae_attention(plan, prog, name + ".attentions.0", a, nb, h, w)
        self._free(plan, a)
        c = self._vae_resnet(plan, prog, name + ".resnets.1", b, nb, h, w, C)
        self._free(plan, b)
        return c

    def _get_plan(self, key, build):
        p = self._plans.pop(key, None)
        if p is None:
            p = build()
            p.program.build_graph()
            while len(self._plans) >= self.MAX_PLANS:
                self._plans.popitem(last=False)
        self._plans[key] = p
        return p

    # ---- encoder
    def encode_plan(self, nb: int, H: int, W: int) -> Plan:
        return self._get_plan(("enc", nb, H, W), lambda: self._build_encode(nb, H, W))

    def _build_encode(self, nb, H, W) -> Plan:
        cfg = self.cfg
        boc = cfg.block_out_channels
        plan = Plan()
        prog = plan.program = ops.Program()
        x_in = torch.zeros(nb, H * W, 8, dtype=BF16, device=self.device)
        plan.inputs["x"] = x_in
        h, w = H, W
        cur = self._conv3(plan, prog, x_in, nb, h, w, "encoder.conv_in", boc[0], pad_in=8, stats=True)
        for i, co in enumerate(boc):
            for j in range(cfg.layers_per_block):
                nxt = self._vae_resnet(plan, prog, f"encoder.down_blocks.{i}.resnets.{j}", cur, nb, h, w, co)
                self._free(plan, cur)
                cur = nxt
            if i != len(boc) - 1:
                if h < 2 or w < 2:
                    raise ValueError(f"image {H}x{W} is too small for the VAE encoder")
                nxt = self._conv3(plan, prog, cur, nb, h, w, f"encoder.down_blocks.{i}.downsamplers.0.conv", co,
                                  a_mode=N.PP_A_CONV3X3_S2P0, stats=True)
                self._free(plan, cur)
                cur = nxt
                h, w = h // 2, w // 2
        cur = self._vae_mid(plan, prog, "encoder.mid_block", cur, nb, h, w)
        gno = self._gn(plan, prog, cur, None, nb, h * w, "encoder.conv_norm_out", self.GN_EPS, True)
        self._free(plan, cur)
        # conv_out followed by the 1x1 quant_conv: both linear, folded into one conv on the host (fp32)
        lc2 = 2 * cfg.latent_channels

        def folded():
            wq = self._raw("quant_conv.weight").reshape(lc2, lc2)
            wc = self._raw("encoder.conv_out.weight")
            wf = torch.einsum("oc,cikl->oikl", wq, wc)
            bf = wq @ self._raw("encoder.conv_out.bias") + self._raw("quant_conv.bias")
            return ops.pack_conv3x3_weight(wf), bf.contiguous()
        wf, bf = self._cached("enc_out_folded", folded)
        moments = torch.empty(nb, h * w, lc2, dtype=torch.float32, device=self.device)
        prog.add(ops.gemm_desc(a0=gno, w=wf, out=moments, N_=lc2, a_mode=N.PP_A_CONV3X3, c0=gno.shape[-1], nb=nb, h=h,
                               w_=w, bias=bf, out_fp32=True))
        self._free(plan, gno)
        plan.outputs["moments"] = moments
        plan.outputs["hw"] = (h, w)
        return plan

    # ---- decoder
    def decode_plan(self, nb: int, h: int, w: int) -> Plan:
        return self._get_plan(("dec", nb, h, w), lambda: self._build_decode(nb, h, w))

    def _build_decode(self, nb, h, w) -> Plan:
        cfg = self.cfg
        rev = list(reversed(cfg.block_out_channels))
        lc = cfg.latent_channels
        plan = Plan()
        prog = plan.program = ops.Program()
        z_in = torch.zeros(nb, h * w, 8, dtype=BF16, device=self.device)
        plan.inputs["z"] = z_in
        # post_quant_conv (1x1, latent -> latent): its own GEMM; channels >= lc of the output stay zero
        zq = torch.zeros(nb, h * w, 8, dtype=BF16, device=self.device)
        wpq = self._cached("pq:w", lambda: torch.nn.functional.pad(
            self._raw("post_quant_conv.weight").reshape(lc, lc), (0, 8 - lc)).to(BF16).contiguous())
        prog.add(ops.gemm_desc(a0=z_in, w=wpq, out=zq, N_=lc, M=nb * h * w, bias=self.vec("post_quant_conv.bias"), ldc=8))
        cur = self._conv3(plan, prog, zq, nb, h, w, "decoder.conv_in", rev[0], pad_in=8, stats=True)
        cur = self._vae_mid(plan, prog, "decoder.mid_block", cur, nb, h, w)
        for i, co in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                nxt = self._vae_resnet(plan, prog, f"decoder.up_blocks.{i}.resnets.{j}", cur, nb, h, w, co)
                self._free(plan, cur)
                cur = nxt
            if i != len(rev) - 1:
                up = self._buf(plan, nb, 4 * h * w, co)
                prog.add_upsample_nearest(cur, up, nb, h, w, co, 2 * h, 2 * w)
                self._free(plan, cur)
                h, w = 2 * h, 2 * w
                cur = self._conv3(plan, prog, up, nb, h, w, f"decoder.up_blocks.{i}.upsamplers.0.conv", co, stats=True)
                self._free(plan, up)
        gno = self._gn(plan, prog, cur, None, nb, h * w, "decoder.conv_norm_out", self.GN_EPS, True)
        self._free(plan, cur)
        # fp32 image, 4 floats per pixel (3 used): the uint8 conversion must not see a bf16-rounded value
        image = torch.zeros(nb, h * w, 4, dtype=torch.float32, device=self.device)
        prog.add(ops.gemm_desc(a0=gno, w=self.w_conv3("decoder.conv_out"), out=image, N_=cfg.out_channels,
                               a_mode=N.PP_A_CONV3X3, c0=gno.shape[-1], nb=nb, h=h, w_=w,
                               bias=self.vec("decoder.conv_out.bias"), out_fp32=True, ldc=4))
        self._free(plan, gno)
        plan.outputs["image"] = image
        plan.outputs["hw"] = (h, w)
        return plan


# --------------------------------------------------------------------------- diffusers-facing module
class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        dev = generator.device if generator is not None else self.mean.device
        noise = torch.randn(self.mean.shape, generator=generator, device=dev, dtype=self.mean.dtype)
        return self.mean + self.std * noise.to(self.mean.device)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Tuple[int, ...] = (128, 256, 512, 512), layers_per_block: int = 2,
                 norm_num_groups: int = 32, scaling_factor: float = 0.18215):
        super().__init__()
        if in_channels != 3 or out_channels != 3 or latent_channels > 4:
            raise NotImplementedError("the kernel path is built for RGB images and <= 4 latent channels (SD VAE)")
        if any(c % 8 for c in block_out_channels):
            raise ValueError("block_out_channels must be multiples of 8")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      latent_channels=latent_channels, block_out_channels=tuple(block_out_channels),
                                      layers_per_block=layers_per_block, scaling_factor=scaling_factor,
                                      norm_num_groups=norm_num_groups)
        for name, shape in vae_param_shapes(in_channels, out_channels, latent_channels, block_out_channels,
                                            layers_per_block).items():
            self._register(name, torch.zeros(shape))
        self._engine: Optional[VaeEngine] = None
        self._out_dtype = torch.float32
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _register(self, name: str, value: torch.Tensor):
        parts = name.split(".")
        mod = self
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))

    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def to(self, *args, **kwargs):
        # like the UNet: compute is always bf16 x bf16 -> fp32; a requested dtype only selects the output dtype
        dtype = kwargs.pop("dtype", None)
        rest = []
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            else:
                rest.append(a)
        if dtype is not None:
            self._out_dtype = dtype
        if not rest and not kwargs:
            return self
        return super().to(*rest, **kwargs)

    @property
    def dtype(self):
        return self._out_dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def engine(self) -> VaeEngine:
        if self._engine is None:
            dev = self.device
            if dev.type != "cuda":
                raise RuntimeError(f"AutoencoderKL must be on a CUDA device (got {dev}): there is no CPU path")
            self._engine = VaeEngine(self.config, {k: v for k, v in self.state_dict().items()}, dev)
        return self._engine

    # ---- encode
    @torch.no_grad()
    def _moments(self, x_nhwc: torch.Tensor, nb: int, H: int, W: int) -> torch.Tensor:
        plan = self.engine().encode_plan(nb, H, W)
        plan.inputs["x"].copy_(x_nhwc.view_as(plan.inputs["x"]))
        plan.program.launch()
        h, w = plan.outputs["hw"]
        return ops.nhwc_to_nchw(plan.outputs["moments"].view(nb, h, w, -1)).to(self._out_dtype)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [B,3,H,W] in [-1,1] -> `.latent_dist` (DiagonalGaussianDistribution over [B,8,H/8,W/8] moments)"""
        nb, c, H, W = x.shape
        self.engine()  # raises on CPU parameters: there is no CPU path
        with torch.cuda.device(self.device):
            xn = ops.nchw_to_nhwc(x.to(self.device).float().contiguous(), 8)
            dist = DiagonalGaussianDistribution(self._moments(xn, nb, H, W))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def encode_uint8(self, image_u8: torch.Tensor, mask: Optional[torch.Tensor] = None) -> DiagonalGaussianDistribution:
        """uint8 NCHW image (hole zeroed where mask >= 0.5) straight into the encoder: `image / 127.5 - 1`,
        binarise, `image * (mask < 0.5)` (pipeline_PowerPaint.py:123-147) and the layout change are one kernel"""
        nb, c, H, W = image_u8.shape
        self.engine()
        with torch.cuda.device(self.device):
            xn = ops.image_preprocess_u8(image_u8.contiguous(), mask, c_pad=8)
            return DiagonalGaussianDistribution(self._moments(xn, nb, H, W))

    # ---- decode
    @torch.no_grad()
    def _decode_nhwc(self, z: torch.Tensor):
        nb, c, h, w = z.shape
        plan = self.engine().decode_plan(nb, h, w)
        zi = plan.inputs["z"]
        zi.copy_(ops.nchw_to_nhwc(z.to(self.device).float().contiguous(), 8).view_as(zi))
        plan.program.launch()
        H, W = plan.outputs["hw"]
        return plan.outputs["image"], nb, H, W

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        self.engine()
        with torch.cuda.device(self.device):
            img, nb, H, W = self._decode_nhwc(z)
            out = ops.nhwc_to_nchw(img.view(nb, H, W, 4), 3).to(self._out_dtype)
        return SimpleNamespace(sample=out) if return_dict else (out,)

    @torch.no_grad()
    def decode_postprocessed(self, z: torch.Tensor, uint8: bool) -> torch.Tensor:
        """decode + `VaeImageProcessor.postprocess` denormalisation in one pass over the image:
        uint8 NHWC [B,H,W,3] (what "pil" output is built from) or fp32 NCHW in [0,1] ("pt" / "np")"""
        self.engine()
        with torch.cuda.device(self.device):
            img, nb, H, W = self._decode_nhwc(z)
            return ops.image_postprocess(img, nb, H, W, uint8=uint8)

    # ---- checkpoint directories in the diffusers layout (`<root>/vae`)
    @classmethod
    def from_config(cls, config: dict) -> "AutoencoderKL":
        if config.get("act_fn", "silu") != "silu":
            raise NotImplementedError(f"config.act_fn = {config['act_fn']!r}: the SD VAE uses silu")
        keys = ("in_channels", "out_channels", "latent_channels", "block_out_channels", "layers_per_block",
                "norm_num_groups", "scaling_factor")
        return cls(**{k: (tuple(config[k]) if isinstance(config[k], list) else config[k]) for k in keys if k in config})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, torch_dtype=None,
                        revision: Optional[str] = None, variant: Optional[str] = None, local_files_only: bool = False,
                        cache_dir: Optional[str] = None, **unused) -> "AutoencoderKL":
        """`config.json` + `diffusion_pytorch_model[.variant].safetensors | .bin`; the deprecated attention parameter
        names of older VAE checkpoints (query / key / value / proj_attn) are renamed like diffusers does"""
        from ..loading import load_json, load_weights, rename_deprecated_vae_attention, resolve_checkpoint_dir

        d = resolve_checkpoint_dir(pretrained_model_name_or_path, subfolder, revision, local_files_only, cache_dir)
        model = cls.from_config(load_json(d, "config.json"))
        model.load_state_dict(rename_deprecated_vae_attention(load_weights(d, variant)), strict=True)
        if torch_dtype is not None:
            model.to(dtype=torch_dtype)
        return model

    def save_pretrained(self, save_directory, variant: Optional[str] = None, **unused):
        import json
        import os

        from ..loading import save_weights

        os.makedirs(save_directory, exist_ok=True)
        cfg = {"_class_name": "AutoencoderKL", "_diffusers_version": "0.27.0", "act_fn": "silu"}
        cfg.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()})
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        save_weights(self.state_dict(), save_directory, variant=variant)

    @classmethod
    def synthetic(cls, seed: int = 4321, tiny: bool = False, **kw) -> "AutoencoderKL":
        """deterministic random weights (no checkpoint reachable offline); `tiny` = small widths for tests"""
        if tiny:
            kw.setdefault("block_out_channels", (16, 32, 32, 32))
            kw.setdefault("norm_num_groups", 8)
            kw.setdefault("layers_per_block", 1)
        m = cls(**kw)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
                elif "norm" in name and name.endswith("weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
        return m.eval()
