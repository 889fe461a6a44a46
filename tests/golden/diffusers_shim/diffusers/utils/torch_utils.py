def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **freeu_kwargs):
    raise NotImplementedError("FreeU is outside the hot path (the reference never sets s1/s2/b1/b2)")


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers' contract: one generator draws the whole batch, a list draws sample by sample; a CPU generator draws on
    the CPU whatever the target device"""
    import torch

    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        return torch.cat([torch.randn(one, generator=g, dtype=dtype) for g in generator], dim=0).to(device)
    return torch.randn(tuple(shape), generator=generator, dtype=dtype).to(device)


def is_compiled_module(module):
    return False


from . import is_torch_version  # noqa: E402,F401
