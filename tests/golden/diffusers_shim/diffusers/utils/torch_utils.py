def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **freeu_kwargs):
    raise NotImplementedError("FreeU is outside the hot path (the reference never sets s1/s2/b1/b2)")
