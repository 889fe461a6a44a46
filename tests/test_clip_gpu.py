"""GPU parity of the kernel-backed CLIP text encoder.

Pinned two ways: (a) against tests/golden/token_api.json — encoder outputs produced by the REFERENCE's own
powerpaint/utils/utils.py + transformers CLIPTextModel on the synthetic CLIP (make_token_api_golden.py); (b) against
`transformers.CLIPTextModel` itself (the dependency the reference calls; installed in this image) in fp32 on the GPU,
for the synthetic net and for the full ViT-L/14 text-tower shape with random weights. Tolerance (bf16 storage, 12
layers): rel-L2 <= 2e-2."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def test_text_encoder_matches_reference_golden_and_transformers():
    from synthetic_clip import make_text_encoder, make_tokenizer

    from powerpaint_b200.models import CLIPTextModel
    from powerpaint_b200.utils import TokenizerWrapper, add_tokens

    with open(os.path.join(HERE, "golden", "token_api.json")) as f:
        gold = json.load(f)
    tok = TokenizerWrapper.from_tokenizer(make_tokenizer())
    te = make_text_encoder(len(tok.wrapped), seed=0)
    ours = CLIPTextModel.from_transformers(te)
    # add_tokens on OUR model, like the app does on the transformers one (app.py:100-107)
    add_tokens(tokenizer=tok, text_encoder=ours, placeholder_tokens=["P_ctxt", "P_shape", "P_obj"],
               initialize_tokens=["a", "a", "a"], num_vectors_per_token=10)
    g = torch.Generator().manual_seed(42)
    layer = ours.text_model.embeddings.token_embedding
    with torch.no_grad():
        for name in ["P_ctxt", "P_shape", "P_obj"]:
            layer.trainable_embeddings[name].copy_(torch.randn(10, 32, generator=g))
    ours = ours.to(DEV)
    keys = sorted(k for k in ours.state_dict().keys() if "token_embedding" in k)
    assert keys == gold["state_dict_keys"]
    prompts = [c["prompt"] for c in gold["cases"]]
    ids = tok(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
    hid = ours(ids)[0]
    assert hid.shape == (len(prompts), 77, 32) and torch.isfinite(hid).all()
    for i, case in enumerate(gold["cases"]):
        want = torch.tensor(case["hidden_sum"], device=DEV)
        got = hid[i].sum(-1)
        assert (got - want).abs().max().item() < 0.05 * want.abs().max().item() + 0.05, case["prompt"]
    # one prompt at a time == batched (plans per batch size)
    one = ours(ids[3:4])[0]
    assert _rel(one[0], hid[3]) < 1e-3


@pytest.mark.parametrize("full", [False, True])
def test_text_encoder_matches_transformers(full):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFCLIPTextModel

    from powerpaint_b200.models import CLIPTextModel

    torch.manual_seed(1)
    if full:  # the SD-1.5 text tower: openai/clip-vit-large-patch14 text config
        cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                             num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    else:
        cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                             num_attention_heads=4, max_position_embeddings=77, hidden_act="quick_gelu")
    hf = HFCLIPTextModel(cfg).eval()
    with torch.no_grad():  # default init is tiny (std 0.02): scale up so that every layer matters
        for n, p in hf.named_parameters():
            if p.dim() == 2 and "embedding" not in n:
                p.mul_(3.0)
    ours = CLIPTextModel.from_transformers(hf).to(DEV)
    hf = hf.to(DEV)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg.vocab_size - 1, (4, 77), generator=g)
    ids[:, 0] = cfg.vocab_size - 2
    ids[:, -1] = cfg.vocab_size - 1
    with torch.no_grad():
        ref = hf(ids.to(DEV))
    got = ours(ids)
    assert _rel(got[0], ref.last_hidden_state) < 2e-2, _rel(got[0], ref.last_hidden_state)
    assert got.last_hidden_state.shape == ref.last_hidden_state.shape
