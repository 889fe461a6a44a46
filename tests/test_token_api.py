"""Task-prompt token API pinned against the REFERENCE's own powerpaint/utils/utils.py:
tests/golden/token_api.json was produced by tests/golden/make_token_api_golden.py importing the
reference file (with an mmengine stub) on the synthetic CLIP of tests/golden/synthetic_clip.py."""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from synthetic_clip import make_text_encoder, make_tokenizer  # noqa: E402

from powerpaint_b200.utils import EmbeddingLayerWithFixes, TokenizerWrapper, add_tokens  # noqa: E402


@pytest.fixture(scope="module")
def setup():
    with open(os.path.join(HERE, "golden", "token_api.json")) as f:
        gold = json.load(f)
    tok = TokenizerWrapper.from_tokenizer(make_tokenizer())
    te = make_text_encoder(len(tok.wrapped), seed=0)
    add_tokens(tokenizer=tok, text_encoder=te, placeholder_tokens=["P_ctxt", "P_shape", "P_obj"],
               initialize_tokens=["a", "a", "a"], num_vectors_per_token=10)
    g = torch.Generator().manual_seed(42)
    layer = te.text_model.embeddings.token_embedding
    with torch.no_grad():
        for name in ["P_ctxt", "P_shape", "P_obj"]:
            layer.trainable_embeddings[name].copy_(torch.randn(10, 32, generator=g))
    return gold, tok, te, layer


def test_vocab_and_state_dict_names(setup):
    gold, tok, te, layer = setup
    assert len(tok.wrapped) == gold["vocab_after"]
    assert tok.token_map == gold["token_map"]
    keys = sorted(k for k in te.state_dict().keys() if "token_embedding" in k)
    assert keys == gold["state_dict_keys"]
    assert "text_model.embeddings.token_embedding.wrapped.weight" in keys
    assert "text_model.embeddings.token_embedding.trainable_embeddings.P_obj" in keys
    for n in ["P_ctxt", "P_shape", "P_obj"]:
        assert tok.get_token_info(n) == gold["token_info"][n]
    assert isinstance(layer, EmbeddingLayerWithFixes) and layer.num_embeddings == gold["base_vocab"]


def test_ids_embeddings_and_encoder_output_match_reference(setup):
    gold, tok, te, layer = setup
    for case in gold["cases"]:
        p = case["prompt"]
        assert tok.replace_placeholder_tokens_in_text(p) == case["replaced_text"]
        ids = tok(p, padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        assert ids[0].tolist() == case["input_ids"], p
        with torch.no_grad():
            emb = layer(ids)
            hid = te(ids)[0]
        assert torch.allclose(emb[0].sum(-1), torch.tensor(case["token_embedding_sum"]), atol=1e-5), p
        first = emb[0, :, :8].flatten()[: 8 * 20]
        assert torch.allclose(first, torch.tensor(case["token_embedding_first8"]), atol=1e-6), p
        assert torch.allclose(hid[0].sum(-1), torch.tensor(case["hidden_sum"]), atol=1e-4), p
        assert tok.decode(ids[0].tolist(), skip_special_tokens=True) == case["decode"]


def test_placeholder_rules_and_errors(setup):
    _, tok, te, layer = setup
    with pytest.raises(ValueError):
        tok.add_placeholder_token("P_obj_extra", num_vec_per_token=2)  # confusable with P_obj
    with pytest.raises(AssertionError):
        tok.try_adding_tokens("P_obj_0")  # already in the vocabulary
    # a run truncated by max_length is an invalid id sequence -> AssertionError like the reference
    info = tok.get_token_info("P_obj")
    ids = torch.tensor([[info["start"], info["start"] + 1, 5, 6]])
    with pytest.raises(AssertionError):
        layer(ids)
    # ids without any placeholder take the plain embedding path
    plain = torch.tensor([[1, 2, 3]])
    assert torch.equal(layer(plain), layer.wrapped(plain))
    # 1-D input is batched like the reference
    assert layer(torch.tensor([1, 2, 3])).shape == (1, 3, 32)
    # external embeddings passed at call time
    ext = {"name": "tmp", "start": 1000, "end": 1002, "embedding": torch.ones(2, 32)}
    out = layer(torch.tensor([[4, 1000, 1001, 7]]), external_embeddings=[ext])
    assert torch.equal(out[0, 1:3], torch.ones(2, 32))
    with pytest.raises(AssertionError):
        layer.add_embeddings([{"name": "P_obj", "start": 5000, "end": 5001, "embedding": torch.zeros(1, 32)}])


def test_tradoff_blend_property():
    """tradoff = 1 => blended embeddings == promptA embeddings (pipeline_PowerPaint.py:423)"""
    a, b = torch.randn(2, 77, 8), torch.randn(2, 77, 8)
    assert torch.equal(a * 1.0 + (1 - 1.0) * b, a)


def EmbeddingRuns(row, start):
    return (row == start).nonzero().flatten().tolist()


def test_gpu_text_encoder_gather_plan_equals_embedding_layer(setup):
    """host logic of the kernel-backed CLIPTextModel: the task-prompt splice resolved to ONE gather index per
    position reproduces EmbeddingLayerWithFixes.forward (pinned above against the reference) exactly — including
    the adjacent-run scan quirk and the asserts on malformed runs"""
    from powerpaint_b200.models.clip_text import CLIPTextModel

    gold, tok, te, layer = setup
    m = CLIPTextModel.from_transformers(te)
    assert m.text_model.embeddings.token_embedding is layer
    V = m.config.vocab_size
    for case in gold["cases"]:
        ids = tok(case["prompt"], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        idx, ext = m._gather_plan(ids)
        table = torch.cat([layer.wrapped.weight.detach(), ext], 0)
        got = table[idx.long()]
        with torch.no_grad():
            want = layer(ids)
        assert torch.equal(got, want), case["prompt"]
        n_runs = sum(len(EmbeddingRuns(ids[0], tok.get_token_info(n)["start"])) for n in ("P_ctxt", "P_shape", "P_obj"))
        assert int((idx >= V).sum()) <= 10 * n_runs  # <=: the quirk leaves an adjacent second run un-replaced
    bad = torch.tensor([[tok.get_token_info("P_obj")["start"], 5, 6] + [0] * 74])
    with pytest.raises(AssertionError):
        m._gather_plan(bad)
    with pytest.raises(RuntimeError):
        m(ids)  # CPU parameters: there is no CPU path
