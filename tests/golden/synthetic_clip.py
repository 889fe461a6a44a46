"""A tiny deterministic CLIP tokenizer + text encoder for offline tests (no checkpoint is reachable).
Shared by the golden-vector generator (which runs the REFERENCE's powerpaint/utils/utils.py) and by the
tests (which run powerpaint_b200/utils/utils.py) so both see byte-identical vocab and weights."""
import torch


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


WORDS = ["a", "cat", "photo", "of", "the", "dog", "empty", "scene", "blur", "wall", "sky", "chair", "on"]


def make_tokenizer():
    from transformers import CLIPTokenizer

    chars = list(_bytes_to_unicode().values())
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = []
    for w in WORDS:
        if len(w) == 1:
            continue
        syms = list(w[:-1]) + [w[-1] + "</w>"]
        while len(syms) > 1:
            pair = (syms[0], syms[1])
            if pair not in merges:
                merges.append(pair)
            vocab.setdefault(syms[0] + syms[1], len(vocab))
            syms = [syms[0] + syms[1]] + syms[2:]
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    return CLIPTokenizer(vocab=vocab, merges=[(a, b) for a, b in merges], model_max_length=77)


def make_text_encoder(vocab_size: int, hidden: int = 32, seed: int = 0):
    from transformers import CLIPTextConfig, CLIPTextModel

    torch.manual_seed(seed)
    cfg = CLIPTextConfig(vocab_size=vocab_size, hidden_size=hidden, intermediate_size=64, num_hidden_layers=2,
                         num_attention_heads=4, max_position_embeddings=77, projection_dim=hidden)
    return CLIPTextModel(cfg).eval()
