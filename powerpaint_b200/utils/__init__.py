from .utils import EmbeddingLayerWithFixes, TokenizerWrapper, add_tokens

__all__ = ["EmbeddingLayerWithFixes", "TokenizerWrapper", "add_tokens"]
