class PeftAdapterMixin:
    pass


class UNet2DConditionLoadersMixin:
    pass


class FromSingleFileMixin:
    pass


class LoraLoaderMixin:
    pass


class TextualInversionLoaderMixin:
    def maybe_convert_prompt(self, prompt, tokenizer):
        """identity: no multi-vector textual-inversion tokens are ever loaded here"""
        return prompt


class IPAdapterMixin:
    pass
