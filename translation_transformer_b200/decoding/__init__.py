from .speculative_decoding import TranslationInferenceGreedySpeculative  # noqa: F401
