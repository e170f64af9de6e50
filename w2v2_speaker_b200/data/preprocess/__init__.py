from .input_normalisation import InputNormalizer2D  # noqa: F401
