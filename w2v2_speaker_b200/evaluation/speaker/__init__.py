from .cosine_distance import CosineDistanceEvaluator, EmbeddingSample, EvaluationPair  # noqa: F401
