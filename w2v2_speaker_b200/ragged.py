"""Length-bucketed batching of full-length test utterances.

The reference embeds test utterances one at a time (``test_step`` insists on a batch of one,
R:src/lightning_modules/speaker/speaker_recognition_module.py:462-470; ``predict.py`` loops over files,
R:src/predict.py:132-170): a latency-bound use of the GPU.  Here utterances of similar length are zero-padded into
one batch and carried with their lengths; the kernels mask by length (conv-0 GroupNorm statistics, zeros behind the
end for the positional conv, attention keys, pooling), so every utterance gets the numbers a batch of one gives it.

Host-side planning only: which utterances share a batch."""
from __future__ import annotations

from typing import List, Sequence


def plan_buckets(lengths: Sequence[int], max_batch: int = 32, max_pad_fraction: float = 0.15,
                 max_batch_samples: int = 64 * 48000) -> List[List[int]]:
    """Greedy buckets over the utterances sorted by length (longest first).  A bucket is closed when it holds
    `max_batch` utterances, when the next (shorter) utterance would be padded by more than `max_pad_fraction` of the
    bucket's longest, or when the padded batch would exceed `max_batch_samples` samples (memory of a forward).
    -> list of buckets, each a list of indices into `lengths`; every index appears exactly once."""
    if max_batch < 1 or not 0.0 <= max_pad_fraction < 1.0:
        raise ValueError("max_batch >= 1 and 0 <= max_pad_fraction < 1 are required")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    buckets: List[List[int]] = []
    cur: List[int] = []
    longest = 0
    for i in order:
        n = int(lengths[i])
        if n < 1:
            raise ValueError(f"utterance {i} is empty")
        if cur and (len(cur) >= max_batch or n < (1.0 - max_pad_fraction) * longest
                    or (len(cur) + 1) * longest > max_batch_samples):
            buckets.append(cur)
            cur = []
        if not cur:
            longest = n
        cur.append(i)
    if cur:
        buckets.append(cur)
    return buckets
