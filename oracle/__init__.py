"""CPU oracle for the wav2vec2 speaker hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package ``w2v2_speaker_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it, and there only
as the checker or as the timed CPU baseline -- never as the thing shipped.

Parity pin: the reference repository has no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``oracle/make_golden.py`` imports the reference's own ``Wav2vec2FCModule``
(/root/reference/src/lightning_modules/speaker/wav2vec2_fc.py) under import shims
in the build container, runs it on seeded inputs and commits the results as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures, and against the HuggingFace ``Wav2Vec2Model`` (the
third-party dependency that holds the encoder arithmetic) when it is importable.
The speechbrain attentive-statistics pooling source is absent offline; its
restatement (``w2v2_oracle.attentive_stat_pool``) follows the published
speechbrain 0.5.x algorithm.  It is checked (tests/test_oracle_golden.py) against an
independent port of the same speechbrain layer that IS in the image --
``transformers.models.qwen2_5_omni.modeling_qwen2_5_omni.AttentiveStatisticsPooling``
(ECAPA-TDNN of the Qwen2.5-Omni token2wav speaker encoder): statistics with the
1e-12 clamp, the [x | mean | std] context, tanh, the 1x1 convs, softmax over time and
the [mean | std] output agree exactly.  That port's TDNN block has no BatchNorm, so
the one thing still "parity unpinned" is the position of speechbrain's BatchNorm1d
(conv -> ReLU -> BatchNorm, speechbrain ``TDNNBlock``), which the check neutralises.
"""
