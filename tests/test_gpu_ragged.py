"""Length-bucketed (ragged) evaluation batches, SURVEY 8f-1: utterances of different lengths zero-padded into one batch
and masked by length in conv-0's GroupNorm statistics, the positional conv input, attention keys and pooling must give
every utterance what a batch of one gives it (the reference's test loop, R:src/lightning_modules/speaker/
speaker_recognition_module.py:462-500, runs one utterance per step), through all three attention kernels
(persistent T <= 160, single-tile T <= 256, key-tiled T > 256) and both positional-conv paths."""
import pytest
import torch

pytestmark = pytest.mark.gpu
S = 5994


def rows(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()


def _module(pooling, loss, base_params):
    try:
        from test_gpu_round2 import _module as build
    except ImportError:
        from tests.test_gpu_round2 import _module as build
    return build(pooling, loss, base_params).eval()


def _utterances(lengths, seed=3):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in lengths:
        x = torch.randn(n, generator=g)
        out.append((x - x.mean()) / (x.std() + 1e-5))
    return out


@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")])
def test_ragged_batch_equals_one_utterance_at_a_time(base_params, pooling, loss):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    m = _module(pooling, loss, base_params)
    # frames: 49, 73, 99, 149 (persistent kernel), 162 (single-tile kernel), 281 and 405 (key-tiled kernel, chunked pos conv)
    lengths = [16000, 23456, 31999, 48000, 52000, 90000, 129777]
    utts = _utterances(lengths)
    with torch.no_grad():
        single = torch.cat([m.compute_speaker_embedding(u[None].cuda()).reshape(1, -1) for u in utts])
        # one bucket (padding up to 88 %): every masking path is exercised hard
        ragged = m.compute_speaker_embeddings_ragged(utts, max_batch=8, max_pad_fraction=0.9)
        # and the default planning (several buckets, results scattered back in input order)
        planned = m.compute_speaker_embeddings_ragged(utts)
    assert ragged.shape == single.shape == planned.shape
    assert torch.isfinite(ragged).all()
    assert rows(ragged, single) < 5e-4, rows(ragged, single)
    assert rows(planned, single) < 5e-4, rows(planned, single)


def test_ragged_batch_matches_the_oracle(base_params):
    """Anchor: the shortest and a mid-length utterance of a heavily padded batch against the CPU oracle on that utterance
    alone (north_star: 1e-3 on fp32 embeddings)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import w2v2_oracle as O
    m = _module("mean", "ce", base_params)
    lengths = [12000, 40000, 70000]
    utts = _utterances(lengths, seed=8)
    with torch.no_grad():
        got = m.compute_speaker_embeddings_ragged(utts, max_batch=8, max_pad_fraction=0.95)
        torch.set_num_threads(max(8, torch.get_num_threads()))
        for i in (0, 1):
            ref = O.speaker_embedding(utts[i][None], base_params, "mean")
            assert rows(got[i:i + 1], ref) < 1e-3, (i, rows(got[i:i + 1], ref))


def test_ragged_rejects_training_mode(base_params):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    m = _module("mean", "ce", base_params)
    with pytest.raises(RuntimeError):
        m.compute_speaker_embeddings_ragged(_utterances([16000, 17000]))          # gradients enabled
    with torch.no_grad(), pytest.raises(ValueError):
        m.wav2vec.model(torch.zeros(2, 16000, device="cuda"), lengths=[16000, 300])   # shorter than the receptive field


def test_ragged_and_raw_input_with_the_layer_norm_feature_extractor():
    """-lv60 / XLSR variant (LayerNorm conv layers, pre-LN encoder): a zero-padded ragged batch gives every utterance the
    embedding it gets alone (no length-aware statistics are needed in front: every frame is normalised on its own), and raw
    16-bit PCM input equals the standardised float waveform."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dataclasses
    from oracle.params import LARGE_LV60 as O_LV60, make_inputs, make_params
    from w2v2_speaker_b200.engine import LARGE_LV60
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2ModelB200
    arch = dataclasses.replace(LARGE_LV60, layers=3)
    p = make_params(dataclasses.replace(O_LV60, layers=3), seed=6)
    m = Wav2Vec2ModelB200(arch)
    m.load_state_dict(p)
    m = m.cuda().eval()
    lens = [16000, 11283, 24000]
    wavs = [make_inputs(1, n, seed=50 + i)[0][0] for i, n in enumerate(lens)]
    batch = torch.zeros(3, max(lens))
    for i, w in enumerate(wavs):
        batch[i, :len(w)] = w
    with torch.no_grad():
        rag = m(batch.cuda(), lengths=lens).last_hidden_state
        for i, w in enumerate(wavs):
            one = m(w[None, :].cuda()).last_hidden_state[0]
            T = one.shape[0]
            err = ((rag[i, :T] - one).norm() / one.norm()).item()
            assert err < 2e-3, (i, err)
        pcm = (wavs[0] * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
        raw = m(pcm[None, :].cuda()).last_hidden_state
        x = pcm.float()
        ref = m(((x - x.mean()) / (x.std() + 1e-5))[None, :].cuda()).last_hidden_state
        assert ((raw - ref).norm() / ref.norm()).item() < 2e-3
        with pytest.raises(NotImplementedError):
            m(batch.cuda().mul(3000).to(torch.int16), lengths=lens)
