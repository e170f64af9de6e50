"""GPU parity of the SPLIT call path -- ``model.feature_extractor`` / ``model.feature_projection`` / ``model.encoder``
called one by one on a caller-built sequence (the reference's CLS-token wrapper, R:src/models/wav2vec2.py:128-140, and
its paired-input model, R:src/lightning_modules/speaker/wav2vec2_paired_input.py:162-207) -- in training, against
autograd of the CPU oracle; and of the Function boundary contract: gradients crossing it are plain unscaled fp32, so a
torch head and a torch loss around the encoder give the right parameter gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ZERO_REG = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                mask_time_prob=0.0, mask_feature_prob=0.0)


def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _compare_encoder_grads(got, ref, train_cnn, tol=1e-2, kbias_tol=1e-2):
    """got: name -> cuda Parameter; ref: name -> CPU leaf with .grad.  Same exemptions as test_gpu_training."""
    worst = (0.0, None)
    for k, v in ref.items():
        if k.startswith("feature_extractor") and not train_cnn:
            assert got[k].grad is None, k
            continue
        if k == "masked_spec_embed":                  # never used on these paths
            assert got[k].grad is None or got[k].grad.abs().max().item() == 0.0
            continue
        assert got[k].grad is not None, k
        g, r = got[k].grad.detach().cpu().double(), v.grad.double()
        if k.endswith("k_proj.bias"):                 # exactly 0 in exact arithmetic (softmax shift invariance)
            scale = ref[k.replace("k_proj", "q_proj")].grad.double().norm()
            assert g.norm() < kbias_tol * scale and r.norm() < 1e-2 * scale, k
            continue
        rel = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
        worst = max(worst, (rel, k))
        assert rel < tol, (k, rel)
    return worst


def _oracle_sequence_forward(wavs, p, tokens):
    """Oracle restatement of the split path: per-utterance CNN + projection, constant tokens in between, encoder."""
    from oracle import w2v2_oracle as O
    return O.split_path_forward(wavs, p, tokens)                       # [B, T', H]


@pytest.mark.parametrize("cls_token,train_cnn", [(False, False), (True, False), (True, True)])
def test_wrapper_trains_through_a_torch_head_and_loss(base_params, cls_token, train_cnn):
    """Wav2Vec2WrapperModule (fused path and CLS-token split path) -> torch readout -> torch MSE: every parameter
    gradient against the oracle.  The Functions' internal loss scale must not leak through their boundaries."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import BASE, make_inputs
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule
    wav, _ = make_inputs(3, 11283, seed=77)
    gen = torch.Generator().manual_seed(5)
    Wr = torch.randn(4, 768, generator=gen) * 0.05
    target = torch.randn(3, 4, generator=gen)
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False, reg_cfg=Wav2Vec2RegularisationConfig(**ZERO_REG),
                              insert_clc_token=cls_token)
    w.model.load_state_dict(base_params)
    w = w.cuda().train()
    w.model.feature_extractor.requires_grad_(train_cnn)
    readout = Wr.clone().cuda().requires_grad_(True)
    out = w(wav.cuda())                                               # [B, H, T(+1)]
    feat = out[:, :, 0] if cls_token else out.mean(dim=2)
    loss = F.mse_loss(feat @ readout.t(), target.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(train_cnn or not k.startswith("feature_extractor")) for k, v in base_params.items()}
    rr = Wr.clone().requires_grad_(True)
    if cls_token:
        ref_out = _oracle_sequence_forward([wav], p, [1.0])
        ref_feat = ref_out[:, 0, :]
    else:
        ref_out = O.wav2vec2_forward(wav, p, BASE)
        ref_feat = ref_out.mean(dim=1)
    ref_loss = F.mse_loss(ref_feat @ rr.t(), target)
    ref_loss.backward()
    assert tuple(out.shape) == (3, 768, ref_out.shape[1])
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 2e-3
    rel = ((readout.grad.cpu().double() - rr.grad.double()).norm() / rr.grad.double().norm()).item()
    assert rel < 1e-2, rel
    # CNN gradients on the CLS path are second-order small (they only flow through the attention of one token): 2e-2
    worst = _compare_encoder_grads(dict(w.model.named_parameters()), p, train_cnn, tol=2e-2 if train_cnn else 1e-2)
    print("worst parameter-gradient error", worst)


def test_paired_input_model_matches_oracle(base_params):
    """Wav2vec2PairedSpeakerModule: scores in eval mode, then loss and every gradient (encoder, torch Linear head)
    of a training step with torch's BCE, against the oracle's composition of the same steps."""
    _need_cuda()
    from oracle.params import make_inputs
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    B = 3
    wav_a, _ = make_inputs(B, 16000, seed=31)                          # 49 frames
    wav_b, _ = make_inputs(B, 11283, seed=32)                          # 35 frames
    labels = torch.tensor([1, 0, 1])
    cfg = Wav2vec2PairedSpeakerModuleConfig(**ZERO_REG)
    torch.manual_seed(3)
    m = Wav2vec2PairedSpeakerModule(cfg, BinaryCrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params)
    lin_w, lin_b = m.linear.weight.detach().clone(), m.linear.bias.detach().clone()
    m = m.cuda()

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    lw, lb = lin_w.clone().requires_grad_(True), lin_b.clone().requires_grad_(True)
    ref_tokens = _oracle_sequence_forward([wav_a, wav_b], p, [1.0, -1.0, -1.0])
    assert ref_tokens.shape[1] == 1 + 49 + 1 + 35 + 1
    ref_scores = F.linear(ref_tokens[:, 0, :], lw, lb)
    ref_loss = F.binary_cross_entropy_with_logits(ref_scores.squeeze(), labels.float())
    ref_loss.backward()

    m.eval()
    with torch.no_grad():
        scores = m(wav_a.cuda(), wav_b.cuda())
    assert scores.shape == (B, 1)
    # a readout of ONE token after 12 layers on fp16 operands (the 1e-3 bar of the path is on pooled embeddings)
    assert (scores.cpu() - ref_scores.detach()).abs().max().item() < 5e-3 * max(1.0, ref_scores.abs().max().item())

    m.train()
    m.on_train_start()                                                # freezes the CNN (cfg default)
    scores = m(wav_a.cuda(), wav_b.cuda())
    loss, prediction = m.loss_fn(scores, labels.cuda())
    loss.backward()
    m.on_after_backward()
    torch.cuda.synchronize()
    assert prediction.shape == (B,) and m.steps == 1
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 5e-3
    for g, r, k in ((m.linear.weight.grad, lw.grad, "linear.weight"), (m.linear.bias.grad, lb.grad, "linear.bias")):
        rel = ((g.cpu().double() - r.double()).norm() / r.double().norm()).item()
        assert rel < 1e-2, (k, rel)
    worst = _compare_encoder_grads(dict(m.wav2vec.model.named_parameters()), p, False)
    print("worst parameter-gradient error", worst)


def test_paired_model_trains_with_default_regularisation(base_params):
    """The split path with the reference's default regularisation (dropouts, LayerDrop; no SpecAugment on this path):
    a few steps of a stock torch optimizer reduce the (eval-mode) BCE loss; gradients are finite, reach the trained
    tensors and leave the frozen CNN alone."""
    _need_cuda()
    from oracle.params import make_inputs
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    B = 4
    wav_a, _ = make_inputs(B, 16000, seed=41)
    wav_b, _ = make_inputs(B, 16000, seed=42)
    wav_a, wav_b = wav_a.cuda(), wav_b.cuda()
    labels = torch.tensor([1, 0, 0, 1]).cuda()
    torch.manual_seed(11)
    m = Wav2vec2PairedSpeakerModule(Wav2vec2PairedSpeakerModuleConfig(layerdrop=0.25), BinaryCrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params)
    m = m.cuda()

    def eval_loss():
        m.eval()
        with torch.no_grad():
            return m.loss_fn(m(wav_a, wav_b), labels)[0].item()

    before = eval_loss()
    m.train()
    m.on_train_start()
    opt = torch.optim.Adam([q for q in m.parameters() if q.requires_grad], lr=1e-4)
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss, _ = m.loss_fn(m(wav_a, wav_b), labels)
        loss.backward()
        for n, q in m.named_parameters():
            if q.grad is not None:
                assert torch.isfinite(q.grad).all(), n
        opt.step()
    named = dict(m.wav2vec.model.named_parameters())
    assert named["feature_projection.projection.weight"].grad is not None
    assert named["encoder.layers.0.attention.q_proj.weight"].grad is not None
    assert named["feature_extractor.conv_layers.1.conv.weight"].grad is None
    after = eval_loss()
    assert after < before, (before, after)
