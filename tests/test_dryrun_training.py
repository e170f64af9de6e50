"""CPU tier: the host side of the training / evaluation path with the kernels stubbed out (tests/dryrun.py).
What is checked here is what a GPU is not needed for: that every module variant drives the launch schedules end to end
(forward, loss, ``loss.backward()``), that each C entry point is called with an argument list its ctypes signature
accepts, how many per-layer schedule calls a step makes, and which parameters end up with a gradient."""
import collections

import pytest
import torch

from dryrun import dry_library

S = 64


def _fc_module(pooling, loss, **cfg_kw):
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling, mask_time_prob=0.0, **cfg_kw)
    ctor = CrossEntropyLoss if loss == "ce" else (
        lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30))
    return Wav2vec2FCModule(cfg, S, ctor)


@pytest.mark.parametrize("pooling,loss,freeze_cnn", [("mean", "ce", True), ("mean+std", "aam", True), ("attentive", "aam", True),
                                                     ("first+cls", "ce", True), ("mean", "ce", False), ("first+cls", "ce", False),
                                                     ("max", "ce", True), ("quantile", "ce", True), ("first", "ce", True),
                                                     ("last", "ce", True)])
def test_training_step_control_flow(pooling, loss, freeze_cnn):
    with dry_library() as lib:
        m = _fc_module(pooling, loss, layerdrop=0.0, completely_freeze_feature_extractor=freeze_cnn).train()
        m.wav2vec.model.feature_extractor.requires_grad_(not freeze_cnn)
        emb, pred = m(torch.randn(3, 1, 16000))
        out, prob = m.loss_fn(pred, torch.tensor([1, 2, 3]))
        width = {"mean+std": 1536, "attentive": 1536, "quantile": 5 * 768}.get(pooling, 768)
        assert emb.shape == (3, width) and prob.shape == (3, S)
        out.backward()
        calls = collections.Counter(lib.calls)
    assert calls["w2v2_encoder_layer_fwd"] == 12 and calls["w2v2_encoder_layer_bwd"] == 12
    # the loss scale enters the encoder backward exactly once per step, out of place (autograd owns the incoming gradient)
    # (w2v2_grad_entry_scale: the same copy with the device-chosen normalisation that keeps an outer GradScaler harmless)
    assert calls["w2v2_grad_entry_scale"] >= 1 and calls["w2v2_scale_f32_dev"] >= 1
    for n, q in m.wav2vec.model.named_parameters():
        if n == "masked_spec_embed":
            continue
        frozen = n.startswith("feature_extractor.") and freeze_cnn
        assert (q.grad is None) == frozen, n
    for n, q in m.named_parameters():
        if not n.startswith("wav2vec.") and q.requires_grad:
            assert q.grad is not None, n


def test_layerdrop_of_every_layer_is_the_identity_stack():
    with dry_library() as lib:
        m = _fc_module("mean", "ce", layerdrop=1.0).train()
        m.wav2vec.model.feature_extractor.requires_grad_(False)
        emb, pred = m(torch.randn(2, 1, 8000))
        m.loss_fn(pred, torch.tensor([0, 1]))[0].backward()
        calls = collections.Counter(lib.calls)
    assert calls["w2v2_encoder_layer_fwd"] == 0 and calls["w2v2_encoder_layer_bwd"] == 0
    named = dict(m.wav2vec.model.named_parameters())
    assert named["feature_projection.projection.weight"].grad is not None


def test_paired_input_step_control_flow():
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    with dry_library() as lib:
        m = Wav2vec2PairedSpeakerModule(Wav2vec2PairedSpeakerModuleConfig(layerdrop=0.0), BinaryCrossEntropyLoss).train()
        m.on_train_start()
        scores = m(torch.randn(2, 16000), torch.randn(2, 12000))
        assert scores.shape == (2, 1)
        loss, prediction = m.loss_fn(scores, torch.tensor([1, 0]))
        loss.backward()
        m.on_after_backward()
        calls = collections.Counter(lib.calls)
    assert prediction.shape == (2,) and m.steps == 1
    assert calls["w2v2_encoder_layer_fwd"] == 12 and calls["w2v2_encoder_layer_bwd"] == 12
    named = dict(m.wav2vec.model.named_parameters())
    assert named["feature_extractor.conv_layers.3.conv.weight"].grad is None         # completely_freeze_feature_extractor
    assert named["feature_projection.projection.weight"].grad is not None           # two calls, gradients summed by autograd
    assert named["encoder.layers.11.final_layer_norm.bias"].grad is not None
    assert m.linear.weight.grad is not None


def test_paired_input_freeze_protocol():
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    cfg = Wav2vec2PairedSpeakerModuleConfig(wav2vec_initially_frozen=True, num_frozen_steps=2,
                                            completely_freeze_feature_projector=True)
    m = Wav2vec2PairedSpeakerModule(cfg, BinaryCrossEntropyLoss)
    m.on_train_start()
    assert not any(q.requires_grad for q in m.wav2vec.parameters()) and m.linear.weight.requires_grad
    m.on_after_backward()
    assert not any(q.requires_grad for q in m.wav2vec.parameters())
    m.on_after_backward()                                                           # step 2: unfreeze, keep the permanent ones
    named = dict(m.wav2vec.model.named_parameters())
    assert named["encoder.layers.0.attention.q_proj.weight"].requires_grad
    assert not named["feature_extractor.conv_layers.0.conv.weight"].requires_grad
    assert not named["feature_projection.projection.weight"].requires_grad


@pytest.mark.parametrize("seconds", [1, 6])
def test_evaluation_forward_control_flow(seconds):
    """Eval forward of a short and of a 6 s utterance (299 frames > 256: key-tiled attention, chunked positional conv)."""
    with dry_library() as lib:
        m = _fc_module("mean+std", "aam").eval()
        with torch.no_grad():
            emb, pred = m(torch.randn(2, 1, 16000 * seconds))
        calls = collections.Counter(lib.calls)
    assert emb.shape == (2, 1536)
    assert calls["w2v2_encoder_layer_fwd"] == 12 and calls["w2v2_encoder_layer_bwd"] == 0


def test_bce_loss_matches_torch():
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    logits = torch.tensor([[0.3], [-1.2], [2.0]], requires_grad=True)
    labels = torch.tensor([1, 0, 0])
    loss, prediction = BinaryCrossEntropyLoss()(logits, labels)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(logits.detach().squeeze(), labels.float())
    assert torch.allclose(loss, ref) and torch.allclose(prediction, torch.sigmoid(logits.detach().squeeze()))
    assert not prediction.requires_grad
    loss.backward()
    assert logits.grad is not None


def test_derived_weights_follow_parameter_versions():
    """The fp16 operand copies are rebuilt when a parameter's autograd version or storage changes, updated in place by
    refresh(), and (documented limit) not touched by edits made through .data."""
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2WrapperModule
    with dry_library() as lib:
        w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False).eval()
        model = w.model
        eng = model._engine()
        assert model._engine() is eng                                   # nothing changed: cached
        q = dict(model.named_parameters())["encoder.layers.3.attention.q_proj.weight"]
        with torch.no_grad():
            q.data.mul_(1.0)                                            # invisible to the version counter
        assert model._engine() is eng
        lib.calls.clear()
        model.refresh()                                                 # re-derives the copies with one batched launch
        assert model._engine() is eng and lib.calls.count("w2v2_prepare_weights") == 1
        with torch.no_grad():
            q.mul_(1.0)                                                 # what torch.optim does: bumps the version
        eng2 = model._engine()
        assert eng2 is not eng
        model.load_state_dict(model.state_dict())                       # copies in place: versions change again
        assert model._engine() is not eng2


@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("attentive", "aam")])
def test_fc_module_freeze_protocol(pooling, loss):
    """SURVEY 8 row a13 (R:src/lightning_modules/speaker/wav2vec2_fc.py:339-361): the encoder starts frozen, the heads
    train on their own (their kernels must still produce gradients with a gradient-less embedding), the encoder is
    released after num_frozen_steps while the CNN stays frozen for good."""
    with dry_library() as lib:
        m = _fc_module(pooling, loss, layerdrop=0.0, wav2vec_initially_frozen=True, num_frozen_steps=2).train()
        m.on_train_start()
        assert not any(q.requires_grad for q in m.wav2vec.parameters()) and not m.wav2vec.training
        heads = [(n, q) for n, q in m.named_parameters() if not n.startswith("wav2vec.")]
        assert heads and all(q.requires_grad for _, q in heads)
        for step in (1, 2, 3):
            m.zero_grad(set_to_none=True)
            lib.calls.clear()
            emb, pred = m(torch.randn(3, 1, 16000))
            out, _ = m.loss_fn(pred, torch.tensor([1, 2, 3]))
            out.backward()
            calls = collections.Counter(lib.calls)
            assert all(q.grad is not None for _, q in heads), step
            named = dict(m.wav2vec.model.named_parameters())
            if step <= 2:                                   # frozen phase: forward schedules only, no encoder backward
                assert calls["w2v2_encoder_layer_bwd"] == 0 and named["encoder.layers.0.attention.q_proj.weight"].grad is None
            else:                                           # released after the second on_after_backward
                assert calls["w2v2_encoder_layer_bwd"] == 12 and named["encoder.layers.0.attention.q_proj.weight"].grad is not None
                assert named["feature_extractor.conv_layers.0.conv.weight"].grad is None
            m.on_after_backward()
        assert m.steps == 3 and not m._is_wav2vec_frozen and m.wav2vec.training
    assert m.generate_example_input(True, 4).shape == (4, 16000) and m.generate_example_input(False).shape == (16000,)


def test_evaluation_handles_any_utterance_length():
    """Host side of the any-length evaluation forward (frame arithmetic HF:1012-1018, single-slab / chunked positional
    conv, single-tile / key-tiled attention): one frame up to 70 s; shorter than the receptive field raises."""
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2WrapperModule
    with dry_library():
        w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False).eval()
        with torch.no_grad():
            for n, frames in ((400, 1), (719, 1), (720, 2), (16000, 49), (82000, 256), (82320, 257), (160400, 501),
                              (1120000, 3499)):
                assert w(torch.randn(1, n)).shape == (1, 768, frames), n
            with pytest.raises(ValueError):
                w(torch.randn(2, 399))


def test_frozen_encoder_in_train_mode_runs_the_stochastic_forward():
    """`wav2vec_initially_frozen` + Lightning's `.train()` after a validation loop leaves a frozen encoder in
    train mode; HF then runs a dropout-active forward without gradients -- so must the mirror (it used to raise)."""
    with dry_library() as lib:
        m = _fc_module("mean", "ce", layerdrop=0.0, wav2vec_initially_frozen=True, num_frozen_steps=100).train()
        m.on_train_start()
        m.train()                                            # what Lightning does after every validation loop
        assert m.wav2vec.training and not any(q.requires_grad for q in m.wav2vec.parameters())
        lib.calls.clear()
        emb, pred = m(torch.randn(3, 1, 16000))
        calls = collections.Counter(lib.calls)
        assert calls["w2v2_encoder_layer_fwd"] == 12 and calls["w2v2_dropout"] >= 1      # the regularised forward ran
        assert not emb.requires_grad or pred.requires_grad   # heads still differentiable
        out, _ = m.loss_fn(pred, torch.tensor([1, 2, 3]))
        out.backward()
        assert collections.Counter(lib.calls)["w2v2_encoder_layer_bwd"] == 0


def test_flat_trainer_follows_the_freeze_protocol():
    """a FlatAdamTrainer built while the encoder is frozen must pick the encoder up once
    on_after_backward() releases it (flat buffers + gradient sink rebuilt, Adam state of the heads carried over)."""
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    with dry_library() as lib:
        m = _fc_module("mean", "ce", layerdrop=0.0, wav2vec_initially_frozen=True, num_frozen_steps=1).train()
        m.on_train_start()
        tr = FlatAdamTrainer(m, lr=1e-3)
        n_heads = tr.flat_p.numel()
        assert tr.n0 == 0 and tr.model._grad_sink is None and not tr._encoder_joins
        tr.m.fill_(0.5)                                      # stand-in for accumulated Adam state of the heads
        tr.step(torch.randn(2, 1, 16000), torch.tensor([1, 2]))
        m.on_after_backward()                                # releases the encoder (CNN stays frozen)
        lib.calls.clear()
        tr.step(torch.randn(2, 1, 16000), torch.tensor([1, 2]))
        assert tr.n0 > 90_000_000 and tr.model._grad_sink is not None and tr._encoder_joins
        assert tr.flat_p.numel() == tr.n0 + n_heads
        assert collections.Counter(lib.calls)["w2v2_encoder_layer_bwd"] == 12
        assert torch.all(tr.m[tr.n0:] == 0.5) and torch.all(tr.m[:tr.n0] == 0)          # heads kept, encoder fresh
        q = dict(m.wav2vec.model.named_parameters())["encoder.layers.0.attention.q_proj.weight"]
        assert q.data_ptr() >= tr.flat_p.data_ptr() and q.data_ptr() < tr.flat_p.data_ptr() + 4 * tr.flat_p.numel()


def test_hidden_states_under_grad_are_returned_detached():
    """`model(x, output_hidden_states=True)` used to raise with gradients enabled; it now returns the
    13 per-layer outputs the training forward keeps anyway (detached: gradients flow through last_hidden_state)."""
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule
    with dry_library():
        w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False, reg_cfg=Wav2Vec2RegularisationConfig(layerdrop=0.5)).train()
        w.model.feature_extractor.requires_grad_(False)
        out = w.model(torch.randn(2, 16000), output_hidden_states=True)
        assert out.last_hidden_state.requires_grad and out.last_hidden_state.shape == (2, 49, 768)
        assert len(out.hidden_states) == 13 and all(h.shape == (2, 49, 768) and not h.requires_grad for h in out.hidden_states)
        out.last_hidden_state.sum().backward()


def test_stable_layer_norm_variant_evaluation_control_flow():
    """-lv60 / XLSR architecture (2 layers of it): the evaluation forward composes, per conv layer, tap-GEMM -> LayerNorm
    (+ conv bias) -> GELU, and per encoder layer LN -> QKV -> attention -> out_proj -> add -> LN -> FFN1 -> FFN2 -> add;
    hidden_states has layers + 1 entries (the input of every layer and the normalised output, as HF returns them)."""
    import dataclasses
    from w2v2_speaker_b200.engine import LARGE_LV60
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2ModelB200
    with dry_library() as lib:
        m = Wav2Vec2ModelB200(dataclasses.replace(LARGE_LV60, layers=2)).eval()
        with torch.no_grad():
            out = m(torch.zeros(2, 12000), output_hidden_states=True)
        assert tuple(out.last_hidden_state.shape) == (2, 37, 1024) and len(out.hidden_states) == 3
        c = collections.Counter(lib.calls)
        assert c["w2v2_gelu_fwd"] == 7 and c["w2v2_attention_ex"] == 2 and c["w2v2_add2_cast"] == 1 + 2 * 2
        assert c["w2v2_layernorm_ex"] == 7 + 1 + 2 * 2 + 1               # conv layers, projection, two per layer, final
        assert c["w2v2_gemm_f16"] == 7 + 1 + 2 * 4
        assert "w2v2_encoder_layer_fwd" not in c                         # (the post-LN schedule is not used)


def test_stable_layer_norm_variant_training_control_flow():
    """The pre-LN training path (training_stable.py) on 2 layers of the -lv60 architecture, default regularisation, CNN
    frozen: forward + backward drive the kernels launch by launch (no native layer schedule); every parameter behind the
    CNN gets a gradient, the CNN gets none."""
    import dataclasses
    from w2v2_speaker_b200.engine import LARGE_LV60
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2ModelB200, Wav2Vec2RegularisationConfig
    with dry_library() as lib:
        m = Wav2Vec2ModelB200(dataclasses.replace(LARGE_LV60, layers=2), Wav2Vec2RegularisationConfig(layerdrop=0.0)).train()
        m.feature_extractor.requires_grad_(False)
        out = m(torch.zeros(2, 16000)).last_hidden_state
        assert tuple(out.shape) == (2, 49, 1024)
        out.sum().backward()
        c = collections.Counter(lib.calls)
        # (activation dropout, off by default, takes the two-pass GELU backward: same argument-list check)
        m2 = Wav2Vec2ModelB200(dataclasses.replace(LARGE_LV60, layers=1),
                               Wav2Vec2RegularisationConfig(layerdrop=0.0, activation_dropout=0.1)).train()
        m2.feature_extractor.requires_grad_(False)
        n0 = lib.calls.count("w2v2_gelu_bwd_colsum")
        m2(torch.zeros(2, 16000)).last_hidden_state.sum().backward()
        assert lib.calls.count("w2v2_gelu_bwd_colsum") == n0 + 2      # positional conv + the layer's FFN
        assert "w2v2_encoder_layer_fwd" not in c and "w2v2_encoder_layer_bwd" not in c
        assert c["w2v2_attention_ex"] == 2 and c["w2v2_attention_bwd_ex2"] == 2
        assert c["w2v2_gemm_wgrad_f16"] == 2 * 4 + 1                    # four per layer + the feature projection
        assert c["w2v2_layernorm_bwd_ex"] == 2 * 2 + 1 + 1              # two per layer, final LayerNorm, projection
        for n, q in m.named_parameters():
            if n.startswith("feature_extractor"):
                assert q.grad is None, n
            elif n != "masked_spec_embed":
                assert q.grad is not None, n


def test_feature_axis_specaugment_control_flow():
    """mask_feature_prob > 0 (HF:1312-1322): the feature mask is drawn per step on the [B, hidden] grid and applied once in
    the forward and once to the gradient in the backward (fused call path only, like HF)."""
    with dry_library() as lib:
        m = _fc_module("mean", "ce", layerdrop=0.0, mask_feature_prob=0.2).train()
        m.on_train_start()
        emb, pred = m(torch.randn(2, 1, 16000))
        out, _ = m.loss_fn(pred, torch.tensor([1, 2]))
        out.backward()
        assert lib.calls.count("w2v2_feature_mask") == 2
