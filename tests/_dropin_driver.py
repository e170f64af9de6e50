"""Run in a subprocess by tests/test_reference_dropin.py (the reference's modules must not leak into the test process).

The REFERENCE'S OWN ``Wav2vec2FCModule`` (R:src/lightning_modules/speaker/wav2vec2_fc.py:101-236, 339-438) is imported from
/root/reference, unmodified, after ``w2v2_speaker_b200.integration.install()`` has put this package's mirrors under the
reference's module names.  With a CUDA device the kernels run and the numbers are compared with the committed fixtures (which
the reference produced with its own HF / torch path); without one the C entry points are replaced by the argument-checking
stubs of tests/dryrun.py, so this checks construction, call protocol, autograd plumbing and the freeze protocol.
Prints one JSON object."""
import collections
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np          # noqa: E402
import torch                # noqa: E402


def main():
    from oracle.make_golden import install_shims
    install_shims()                                   # pytorch_lightning / omegaconf / ... stand-ins, /root/reference on sys.path
    import w2v2_speaker_b200.integration as b200
    b200.install()
    from src.lightning_modules.speaker.wav2vec2_fc import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    import src.models.wav2vec2 as ref_models
    from src.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    assert ref_models.__name__ == "w2v2_speaker_b200.models.wav2vec2"
    assert Wav2vec2FCModule.__module__ == "src.lightning_modules.speaker.wav2vec2_fc"        # the reference's class
    assert CrossEntropyLoss.__module__.startswith("w2v2_speaker_b200.")

    gpu = torch.cuda.is_available()
    if gpu:
        ctx = contextlib.nullcontext(None)
        dev = torch.device("cuda", 0)
    else:
        from dryrun import dry_library
        ctx = dry_library()
        dev = torch.device("cpu")
    from w2v2_speaker_b200.layers.linear import SpeakerLinear

    def build(pooling, loss, **over):
        cfg = dict(wav2vec_hunggingface_id="facebook/wav2vec2-base", reset_weights=False, wav2vec_feature_encoder_only=False,
                   wav2vec_initially_frozen=False, num_frozen_steps=None, completely_freeze_feature_extractor=True,
                   hidden_fc_layers_out=[], embedding_layer_idx=-1, stat_pooling_type=pooling, test_stat_pooling_type=pooling,
                   activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                   mask_feature_length=10, mask_feature_prob=0.0, mask_time_length=10, mask_time_prob=0.0,
                   final_channel_mask_prob=0.0, final_channel_mask_width=5, explicit_stat_pool_embedding_size=None,
                   explicit_num_speakers=None)
        cfg.update(over)
        ctor = (lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30)) \
            if loss == "aam" else (lambda: CrossEntropyLoss())

        class _Eval:
            max_num_training_samples = 0
        return Wav2vec2FCModule(hyperparameters_to_save={}, cfg=Wav2vec2FCModuleConfig(**cfg), num_speakers=5994,
                                loss_fn_constructor=ctor, validation_pairs=[], test_pairs=[], evaluator=_Eval())

    out = {"gpu": gpu, "cases": {}}
    with ctx as lib:
        for pooling, loss in (("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")):
            m = build(pooling, loss)
            case = {"heads": [type(l[0]).__name__ for l in m.fc_list],
                    "wrapper": type(m.wav2vec).__module__, "pool": type(m.stat_pooling).__module__,
                    "loss": type(m.loss_fn).__module__}
            if gpu:
                from oracle.params import BASE, make_head_params, make_params
                fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_cfg0_b2_1s.npz"))
                tr = np.load(os.path.join(ROOT, "tests", "golden", "ref_train_b2_1s.npz"))
                case["fixture_keys"] = len(fx.files) + len(tr.files)
                m.wav2vec.model.load_state_dict(make_params(BASE, seed=0), strict=False)
            m = m.to(dev)
            wav = torch.randn(2, 16000, device=dev)
            labels = torch.tensor([3, 5], device=dev)
            m.eval()
            with torch.no_grad():
                emb = m.compute_speaker_embedding(wav)
                pred = m.compute_speaker_prediction(emb)
                loss_v, prob = m.loss_fn(pred, labels)
            case["eval_shapes"] = [list(emb.shape), list(pred.shape), list(prob.shape)]
            m.train()
            m.on_train_start()
            if lib is not None:
                lib.calls.clear()
            emb = m.compute_speaker_embedding(wav)
            pred = m.compute_speaker_prediction(emb)
            loss_v, prob = m.loss_fn(pred, labels)
            loss_v.backward()
            m.on_after_backward()
            named = dict(m.named_parameters())
            case["grads"] = {"encoder": named["wav2vec.model.encoder.layers.0.attention.q_proj.weight"].grad is not None,
                             "cnn": named["wav2vec.model.feature_extractor.conv_layers.0.conv.weight"].grad is not None,
                             "heads": all(p.grad is not None for n, p in named.items()
                                          if not n.startswith("wav2vec.") and p.requires_grad)}
            if lib is not None:
                c = collections.Counter(lib.calls)
                case["calls"] = {k: c[k] for k in ("w2v2_encoder_layer_fwd", "w2v2_encoder_layer_bwd", "w2v2_gemm_f16",
                                                   "w2v2_softmax_ce", "w2v2_aam_softmax_ce_ex", "w2v2_stat_pool")}
            out["cases"][f"{pooling}/{loss}"] = case

        # freeze protocol on the reference's own module (a13): encoder initially frozen, released after 2 steps; Lightning
        # calls .train() after a validation loop while the encoder is still frozen
        m = build("mean", "ce", wav2vec_initially_frozen=True, num_frozen_steps=2, attention_dropout=0.1, hidden_dropout=0.1,
                  feat_proj_dropout=0.1, layerdrop=0.0, mask_time_prob=0.05).to(dev)
        m.train()
        m.on_train_start()
        phases = []
        for step in (1, 2, 3):
            m.train()
            m.zero_grad(set_to_none=True)
            if lib is not None:
                lib.calls.clear()
            emb = m.compute_speaker_embedding(torch.randn(3, 16000, device=dev))
            pred = m.compute_speaker_prediction(emb)
            loss_v, _ = m.loss_fn(pred, torch.tensor([1, 2, 3], device=dev))
            loss_v.backward()
            named = dict(m.named_parameters())
            phases.append({"frozen": bool(m._is_wav2vec_frozen),
                           "encoder_grad": named["wav2vec.model.encoder.layers.0.attention.q_proj.weight"].grad is not None,
                           "cnn_grad": named["wav2vec.model.feature_extractor.conv_layers.0.conv.weight"].grad is not None,
                           "head_grad": m.fc_list[-1][0].weight.grad is not None,
                           "bwd_calls": collections.Counter(lib.calls)["w2v2_encoder_layer_bwd"] if lib is not None else None})
            m.on_after_backward()
        out["freeze"] = phases
    print("DROPIN_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
