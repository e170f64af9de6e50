"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF (build container only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Run:  python -m oracle.make_golden          (ref_cfg0_*, ref_b3_*)
                                                     python -m oracle.make_golden paired   (ref_paired_b3)
                                                     python -m oracle.make_golden train    (ref_train_b2_1s)
                                                     python -m oracle.make_golden eval     (ref_eval)

The reference (/root/reference, read-only, python) is imported unmodified under import
shims for the packages missing from this image (pytorch_lightning, omegaconf, torchmetrics,
speechbrain, fairseq, ...).  ``Wav2Vec2Model.from_pretrained`` cannot reach the
network, so it is patched to build the same architecture from ``Wav2Vec2Config`` and the
parameters are then overwritten with ``oracle.params.make_params(seed)`` -- the very
parameters the oracle and the CUDA engine are tested with.  The reference's own classes do
the arithmetic:
    Wav2vec2FCModule.compute_speaker_embedding / compute_speaker_prediction
        (R:src/lightning_modules/speaker/wav2vec2_fc.py:414-438)
    Wav2Vec2WrapperModule.forward (R:src/models/wav2vec2.py:126-146)
    MeanStatPool1D / MeanStdStatPool1D / AttentiveStatPool1D (R:src/layers/pooling.py)
    CrossEntropyLoss / AngularAdditiveMarginSoftMaxLoss (R:src/optim/loss/*.py)
speechbrain's AttentiveStatisticsPooling is absent offline; the shim provides a torch.nn
restatement of the published module (the only part of these fixtures not produced by
reference/HF code).

Fixtures are small: full tensors for embeddings / logits / last hidden state, and
(norm, mean, strided sample) summaries for the big intermediates.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

STUB_PREFIXES = ("pytorch_lightning", "omegaconf", "torchmetrics", "speechbrain", "fairseq", "hurry",
                 "webdataset", "comet_ml", "pl_bolts", "jiwer", "librosa", "yaspin", "hydra", "dotenv",
                 "torchaudio", "augment", "seaborn", "matplotlib", "pytorch_model_summary", "bob",
                 "wget", "tqdm_joblib", "optuna")


class _Anything:
    """Permissive placeholder: subclassable, callable, attribute-able."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _LightningModule(nn.Module):
    """pytorch_lightning.LightningModule stand-in: nn.Module + freeze/unfreeze semantics
    (all params requires_grad False + eval / True + train)."""
    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad = True
        self.train()

    def log(self, *a, **k):
        pass

    def save_hyperparameters(self, *a, **k):
        pass


class _ASP(nn.Module):
    """speechbrain.lobes.models.ECAPA_TDNN.AttentiveStatisticsPooling(channels) restated with
    torch.nn modules and speechbrain's parameter names (see oracle.w2v2_oracle.attentive_stat_pool)."""
    def __init__(self, channels, attention_channels=128, global_context=True):
        super().__init__()
        self.eps = 1e-12

        class _Wrap(nn.Module):
            def __init__(self, name, mod):
                super().__init__()
                setattr(self, name, mod)
        self.tdnn = nn.Module()
        self.tdnn.conv = _Wrap("conv", nn.Conv1d(channels * 3, attention_channels, 1))
        self.tdnn.norm = _Wrap("norm", nn.BatchNorm1d(attention_channels))
        self.conv = _Wrap("conv", nn.Conv1d(attention_channels, channels, 1))

    def forward(self, x, lengths=None):
        L = x.shape[-1]

        def stats(x, m, dim=2, eps=self.eps):
            mean = (m * x).sum(dim)
            std = torch.sqrt((m * (x - mean.unsqueeze(dim)).pow(2)).sum(dim).clamp(eps))
            return mean, std
        mask = torch.ones(x.shape[0], 1, L, dtype=x.dtype)
        total = mask.sum(dim=2, keepdim=True).float()
        mean, std = stats(x, mask / total)
        attn = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
        attn = self.tdnn.norm.norm(torch.relu(self.tdnn.conv.conv(attn)))
        attn = self.conv.conv(torch.tanh(attn))
        attn = attn.masked_fill(mask == 0, float("-inf"))
        attn = torch.softmax(attn, dim=2)
        mean, std = stats(x, attn)
        return torch.cat((mean, std), dim=1).unsqueeze(2)


def install_shims():
    # import transformers first: its optional-dependency probes must not see the stubs
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    import transformers.models.wav2vec2.modeling_wav2vec2  # noqa: F401
    sys.meta_path.insert(0, _StubFinder())
    import pytorch_lightning as pl
    pl.LightningModule = _LightningModule
    import pytorch_lightning.core.decorators as dec
    dec.auto_move_data = lambda f: f
    import omegaconf
    omegaconf.DictConfig = dict
    omegaconf.OmegaConf.to_container = staticmethod(lambda c, *a, **k: dict(c))
    import speechbrain.lobes.models.ECAPA_TDNN as ecapa
    ecapa.AttentiveStatisticsPooling = _ASP
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # random-init instead of from_pretrained (no network / no HF cache)

    def _from_pretrained(cls_or_id, *args, **kw):
        hid = cls_or_id if isinstance(cls_or_id, str) else args[0]
        kw.pop("gradient_checkpointing", None)
        size = {}
        if "large" in hid:
            size = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
        cfg = Wav2Vec2Config(**size, **kw)
        cfg._attn_implementation = "eager"
        return Wav2Vec2Model(cfg)
    Wav2Vec2Model.from_pretrained = staticmethod(_from_pretrained)


def build_reference_module(pooling: str, loss: str, num_speakers: int = 5994):
    """Construct the reference's Wav2vec2FCModule as R:src/main.py:223-285 would (network cfg
    R:config/network/wav2vec2_fc.yaml), with all stochastic regularisation at 0."""
    from src.lightning_modules.speaker.wav2vec2_fc import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    from src.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    cfg = Wav2vec2FCModuleConfig(
        wav2vec_hunggingface_id="facebook/wav2vec2-base", reset_weights=False,
        wav2vec_feature_encoder_only=False, wav2vec_initially_frozen=False, num_frozen_steps=None,
        completely_freeze_feature_extractor=True, hidden_fc_layers_out=[], embedding_layer_idx=-1,
        stat_pooling_type=pooling, test_stat_pooling_type=pooling,
        activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0,
        layerdrop=0.0, mask_feature_length=10, mask_feature_prob=0.0, mask_time_length=10, mask_time_prob=0.0,
        final_channel_mask_prob=0.0, final_channel_mask_width=5,
        explicit_stat_pool_embedding_size=None, explicit_num_speakers=None)
    if loss == "aam":
        ctor = lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30)
    else:
        ctor = lambda: CrossEntropyLoss()

    class _Eval:
        max_num_training_samples = 0
    m = Wav2vec2FCModule(hyperparameters_to_save={}, cfg=cfg, num_speakers=num_speakers,
                         loss_fn_constructor=ctor, validation_pairs=[], test_pairs=[], evaluator=_Eval())
    return m


def summarise(name: str, t: torch.Tensor, out: dict, n: int = 4096):
    t = t.detach().float()
    out[name + ".shape"] = np.array(t.shape, dtype=np.int64)
    out[name + ".norm"] = np.array(t.double().norm().item())
    out[name + ".mean"] = np.array(t.double().mean().item())
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    out[name + ".sample"] = flat[::step][:n].numpy().copy()


def main():
    install_shims()
    from oracle.params import BASE, make_asp_params, make_head_params, make_inputs, make_params
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    S = 5994
    params = make_params(BASE, seed=0)
    for tag, B, N in (("cfg0_b2_1s", 2, 16000), ("b3_ragged_0p7s", 3, 11283)):
        wav, labels = make_inputs(B, N, S, seed=1234)
        out = {"B": np.array(B), "N": np.array(N), "labels": labels.numpy(),
               "wav.sample": wav.reshape(-1)[::97][:2048].numpy().copy()}
        for pooling, loss in (("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")):
            E = 768 if pooling == "mean" else 1536
            m = build_reference_module(pooling, loss, S)
            res = m.wav2vec.model.load_state_dict(params, strict=False)
            # masked_spec_embed only exists when mask_time_prob > 0 (HF:1258-1262)
            assert not res.missing_keys and set(res.unexpected_keys) <= {"masked_spec_embed"}, res
            head = make_head_params(E, S, seed=1)
            if loss == "ce":
                m.fc_list[-1][0].weight.data.copy_(head["fc.weight"])
                m.fc_list[-1][0].bias.data.copy_(head["fc.bias"])
            else:
                m.loss_fn.fc_weights.data.copy_(head["aam.fc_weights"])
            if pooling == "attentive":
                asp = make_asp_params(768, seed=2)
                m.stat_pooling.pooling_layer.load_state_dict(asp, strict=False)
                sd = m.stat_pooling.pooling_layer.state_dict()
                for k, v in asp.items():
                    assert torch.equal(sd[k], v), k
            m.eval()
            with torch.no_grad():
                # the reference hands [B,1,N] batches (R:.../training_batch_speaker.py:44-75)
                emb = m.compute_speaker_embedding(wav[:, None, :])
                pred = m.compute_speaker_prediction(emb)
                loss_v, softmax = m.loss_fn(pred, labels)
                key = f"{pooling}.{loss}"
                out[key + ".embedding"] = emb.numpy()
                out[key + ".loss"] = np.array(loss_v.item())
                out[key + ".argmax"] = softmax.argmax(1).numpy()
                summarise(key + ".softmax", softmax, out)
                if loss == "ce":
                    out[key + ".logits"] = pred.numpy()
                    hf = m.wav2vec.model(wav, output_hidden_states=True)
                    out["last_hidden_state"] = hf.last_hidden_state.numpy()
                    for i, hs in enumerate(hf.hidden_states):
                        summarise(f"hidden_states.{i}", hs, out)
                    feat = m.wav2vec.model.feature_extractor(wav)
                    summarise("feature_extractor", feat, out)
                    h = wav[:, None, :]
                    for i, layer in enumerate(m.wav2vec.model.feature_extractor.conv_layers):
                        h = layer(h)
                        summarise(f"conv.{i}", h, out)
            print(tag, key, "emb", tuple(emb.shape), "loss", float(loss_v))
        np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
        print("wrote", tag, sum(v.nbytes for v in out.values()) / 1e6, "MB")


def paired_main():
    """tests/golden/ref_paired_b3.npz: the reference's sibling heads on the same encoder, run by the reference's own
    classes -- ``Wav2vec2PairedSpeakerModule.compute_speaker_equality`` + ``BinaryCrossEntropyLoss``
    (R:src/lightning_modules/speaker/wav2vec2_paired_input.py:162-207, R:src/optim/loss/binary_cross_entropy.py) for one
    training step (scores, loss, and the gradients torch autograd gives the reference), and the CLS-token path of
    ``Wav2Vec2WrapperModule`` (R:src/models/wav2vec2.py:128-140)."""
    install_shims()
    from oracle.params import BASE, make_inputs, make_params
    from src.lightning_modules.speaker.wav2vec2_paired_input import (Wav2vec2PairedSpeakerModule,
                                                                      Wav2vec2PairedSpeakerModuleConfig)
    from src.models.wav2vec2 import Wav2Vec2RegularisationConfig, Wav2Vec2WrapperModule
    from src.optim.loss.binary_cross_entropy import BinaryCrossEntropyLoss
    torch.set_num_threads(8)
    params = make_params(BASE, seed=0)
    B = 3
    wav_a, _ = make_inputs(B, 16000, seed=31)
    wav_b, _ = make_inputs(B, 11283, seed=32)
    labels = torch.tensor([1, 0, 1])
    zero = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                mask_feature_length=10, mask_feature_prob=0.0, mask_time_length=10, mask_time_prob=0.0)
    cfg = Wav2vec2PairedSpeakerModuleConfig(
        wav2vec_hunggingface_id="facebook/wav2vec2-base", reset_weights=False, wav2vec_initially_frozen=False,
        num_frozen_steps=None, completely_freeze_feature_extractor=True, completely_freeze_feature_projector=False,
        final_channel_mask_prob=0.0, final_channel_mask_width=5, **zero)
    torch.manual_seed(3)
    m = Wav2vec2PairedSpeakerModule(hyperparameters_to_save={}, cfg=cfg, loss_fn_constructor=BinaryCrossEntropyLoss)
    res = m.wav2vec.model.load_state_dict(params, strict=False)
    assert not res.missing_keys and set(res.unexpected_keys) <= {"masked_spec_embed"}, res
    out = {"labels": labels.numpy(), "linear.weight": m.linear.weight.detach().numpy().copy(),
           "linear.bias": m.linear.bias.detach().numpy().copy()}
    m.eval()
    with torch.no_grad():
        out["scores.eval"] = m.compute_speaker_equality(wav_a, wav_b).numpy()
    m.train()
    m.on_train_start()
    scores = m.compute_speaker_equality(wav_a, wav_b)
    loss, prediction = m.loss_fn(scores, labels)
    loss.backward()
    out["scores.train"] = scores.detach().numpy()
    out["loss"] = np.array(loss.item())
    out["prediction"] = prediction.numpy()
    out["grad.linear.weight"] = m.linear.weight.grad.numpy().copy()
    out["grad.linear.bias"] = m.linear.bias.grad.numpy().copy()
    for k, q in m.wav2vec.model.named_parameters():
        if q.grad is not None:
            summarise("grad." + k, q.grad, out, n=256)
        else:
            assert k.startswith("feature_extractor."), k
    print("paired: scores", out["scores.train"].ravel(), "loss", float(loss))
    # CLS-token wrapper path (eval)
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False, reg_cfg=Wav2Vec2RegularisationConfig(**zero),
                              insert_clc_token=True)
    res = w.model.load_state_dict(params, strict=False)
    assert not res.missing_keys and set(res.unexpected_keys) <= {"masked_spec_embed"}, res
    w.eval()
    with torch.no_grad():
        cls_out = w(wav_b)                                       # [B, 768, 35 + 1]
    out["cls.first_token"] = cls_out[:, :, 0].numpy()
    summarise("cls.output", cls_out, out)
    np.savez_compressed(os.path.join(OUT, "ref_paired_b3.npz"), **out)
    print("wrote ref_paired_b3.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB")


def train_main():
    """tests/golden/ref_train_b2_1s.npz: one TRAINING step of the reference's own Wav2vec2FCModule (regularisation off,
    CNN frozen as R:config/network/wav2vec2_fc.yaml:16 has it) for the three pooling / loss pairs of the forward fixtures:
    loss and every parameter gradient torch autograd gives the reference (norm + strided sample each)."""
    install_shims()
    from oracle.params import BASE, make_asp_params, make_head_params, make_inputs, make_params
    torch.set_num_threads(8)
    S = 5994
    params = make_params(BASE, seed=0)
    wav, labels = make_inputs(2, 16000, S, seed=1234)
    out = {"labels": labels.numpy()}
    for pooling, loss in (("mean", "ce"), ("mean+std", "aam"), ("attentive", "aam")):
        E = 768 if pooling == "mean" else 1536
        m = build_reference_module(pooling, loss, S)
        res = m.wav2vec.model.load_state_dict(params, strict=False)
        assert not res.missing_keys and set(res.unexpected_keys) <= {"masked_spec_embed"}, res
        head = make_head_params(E, S, seed=1)
        if loss == "ce":
            m.fc_list[-1][0].weight.data.copy_(head["fc.weight"])
            m.fc_list[-1][0].bias.data.copy_(head["fc.bias"])
        else:
            m.loss_fn.fc_weights.data.copy_(head["aam.fc_weights"])
        if pooling == "attentive":
            m.stat_pooling.pooling_layer.load_state_dict(make_asp_params(768, seed=2), strict=False)
        m.train()
        m.on_train_start()                                         # freezes the feature extractor
        emb = m.compute_speaker_embedding(wav[:, None, :])
        pred = m.compute_speaker_prediction(emb)
        loss_v, _ = m.loss_fn(pred, labels)
        loss_v.backward()
        key = f"{pooling}.{loss}"
        out[key + ".loss"] = np.array(loss_v.item())
        # packed (one zip entry per array costs more than a small gradient summary): names, norms, strided samples
        names, norms, samples = [], [], []
        for k, q in m.named_parameters():
            if q.grad is not None:
                flat = q.grad.detach().float().reshape(-1)
                step = max(1, flat.numel() // 128)
                smp = np.zeros(128, dtype=np.float32)
                got = flat[::step][:128].numpy()
                smp[:got.size] = got
                names.append(k); norms.append(q.grad.double().norm().item()); samples.append(smp)
        out[key + ".grad.names"] = np.array(names)
        out[key + ".grad.norms"] = np.array(norms)
        out[key + ".grad.samples"] = np.stack(samples)
        print(key, "loss", float(loss_v.item()), "gradients", len(names))
    np.savez_compressed(os.path.join(OUT, "ref_train_b2_1s.npz"), **out)
    print("wrote ref_train_b2_1s.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB")


def eval_main():
    """tests/golden/ref_eval.npz: the reference's own CosineDistanceEvaluator.evaluate (with and without centring) and
    calculate_eer / calculate_mdc (R:src/evaluation/speaker/cosine_distance.py, R:src/eval_metrics.py) on seeded
    embeddings and trial lists."""
    install_shims()
    import contextlib
    import io
    from src.eval_metrics import calculate_eer, calculate_mdc
    from src.evaluation.speaker.cosine_distance import CosineDistanceEvaluator
    from src.evaluation.speaker.speaker_recognition_evaluator import EmbeddingSample, EvaluationPair
    g = torch.Generator().manual_seed(9)
    n_spk, per, E = 30, 5, 96
    centers = torch.randn(n_spk, E, generator=g)
    emb = (centers[:, None, :] + 4.0 * torch.randn(n_spk, per, E, generator=g) + 0.5).reshape(-1, E)
    ids = [f"spk{s}/utt{u}" for s in range(n_spk) for u in range(per)]
    rng = np.random.default_rng(4)
    left, right = [], []
    while len(left) < 2000:
        a, b = rng.integers(0, len(ids), 2)
        if a != b:
            left.append(int(a)); right.append(int(b))
    pairs = [EvaluationPair(ids[a].split("/")[0] == ids[b].split("/")[0], ids[a], ids[b]) for a, b in zip(left, right)]
    samples = [EmbeddingSample(i, e) for i, e in zip(ids, emb)]
    out = {"embeddings": emb.numpy(), "left": np.array(left), "right": np.array(right),
           "same": np.array([p.same_speaker for p in pairs]), "fit_stride": np.array(3)}
    for center in (False, True):
        ev = CosineDistanceEvaluator(center_before_scoring=center, length_norm_before_scoring=True,
                                     max_num_training_samples=0)
        ev.fit_parameters(list(emb[::3]), [])
        with contextlib.redirect_stdout(io.StringIO()):           # the reference prints the centring statistics
            res = ev.evaluate(pairs, samples)
        key = "center" if center else "plain"
        for k, v in res.items():
            out[f"{key}.{k}"] = np.array(float(v))
        print(key, {k: float(v) for k, v in res.items()})
    # the metric functions on their own, on scores with ties and an awkward operating point
    gt = rng.integers(0, 2, 1500).tolist()
    pred = np.round(np.clip(rng.normal(0.45, 0.2, 1500) + 0.15 * np.array(gt), 0, 1), 3).tolist()
    out["metric.gt"], out["metric.pred"] = np.array(gt), np.array(pred)
    e, et = calculate_eer(gt, pred)
    m, mt = calculate_mdc(gt, pred)
    out["metric.eer"], out["metric.eer_threshold"] = np.array(float(e)), np.array(float(et))
    out["metric.mdc"], out["metric.mdc_threshold"] = np.array(float(m)), np.array(float(mt))
    print("metrics", float(e), float(et), float(m), float(mt))
    np.savez_compressed(os.path.join(OUT, "ref_eval.npz"), **out)
    print("wrote ref_eval.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB")


if __name__ == "__main__":
    {"paired": paired_main, "train": train_main, "eval": eval_main}.get((sys.argv[1:] or [""])[0], main)()
