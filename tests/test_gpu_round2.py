"""Round-2 GPU tests (VERDICT r1 "next round" items 2, 5, 8, 9):
* an outer GradScaler (`precision: 16`) does not overflow the internal fp16 gradient operands,
* fp16 operand RANGE under pretrained-like activation magnitudes,
* oracle anchors at BASELINE.json's full size for 8 of the 64 utterances (cfg1) and for cfg2 (attentive + AAM),
* a live transformers state dict renamed to 4.x keys loads through checkpoint.py and reproduces the HF forward,
* (two GPUs) an N-rank trainer step equals the 1-rank step on the concatenated batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
S = 5994
ZERO_REG = dict(activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, hidden_dropout=0.0, layerdrop=0.0,
                mask_time_prob=0.0, mask_feature_prob=0.0)


def _need_cuda(n=1):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} CUDA device(s)")


def rows(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()


def _module(pooling, loss, base_params, **cfg_over):
    from oracle.params import make_asp_params, make_head_params
    from w2v2_speaker_b200.optim.loss import AngularAdditiveMarginSoftMaxLoss, CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling, **{**ZERO_REG, **cfg_over})
    ctor = CrossEntropyLoss if loss == "ce" else (
        lambda: AngularAdditiveMarginSoftMaxLoss(input_features=1, output_features=1, margin=0.2, scale=30))
    m = Wav2vec2FCModule(cfg, S, ctor)
    m.wav2vec.model.load_state_dict(base_params, strict=False)
    head = make_head_params(768 if pooling == "mean" else 1536, S, seed=1)
    with torch.no_grad():
        if loss == "ce":
            m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
        else:
            m.loss_fn.fc_weights.copy_(head["aam.fc_weights"])
        if pooling == "attentive":
            m.stat_pooling.pooling_layer.load_state_dict(make_asp_params(768, seed=2), strict=False)
    return m.cuda()


@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("attentive", "aam")])
def test_outer_grad_scaler_scale_is_absorbed(base_params, pooling, loss):
    """Lightning's native AMP (`precision: 16`, R:config/experiment/speaker_wav2vec2_aam.yaml:17) multiplies the loss by the
    GradScaler's scale, 2^16 at the start.  The backward used to carry that factor into its fp16 gradient copies (inf until
    the scaler had backed off to ~2^10).  Now each Function normalises its incoming gradient by a device-chosen power of two:
    the gradients of scale * loss must be finite and equal scale * (gradients of loss)."""
    _need_cuda()
    from oracle.params import make_inputs
    m = _module(pooling, loss, base_params).train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    wav, labels = make_inputs(2, 16000, S, seed=1234)
    x, y = wav[:, None, :].cuda(), labels.cuda()

    def grads(scale):
        m.zero_grad(set_to_none=True)
        if pooling == "attentive":          # same BatchNorm running statistics for both runs
            m.stat_pooling.pooling_layer.tdnn.norm.norm.reset_running_stats()
        emb, pred = m(x)
        loss_v, _ = m.loss_fn(pred, y)
        (loss_v * scale).backward()
        return {k: q.grad.detach().double().cpu() for k, q in m.named_parameters() if q.grad is not None}

    g1 = grads(1.0)
    for scale in (65536.0, 2.0 ** -12):
        gs = grads(scale)
        assert set(gs) == set(g1) and len(g1) > 190
        for k, g in gs.items():
            assert torch.isfinite(g).all(), (scale, k)
            # exactly 0 in exact arithmetic (softmax shift invariance over keys / over time): rounding noise only
            if k.endswith("k_proj.bias") or k.endswith("pooling_layer.conv.conv.bias"):
                continue
            err = ((g / scale - g1[k]).norm() / g1[k].norm().clamp_min(1e-30)).item()
            assert err < 3e-3, (scale, k, err)


def test_fp16_operand_range_with_large_activations(base_params):
    """Random-init weights keep every activation O(1); pretrained wav2vec2 carries residual-stream / FFN activations two
    orders of magnitude larger.  Scale the weights so that the FFN hidden activations (kept in fp16), the qkv projections
    (fp16) and the residual stream grow ~64x and check the fp16 copies neither overflow nor lose the embedding parity."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import make_inputs
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2WrapperModule
    p = {k: v.clone() for k, v in base_params.items()}
    for l in range(12):
        pre = f"encoder.layers.{l}."
        p[pre + "feed_forward.intermediate_dense.weight"] *= 24.0          # FFN hidden ~ 24 x sqrt-ish larger
        p[pre + "feed_forward.intermediate_dense.bias"] += 2.0
        p[pre + "feed_forward.output_dense.weight"] *= 4.0                # residual branch ~ 100 x
        p[pre + "attention.v_proj.weight"] *= 16.0
        p[pre + "attention.q_proj.weight"] *= 3.0
        p[pre + "final_layer_norm.weight"] *= 8.0                          # residual stream (fp32) and its fp16 copy ~ 8 x
        p[pre + "final_layer_norm.bias"] += 4.0
    wav, _ = make_inputs(2, 16000, S, seed=77)
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False)
    w.model.load_state_dict(p, strict=False)
    w = w.cuda().eval()
    with torch.no_grad():
        out = w.model(wav.cuda(), output_hidden_states=True)
    hs = out.hidden_states
    assert all(torch.isfinite(h).all() for h in hs)
    peak = max(h.abs().max().item() for h in hs)
    assert peak > 30.0, f"the stress did not raise the activations (peak {peak})"
    emb = out.last_hidden_state.mean(1)
    with torch.no_grad():
        ref_emb = O.speaker_embedding(wav, p, "mean")
    assert torch.isfinite(emb).all()
    assert rows(emb, ref_emb) < 2e-3, rows(emb, ref_emb)


@pytest.mark.parametrize("pooling,loss", [("mean", "ce"), ("attentive", "aam")])
def test_full_size_batch_oracle_anchor_8_of_64(base_params, pooling, loss):
    """BASELINE.json configs[1] and configs[2] at full size (64 utterances of 3 s): 8 utterances spread over the batch against
    the CPU oracle (north_star: embeddings within 1e-3, arg-max speaker id bit-exact)."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import make_asp_params, make_head_params, make_inputs
    m = _module(pooling, loss, base_params).eval()
    wav, labels = make_inputs(64, 48000, S, seed=5)
    with torch.no_grad():
        emb, pred = m(wav[:, None, :].cuda())
        loss_v, prob = m.loss_fn(pred, labels.cuda())
    idx = [0, 9, 18, 27, 36, 45, 54, 63]
    torch.set_num_threads(max(8, torch.get_num_threads()))
    head = make_head_params(768 if pooling == "mean" else 1536, S, seed=1)
    with torch.no_grad():
        if pooling == "mean":
            ref = O.speaker_embedding(wav[idx], base_params, "mean")
            ref_logits, _, ref_sm = O.cross_entropy_head(ref, head["fc.weight"], head["fc.bias"], labels[idx])
        else:
            ref = O.speaker_embedding(wav[idx], base_params, "attentive", asp=make_asp_params(768, seed=2))
            ref_sm = O.aam_softmax(ref, head["aam.fc_weights"], labels[idx], 0.2, 30.0)[-1]
    assert rows(emb[idx], ref) < 1e-3, rows(emb[idx], ref)
    assert torch.equal(prob[idx].argmax(1).cpu(), ref_sm.argmax(1))


def test_transformers_4x_checkpoint_reproduces_the_hf_forward():
    """SURVEY 8f-3 on the GPU: a live `transformers.Wav2Vec2Model.state_dict()`, renamed to the 4.x keys the reference's pin
    would save (weight_g / weight_v) and wrapped like a Lightning checkpoint, loads through checkpoint.py and reproduces
    HF's own forward (eager attention, fp32 on the same GPU) within 1e-3."""
    _need_cuda()
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    from w2v2_speaker_b200 import checkpoint as C
    from w2v2_speaker_b200.models.wav2vec2 import Wav2Vec2WrapperModule
    torch.manual_seed(3)
    cfg = Wav2Vec2Config()
    cfg._attn_implementation = "eager"
    hf = Wav2Vec2Model(cfg).eval()
    with torch.no_grad():                                   # make the weight-norm gain non-trivial
        hf.encoder.pos_conv_embed.conv.parametrizations.weight.original0.mul_(1.0 + 0.1 * torch.rand(1, 1, 128))
    sd4 = {}
    for k, v in hf.state_dict().items():
        k4 = k.replace("parametrizations.weight.original0", "weight_g").replace("parametrizations.weight.original1", "weight_v")
        sd4["wav2vec2." + k4] = v.clone()
    sd4["lm_head.weight"] = torch.zeros(32, 768)            # a Wav2Vec2ForCTC-style file
    w = Wav2Vec2WrapperModule("facebook/wav2vec2-base", False)
    res = w.model.load_state_dict(C.convert_hf_state_dict(sd4), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    w = w.cuda().eval()
    wav = torch.randn(3, 24000, generator=torch.Generator().manual_seed(4))
    wav = (wav - wav.mean(1, keepdim=True)) / wav.std(1, keepdim=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    hf = hf.cuda()
    with torch.no_grad():
        ref = hf(wav.cuda()).last_hidden_state              # [B, T, H]
        got = w(wav.cuda()).transpose(1, 2)                 # wrapper returns [B, H, T]
    assert got.shape == ref.shape
    assert rows(got.mean(1), ref.mean(1)) < 1e-3
    assert rows(got.reshape(3, -1), ref.reshape(3, -1)) < 3e-3


# ---- two GPUs: data-parallel step == single-GPU step on the concatenated batch -------------------------------------------


def _dp_worker(rank, world, port, base_params, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle.params import make_inputs
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    torch.manual_seed(100 + rank)                           # ranks start DIFFERENT: the trainer must broadcast rank 0's state
    m = _module("mean", "ce", base_params).train()
    if rank != 0:
        with torch.no_grad():
            for q in m.parameters():
                q.add_(0.01)
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    tr = FlatAdamTrainer(m, lr=1e-4)
    B = 16
    wav, labels = make_inputs(world * B, 48000, S, seed=9)
    x, y = wav[:, None, :].cuda(), labels.cuda()
    lo = rank * B
    for _ in range(2):
        tr.step(x[lo:lo + B], y[lo:lo + B])
    tr.synchronize()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"p": tr.flat_p.cpu(), "m": tr.m.cpu()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_one_rank_step_on_the_concatenated_batch(base_params, tmp_path):
    """R:config/trainer/trainer.yaml:6-9 (`accelerator: ddp`) semantics on hardware: two ranks x 16 utterances, regularisation
    off, two FlatAdamTrainer steps (gradient all-reduce over NCCL, rank 1 deliberately initialised differently) against one
    rank on the 32-utterance batch: first Adam moments (= the averaged gradient) and parameters must agree."""
    _need_cuda(2)
    import torch.multiprocessing as mp
    from oracle.params import make_inputs
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    out = str(tmp_path / "dp.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(2, port, base_params, out), nprocs=2, join=True)
    got = torch.load(out)
    m = _module("mean", "ce", base_params).train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    tr = FlatAdamTrainer(m, lr=1e-4)
    wav, labels = make_inputs(32, 48000, S, seed=9)
    x, y = wav[:, None, :].cuda(), labels.cuda()
    for _ in range(2):
        tr.step(x, y)
    tr.synchronize()
    torch.cuda.synchronize()
    ref_m, ref_p = tr.m.cpu().double(), tr.flat_p.cpu().double()
    em = ((got["m"].double() - ref_m).norm() / ref_m.norm()).item()
    ep = ((got["p"].double() - ref_p).norm() / ref_p.norm()).item()
    assert em < 2e-3, em            # gradients: fp16 operands see batches of 16 vs 32 (other tile counts / atomics order)
    # parameters: two Adam steps of size lr * m / (sqrt(v) + eps) -- where a gradient element is rounding noise its sign,
    # hence a whole step of 1e-4, may differ; measured 2.4e-5 of the parameter norm
    assert ep < 1e-4, ep


@pytest.mark.parametrize("hf_id", ["facebook/wav2vec2-large", "facebook/wav2vec2-large-lv60"])
def test_large_architecture_training_gradients_match_oracle_autograd(hf_id):
    """BASELINE.json configs[4] architecture (wav2vec2-large: 24 layers, H = 1024, 16 heads, FFN 4096) + mean+std pooling +
    CE: one training step (B = 2, 1 s, regularisation off, CNN frozen) against autograd of the CPU oracle -- loss within
    1e-3, every parameter gradient within 1.5e-2 norm-wise (24 layers of fp16-operand arithmetic).  Also for the
    stable-layer-norm sibling (-lv60 / XLSR: LayerNorm conv layers, pre-LN encoder layers; training_stable.py)."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import arch_from_id, make_head_params, make_inputs, make_params
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    LARGE = arch_from_id(hf_id)
    assert LARGE.stable_layer_norm == ("lv60" in hf_id)
    params = make_params(LARGE, seed=3)
    head = make_head_params(2048, S, seed=1)
    cfg = Wav2vec2FCModuleConfig(wav2vec_hunggingface_id=hf_id, stat_pooling_type="mean+std",
                                 test_stat_pooling_type="mean+std", **ZERO_REG)
    m = Wav2vec2FCModule(cfg, S, CrossEntropyLoss)
    res = m.wav2vec.model.load_state_dict(params, strict=False)
    assert not res.unexpected_keys and not res.missing_keys
    with torch.no_grad():
        m.fc_list[-1][0].weight.copy_(head["fc.weight"]); m.fc_list[-1][0].bias.copy_(head["fc.bias"])
    m = m.cuda().train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)
    wav, labels = make_inputs(2, 16000, S, seed=1234)
    emb, pred = m(wav[:, None, :].cuda())
    loss, prob = m.loss_fn(pred, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(max(8, torch.get_num_threads()))
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in params.items()}
    fw, fb = head["fc.weight"].clone().requires_grad_(True), head["fc.bias"].clone().requires_grad_(True)
    ref_emb = O.speaker_embedding(wav, p, "mean+std", arch=LARGE)
    _, ref_loss, ref_sm = O.cross_entropy_head(ref_emb, fw, fb, labels)
    ref_loss.backward()
    assert rows(emb.detach(), ref_emb.detach()) < 1.5e-3
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-3
    assert torch.equal(prob.argmax(1).cpu(), ref_sm.argmax(1))
    got = dict(m.wav2vec.model.named_parameters())
    worst = (0.0, None)
    for k, v in p.items():
        if k.startswith("feature_extractor") or k == "masked_spec_embed":
            continue
        g, r = got[k].grad.detach().cpu().double(), v.grad.double()
        if k.endswith("k_proj.bias"):
            # exactly 0 in exact arithmetic: both sides are rounding noise (fp16 operands here, fp32 there)
            scale = p[k.replace("k_proj", "q_proj")].grad.double().norm()
            assert g.norm() < 5e-2 * scale and r.norm() < 5e-2 * scale, k
            continue
        err = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
        worst = max(worst, (err, k))
        assert err < 1.5e-2, (k, err)
    lin = m.fc_list[-1][0]
    assert ((lin.weight.grad.cpu().double() - fw.grad.double()).norm() / fw.grad.double().norm()).item() < 1e-2
    print("worst LARGE parameter-gradient error", hf_id, worst)
    if LARGE.stable_layer_norm:
        # the same module with the reference's default regularisation (dropouts, LayerDrop, SpecAugment): finite, non-zero
        m2 = Wav2vec2FCModule(Wav2vec2FCModuleConfig(wav2vec_hunggingface_id=hf_id, layerdrop=0.1), S, CrossEntropyLoss)
        m2.wav2vec.model.load_state_dict(params, strict=False)
        m2 = m2.cuda().train()
        m2.wav2vec.model.feature_extractor.requires_grad_(False)
        _, pred2 = m2(wav[:, None, :].cuda())
        loss2, _ = m2.loss_fn(pred2, labels.cuda())
        loss2.backward()
        q = dict(m2.named_parameters())["wav2vec.model.encoder.layers.23.attention.q_proj.weight"]
        assert torch.isfinite(loss2) and q.grad is not None and torch.isfinite(q.grad).all()
        m2.wav2vec.model.feature_extractor.requires_grad_(True)
        with pytest.raises(NotImplementedError, match="CNN frozen"):
            m2(wav[:, None, :].cuda())


def test_ensemble_embedding_matches_oracle_hidden_states(base_params):
    """`use_transformers_as_ensembles` (R:src/lightning_modules/speaker/wav2vec2_fc.py:440-463): one pooled embedding per
    encoder output for the last `num_ensembles` of the 13 hidden states."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import make_inputs
    m = _module("mean", "ce", base_params, use_transformers_as_ensembles=True, num_ensembles=4).eval()
    assert m.test_with_ensemble
    wav, _ = make_inputs(2, 16000, S, seed=21)
    with torch.no_grad():
        got = m.compute_ensemble_embedding(wav[:, None, :].cuda())
        trace = {}
        O.wav2vec2_forward(wav, base_params, trace=trace)
    ref = trace["hidden_states"]
    assert len(got) == 4 and len(ref) == 13
    for e, h in zip(got, ref[9:13]):
        assert e.shape == (2, 768)
        assert rows(e, h.mean(1)) < 1e-3


def test_raw_int16_input_with_the_normaliser_folded_into_conv0(base_params):
    """16-bit PCM goes in, the reference's InputNormalizer2D ((x - mean) / (std + 1e-5) per utterance,
    R:src/data/preprocess/input_normalisation.py:53-67) is folded into conv layer 0's GroupNorm affine.  Must equal the
    float path fed with the waveform normalised the reference's way -- evaluation embeddings, and a training step (CNN
    frozen, regularisation off) through the same autograd path."""
    _need_cuda()
    g = torch.Generator().manual_seed(11)
    pcm = (torch.randn(3, 24000, generator=g) * 3000).round().clamp(-32768, 32767).to(torch.int16)
    pcm[1] = (pcm[1].float() * 0.05 + 700).to(torch.int16)                   # a quiet utterance with a DC offset
    x = pcm.float() / 32768.0
    std, mean = torch.std_mean(x, dim=1, keepdim=True)
    x_norm = (x - mean) / (std + 1e-5)
    labels = torch.tensor([5, 17, 4000])
    m = _module("mean", "ce", base_params).eval()
    with torch.no_grad():
        e_ref = m.compute_speaker_embedding(x_norm[:, None, :].cuda())
        e_raw = m.compute_speaker_embedding(pcm[:, None, :].cuda())
        e_f32 = m.wav2vec.model(x.cuda(), normalize_input=True).last_hidden_state.mean(1)   # float, not yet normalised
    assert rows(e_raw, e_ref) < 3e-4, rows(e_raw, e_ref)
    assert rows(e_f32, e_ref) < 3e-4, rows(e_f32, e_ref)

    m.train()
    m.wav2vec.model.feature_extractor.requires_grad_(False)

    def grads(inp):
        m.zero_grad(set_to_none=True)
        emb, pred = m(inp)
        loss, _ = m.loss_fn(pred, labels.cuda())
        loss.backward()
        return loss.item(), {k: q.grad.detach().double().cpu() for k, q in m.named_parameters() if q.grad is not None}

    l_ref, g_ref = grads(x_norm[:, None, :].cuda())
    l_raw, g_raw = grads(pcm[:, None, :].cuda())
    assert abs(l_raw - l_ref) < 1e-4 * abs(l_ref)
    assert set(g_raw) == set(g_ref)
    for k, gr in g_ref.items():
        if k.endswith("k_proj.bias"):
            continue
        err = ((g_raw[k] - gr).norm() / gr.norm().clamp_min(1e-30)).item()
        assert err < 3e-3, (k, err)
    m.wav2vec.model.feature_extractor.requires_grad_(True)
    with pytest.raises(NotImplementedError):
        m(pcm[:, None, :].cuda())


def test_paired_model_trains_at_the_reference_crop_length(base_params):
    """ADVICE r1 / VERDICT r1 missing #6: the paired-input model at the reference's own pipeline (two 3 s crops ->
    1 + 149 + 1 + 149 + 1 = 301 frames, R:config/experiment/speaker_wav2vec2_pairs.yaml:5) is longer than one key tile
    set of the training attention kernels.  It evaluates AND trains: scores, loss and every gradient of a training
    step against the oracle's composition of the same steps (key-tiled attention forward, attention backward launched
    per key block), then a step with the reference's default regularisation (attention dropout on the long path)."""
    _need_cuda()
    import torch.nn.functional as F
    from oracle import w2v2_oracle as O
    from oracle.params import make_inputs
    from w2v2_speaker_b200.optim.loss import BinaryCrossEntropyLoss
    from w2v2_speaker_b200.paired_speaker_module import Wav2vec2PairedSpeakerModule, Wav2vec2PairedSpeakerModuleConfig
    try:
        from test_gpu_split_path import _compare_encoder_grads
    except ImportError:
        from tests.test_gpu_split_path import _compare_encoder_grads
    torch.manual_seed(3)
    m = Wav2vec2PairedSpeakerModule(Wav2vec2PairedSpeakerModuleConfig(**ZERO_REG), BinaryCrossEntropyLoss)
    m.wav2vec.model.load_state_dict(base_params)
    lin_w, lin_b = m.linear.weight.detach().clone(), m.linear.bias.detach().clone()
    m = m.cuda().eval()
    a, _ = make_inputs(2, 48000, seed=41)
    b, _ = make_inputs(2, 48000, seed=42)
    labels = torch.tensor([1, 0])

    torch.set_num_threads(8)
    p = {k: v.clone().requires_grad_(not k.startswith("feature_extractor")) for k, v in base_params.items()}
    lw, lb = lin_w.clone().requires_grad_(True), lin_b.clone().requires_grad_(True)
    ref_tokens = O.split_path_forward([a, b], p, [1.0, -1.0, -1.0])
    assert ref_tokens.shape[1] == 301
    ref_scores = F.linear(ref_tokens[:, 0, :], lw, lb)
    ref_loss = F.binary_cross_entropy_with_logits(ref_scores.squeeze(), labels.float())
    ref_loss.backward()

    with torch.no_grad():
        s = m(a.cuda(), b.cuda())
    assert s.shape == (2, 1) and torch.isfinite(s).all()
    assert (s.cpu() - ref_scores.detach()).abs().max().item() < 5e-3 * max(1.0, ref_scores.abs().max().item())
    m.train()
    m.on_train_start()
    out = m(a.cuda(), b.cuda())
    loss, _ = m.loss_fn(out, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 5e-3
    for g, r, k in ((m.linear.weight.grad, lw.grad, "linear.weight"), (m.linear.bias.grad, lb.grad, "linear.bias")):
        rel = ((g.cpu().double() - r.double()).norm() / r.double().norm()).item()
        assert rel < 1e-2, (k, rel)
    # k_proj.bias (exactly 0 in exact arithmetic): the cancelling sum has twice the terms of a 149-frame crop -- measured
    # 0.98e-2 of the q_proj.bias gradient at layer 9, so the noise bar is 3e-2 here
    # (the other gradients: 2e-2, the bar of the CLS-readout tests -- everything here flows through ONE token's attention)
    worst = _compare_encoder_grads(dict(m.wav2vec.model.named_parameters()), p, False, tol=2e-2, kbias_tol=3e-2)
    print("worst parameter-gradient error at 301 frames", worst)

    # default regularisation (attention dropout 0.1, LayerDrop, SpecAugment) on the long path: finite, non-zero
    torch.manual_seed(11)
    m2 = Wav2vec2PairedSpeakerModule(Wav2vec2PairedSpeakerModuleConfig(layerdrop=0.25), BinaryCrossEntropyLoss)
    m2.wav2vec.model.load_state_dict(base_params)
    m2 = m2.cuda().train()
    m2.on_train_start()
    loss2, _ = m2.loss_fn(m2(a.cuda(), b.cuda()), labels.cuda())
    loss2.backward()
    q = dict(m2.named_parameters())["wav2vec.model.encoder.layers.0.attention.q_proj.weight"]
    assert torch.isfinite(loss2) and q.grad is not None and torch.isfinite(q.grad).all() and q.grad.abs().max().item() > 0


def test_freeze_protocol_on_hardware(base_params):
    """`wav2vec_initially_frozen` + `num_frozen_steps` (R:src/lightning_modules/speaker/wav2vec2_fc.py:
    339-361) on the GPU, through the flat-buffer trainer built WHILE the encoder is frozen.  Step 1 (frozen): only the head
    moves, and its gradient equals the oracle's with the encoder treated as a constant.  on_after_backward() releases the
    encoder (the CNN stays frozen).  Step 2: every encoder parameter behind the CNN moves, and the gradients of that step
    equal autograd of the oracle at the post-step-1 parameters."""
    _need_cuda()
    from oracle import w2v2_oracle as O
    from oracle.params import make_inputs
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    m = _module("mean", "ce", base_params, wav2vec_initially_frozen=True, num_frozen_steps=1).train()
    m.on_train_start()
    assert not any(q.requires_grad for q in m.wav2vec.parameters())
    tr = FlatAdamTrainer(m, lr=1e-3)
    wav, labels = make_inputs(4, 16000, S, seed=77)
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    head_w0, head_b0 = m.fc_list[-1][0].weight.detach().clone(), m.fc_list[-1][0].bias.detach().clone()
    loss1, _ = tr.step(wav[:, None, :].cuda(), labels.cuda())
    m.on_after_backward()                                          # steps = 1 >= num_frozen_steps: release
    torch.cuda.synchronize()
    after1 = {k: v.detach().clone() for k, v in m.named_parameters()}
    moved1 = {k for k in before if not torch.equal(before[k], after1[k])}
    assert moved1 and all(not k.startswith("wav2vec.") for k in moved1), sorted(moved1)[:5]
    enc = dict(m.wav2vec.model.named_parameters())
    assert all(q.requires_grad == (not k.startswith("feature_extractor")) for k, q in enc.items())
    # oracle loss of step 1 (eval semantics = ZERO_REG training)
    torch.set_num_threads(8)
    lin = m.fc_list[-1][0]
    with torch.no_grad():
        ref_emb = O.speaker_embedding(wav, base_params, "mean")
        _, ref_loss1, _ = O.cross_entropy_head(ref_emb, head_w0.cpu(), head_b0.cpu(), labels)
    assert abs(loss1.item() - ref_loss1.item()) / ref_loss1.item() < 1e-3
    # step 2 with gradients kept: run the module by hand at the post-step-1 parameters and compare with oracle autograd
    emb, pred = m(wav[:, None, :].cuda())
    loss2, _ = m.loss_fn(pred, labels.cuda())
    p = {k: after1["wav2vec.model." + k].cpu().clone().requires_grad_(not k.startswith("feature_extractor")) for k in enc}
    fw = lin.weight.detach().cpu().clone().requires_grad_(True)
    fb = lin.bias.detach().cpu().clone().requires_grad_(True)
    ref_emb2 = O.speaker_embedding(wav, p, "mean")
    _, ref_loss2, _ = O.cross_entropy_head(ref_emb2, fw, fb, labels)
    ref_loss2.backward()
    assert abs(loss2.item() - ref_loss2.item()) / ref_loss2.item() < 1e-3
    # the trainer's second step moves the encoder (everything behind the CNN) and leaves the CNN alone
    loss2b, _ = tr.step(wav[:, None, :].cuda(), labels.cuda())
    torch.cuda.synchronize()
    assert abs(loss2b.item() - ref_loss2.item()) / ref_loss2.item() < 1e-3
    after2 = {k: v.detach().clone() for k, v in m.named_parameters()}
    for k in after1:
        if k.startswith("wav2vec.model.feature_extractor") or k.endswith("masked_spec_embed"):
            assert torch.equal(after1[k], after2[k]), k
        elif k.startswith("wav2vec.model."):
            assert not torch.equal(after1[k], after2[k]), k
            if k.endswith("k_proj.bias"):                      # gradient exactly 0 in exact arithmetic: noise on both sides
                continue
            # Adam's first step for a parameter moves it by lr * sign(g) (m / sqrt(v) = g / |g|): check the direction against
            # the oracle gradient where that gradient is not rounding noise
            g = p[k[len("wav2vec.model."):]].grad
            step = (after2[k] - after1[k]).cpu()
            big = g.abs() > 0.05 * g.abs().max()
            agree = (torch.sign(step[big]) == -torch.sign(g[big])).float().mean().item()
            assert agree > 0.99, (k, agree)
