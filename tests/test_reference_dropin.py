"""VERDICT r1 "prove the drop-in": the reference's OWN caller (`Wav2vec2FCModule`, imported unmodified from
/root/reference) constructed and driven on top of this package's classes after `integration.install()`.
Runs where the reference checkout exists (the build container; it does not travel to the GPU box, so there the numbers
of the same protocol are covered by the mirror module against the reference-generated fixtures, tests/test_gpu_modules.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.fixture(scope="module")
def result():
    if not os.path.isdir(os.path.join(REF, "src", "lightning_modules")):
        pytest.skip("the reference checkout is not on this machine")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_dropin_driver.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DROPIN_RESULT ")][-1]
    return json.loads(line[len("DROPIN_RESULT "):])


def test_reference_module_builds_on_the_mirrors(result):
    for name, case in result["cases"].items():
        assert case["wrapper"] == "w2v2_speaker_b200.models.wav2vec2", name
        assert case["pool"] == "w2v2_speaker_b200.layers.pooling", name
        assert case["loss"].startswith("w2v2_speaker_b200.optim.loss"), name
        # the classifier the reference builds itself (nn.Linear, wav2vec2_fc.py:199-210) was re-typed: no cuBLAS on the path
        assert all(h == "SpeakerLinear" for h in case["heads"]), (name, case["heads"])
    assert result["cases"]["mean/ce"]["heads"] == ["SpeakerLinear"]
    assert result["cases"]["attentive/aam"]["heads"] == []          # AAM head surgery removed the last FC layer


def test_reference_module_forward_loss_backward(result):
    dims = {"mean/ce": 768, "mean+std/aam": 1536, "attentive/aam": 1536}
    for name, case in result["cases"].items():
        emb, pred, prob = case["eval_shapes"]
        assert emb == [2, dims[name]] and prob == [2, 5994], (name, case["eval_shapes"])
        g = case["grads"]
        assert g["encoder"] and g["heads"] and not g["cnn"], (name, g)       # CNN frozen by on_train_start
        if not result["gpu"]:
            c = case["calls"]
            assert c["w2v2_encoder_layer_fwd"] == 12 and c["w2v2_encoder_layer_bwd"] == 12, (name, c)
            assert c["w2v2_gemm_f16"] >= 1                                     # classifier / cosine GEMM on the tensor cores
            assert (c["w2v2_softmax_ce"] if name.endswith("/ce") else c["w2v2_aam_softmax_ce_ex"]) == 1, (name, c)


def test_reference_module_freeze_protocol(result):
    p1, p2, p3 = result["freeze"]
    assert p1["frozen"] and p2["frozen"] and not p3["frozen"]
    assert not p1["encoder_grad"] and not p2["encoder_grad"] and p3["encoder_grad"]
    assert p1["head_grad"] and p2["head_grad"] and p3["head_grad"]
    assert not p3["cnn_grad"]
    if not result["gpu"]:
        assert [p["bwd_calls"] for p in (p1, p2, p3)] == [0, 0, 12]
