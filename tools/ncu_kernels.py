"""One launch of each hot non-attention kernel at the cfg1 shapes (M = 64 x 149, H = 768, FF = 3072), for
`ncu --set full`: FFN1 dual-epilogue GEMM, FFN2 data gradient with the multiply epilogue, an accumulating data-gradient
GEMM, the QKV weight gradient, the LayerNorm backward from the output, conv layer 0's stage, the weight re-preparation.
usage: ncu --set full --import-source on --clock-control none -o out python tools/ncu_kernels.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from w2v2_speaker_b200 import ops
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
M, H, FF = 64 * 149, 768, 3072
g = torch.Generator().manual_seed(0)
r16 = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev).half()
r32 = lambda *s: torch.randn(*s, generator=g).to(dev)
h16, w1, b1 = r16(M, H), r16(FF, H), r32(FF)
dx16, w2t, mul = r16(M, H), r16(FF, H), r16(M, FF)
colsum = torch.zeros(FF, device=dev)
dz16, w1t, acc = r16(M, FF), r16(H, FF), r32(M, H)
dqkv16, hin16, dwqkv = r16(M, 3 * H), r16(M, H), torch.zeros(3 * H, H, device=dev)
dy, y32, rstd, gam, bet = r32(M, H), r32(M, H), torch.rand(M, generator=g).to(dev) + 0.5, torch.rand(H, generator=g).to(dev) + 0.5, r32(H)
dg, db, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
wav = torch.randn(64, 48000, generator=g).to(dev)
w0, gg, gb = torch.randn(512, 10, generator=g).to(dev) * 0.4, torch.ones(512, device=dev), torch.zeros(512, device=dev)
m = bench.build_module(dev, True)                     # for the weight re-preparation job table
eng = m.wav2vec.model._engine()
m.wav2vec.model._train_weights(eng)
for _ in range(2):
    ops.gemm_f16_dual_gelu_grad(h16, w1, b1)
    ops.gemm_f16_mul_colsum(dx16, w2t, mul, colsum)
    ops.call("w2v2_gemm_f16_accum", ops.ptr(dz16), M, FF, FF, ops.ptr(w1t), FF, H, ops.ptr(acc), H, ops.stream_ptr())
    ops.gemm_wgrad_f16(dqkv16, hin16, dwqkv)
    ops.layernorm_bwd_from_output(dy, y32, rstd, gam, bet, dgamma=dg, dbeta=db, dbias=dbias, drop_p=0.1, drop_seed=3)
    ops.conv0_gn_gelu(wav, w0, gg, gb)
    eng.w.prep.run()
torch.cuda.synchronize()
