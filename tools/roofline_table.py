"""Per-kernel roofline table of ONE cfg1 train step from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(the step between two Adam launches, as tools/launch_summary.py cuts it): every launch is attributed to its role by
its position in the schedule, its algorithmic FLOPs / bytes are taken from the cfg1 shapes (B = 64, T = 149,
M = B T = 9536, H = 768, FF = 3072, 12 heads of 64), and the achieved rate is set against the measured peaks.
ncu times are cold-cache and serialised: a guide to which kernel sits how far from its bound, not a step time.
usage: python tools/roofline_table.py profiles/r01_g_train_launches.csv [tensor_peak_TFLOPs hbm_peak_GBs]"""
import collections
import csv
import re
import sys

B, T, H, FF, HEADS, D = 64, 149, 768, 3072, 12, 64
M = B * T
CONV_L = [9599, 4799, 2399, 1199, 599, 299, 149]
CONV_K = [10, 3, 3, 3, 3, 2, 2]
C = 512


N_LAYER = 4 * H * H + 2 * H * FF + 9 * H + FF            # parameters of one transformer layer
N_ENC = 12 * N_LAYER + H * (H // 16) * 128 + 128 + H + 2 * H + 512 * H + H + 2 * 512 + H      # everything behind the CNN
N_HEAD = 5994 * H + 5994


def gemm(m, n, k):
    return 2.0 * m * n * k


def rows_of_step(path):
    rows = []
    for r in csv.reader(l for l in open(path) if l.startswith('"')):
        if r[0] == "ID":
            continue
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        rows.append((name, float(r[-1]) / 1e3))
    idx = [i for i, r in enumerate(rows) if "adam_kernel" in r[0]]
    starts = [i for k, i in enumerate(idx) if k == 0 or idx[k - 1] != i - 1]
    return rows[starts[0] + 2:starts[1] + 2]


def classify(rows):
    """-> list of (role, us, flops, bytes); flops / bytes None where the kernel is not modelled."""
    out = []
    bwd = False
    conv_i = 1
    prev = ""
    wg = 0          # wgrad position inside a layer's backward: FFN2, FFN1, out_proj, QKV
    wg_left = 4 * sum(1 for n, _ in rows if "attention_bwd" in n)      # the layers that ran (LayerDrop skips some)
    act = M * H
    for name, us in rows:
        role, fl, by = None, None, None
        if "softmax_ce_kernel" in name:
            bwd = True
        if "gemm_tc_kernel<128, 0, 1, 2" in name:
            role, fl = "conv0 GEMM + GroupNorm affine + GELU (HBM-bound; bytes = waveform + f16 output)", gemm(B * CONV_L[0], C, 64)
            by = B * (48000 * 4 + CONV_L[0] * C * 2)
        elif "gemm_tc_kernel<256, 0, 1, 0, 2" in name or "gemm_tc_kernel<256, 1, 1, 0, 2" in name:
            role, fl = f"conv{conv_i} (k={CONV_K[conv_i]}) + GELU", gemm(B * CONV_L[conv_i], C, CONV_K[conv_i] * C)
            conv_i += 1
        elif "gemm_tc_kernel<256, 0, 0, 1, 2" in name:
            role, fl = "fwd QKV projection (+bias, f16 out)", gemm(M, 3 * H, H)
        elif "gemm_tc_kernel<256, 0, 1, 1, 2, 1, 1>" in name:
            role, fl = "fwd FFN1, dual epilogue (z and gelu(z))", gemm(M, FF, H)
        elif "gemm_tc_kernel<256, 0, 1, 1, 2, 1, 2>" in name:
            role, fl = "fwd FFN1, dual epilogue (gelu(z) and gelu'(z))", gemm(M, FF, H)
        elif "gemm_tc_kernel<256, 0, 2, 0, 2" in name:
            role, fl = "bwd FFN2 data gradient, GELU-backward + bias-gradient epilogue", gemm(M, FF, H)
        elif "gemm_tc_kernel<256, 0, 3, 0, 2" in name:
            role, fl = "bwd FFN2 data gradient, multiply-by-gelu' + bias-gradient epilogue", gemm(M, FF, H)
        elif "gemm_tc_kernel<256, 0, 0, 0, 2" in name:
            role, fl = "bwd out_proj data gradient (f16 out)", gemm(M, H, H)
        elif "gemm_tc_kernel<256, 1, 0, 0, 2" in name and us > 12:
            if "attention_bwd" in prev:
                role, fl = "bwd QKV data gradient (f32, accumulating)", gemm(M, H, 3 * H)
            elif not bwd:
                role, fl = ("fwd out_proj (f32 out, stream-K)", gemm(M, H, H)) if "attention" in prev else \
                           ("fwd FFN2 (f32 out, stream-K)", gemm(M, H, FF))
            elif "wgrad" in prev and wg == 2:
                role, fl = "bwd FFN1 data gradient (f32 out)", gemm(M, H, FF)
            elif "wgrad" in prev and wg == 0:
                role, fl = "bwd QKV data gradient (f32 out)", gemm(M, H, 3 * H)
        elif "gemm_wgrad_kernel" in name and us > 12 and wg_left > 0:
            wg_left -= 1
            role, fl = [("bwd FFN2 weight gradient", gemm(M, H, FF)), ("bwd FFN1 weight gradient", gemm(M, FF, H)),
                        ("bwd out_proj weight gradient", gemm(M, H, H)), ("bwd QKV weight gradient", gemm(M, 3 * H, H))][wg]
            wg = (wg + 1) % 4
        elif name.endswith("attention_kernel"):
            role, fl, by = "attention forward (dropout on)", 4.0 * B * HEADS * T * T * D, act * 3 * 2 + act * 2 + B * HEADS * T * 4
        elif "attention_persist_kernel" in name:
            role, fl, by = "attention forward, persistent (dropout on)", 4.0 * B * HEADS * T * T * D, act * 3 * 2 + act * 2 + B * HEADS * T * 4
        elif "attention_bwd_fused_kernel" in name or "attention_bwd_persist_kernel" in name:
            role, fl, by = "attention backward" + (", persistent" if "persist" in name else ""), 10.0 * B * HEADS * T * T * D, act * 3 * 2 * 2 + act * 2 * 2
        elif "layernorm_kernel<1, 6, 1>" in name:
            role, by = "LayerNorm forward (+bias +residual, f32 + f16 out)", act * (4 + 4 + 4 + 2)
        elif "layernorm_bwd_kernel<1, 6, 1, 1>" in name or "layernorm_bwd_ring_kernel<6>" in name:
            # round 2: the data-gradient GEMMs accumulate into the residual gradient -> ONE gradient stream in (14 B / element);
            # pass "two" as argv[4] for launch lists taken with W2V2_DGRAD_ACCUM=0 (18 B / element)
            two = len(sys.argv) > 4 and sys.argv[4] == "two"
            role = "LayerNorm backward from the output (%s in, f32 + f16 out)" % ("two gradient branches" if two else "one gradient stream")
            by = act * ((4 + 4 + 4 + 4 + 2) if two else (4 + 4 + 4 + 2))
        elif name.endswith("posconv_kernel"):
            role, fl = "positional conv (forward / data gradient)", gemm(M, H, 128 * (H // 16))
        elif "posconv_wgrad_kernel" in name:
            role, fl = "positional conv weight gradient", gemm(M, H, 128 * (H // 16))
        elif "adam_kernel" in name:
            # p, g, m, v read + p, m, v written + g cleared = 32 B per parameter; the two launches (encoder behind the
            # frozen CNN, heads) are modelled together: bytes go to the role, split evenly over its launches
            role, by = "Adam (fused update + gradient clear), encoder + head parameters", 32.0 * (N_ENC + N_HEAD) / 2
        elif "prepare_weights_kernel" in name or "prepare_weights_v2_kernel" in name or "prepare_weights_v3_kernel" in name:
            # fp32 master read once, fp16 operand copy + transposed fp16 copy (data-gradient operand) written
            role, by = "re-derivation of the fp16 operand copies (one batched launch)", N_ENC * (4 + 2 + 2)
        out.append((role or "other: " + name.replace("w2v2::", "")[:40], us, fl, by))
        if "layernorm" not in name and "dropout" not in name:
            prev = name
    return out


def main():
    tensor_peak = float(sys.argv[2]) if len(sys.argv) > 2 else 1373.2
    hbm_peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6549.0
    agg = collections.OrderedDict()
    for role, us, fl, by in classify(rows_of_step(sys.argv[1])):
        a = agg.setdefault(role, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += fl or 0.0; a[3] += by or 0.0
    total = sum(a[1] for a in agg.values())
    print(f"# peaks: tensor {tensor_peak:.0f} TFLOP/s (measured sustained bf16 cuBLAS; fp16 = same rate), HBM {hbm_peak:.0f} GB/s")
    print(f"# {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms summed cold-cache kernel time")
    print(f"{'role':84s} {'n':>3s} {'us each':>8s} {'share':>6s} {'TFLOP/s':>8s} {'of peak':>7s} {'GB/s':>6s} {'of peak':>7s}")
    other = [0, 0.0]
    for role, (n, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if role.startswith("other:") and us < 60:
            other[0] += n; other[1] += us
            continue
        tf = fl / us / 1e6 if fl else None
        gb = by / us / 1e3 if by else None
        print(f"{role:84s} {n:3d} {us / n:8.1f} {100 * us / total:5.1f}% "
              f"{(f'{tf:8.0f}' if tf else '       -')} {(f'{100 * tf / tensor_peak:6.0f}%' if tf else '      -')} "
              f"{(f'{gb:6.0f}' if gb else '     -')} {(f'{100 * gb / hbm_peak:6.0f}%' if gb else '      -')}")
    print(f"{'other (' + str(other[0]) + ' small launches)':84s} {other[0]:3d} {other[1] / max(other[0], 1):8.1f} {100 * other[1] / total:5.1f}%")


if __name__ == "__main__":
    main()
