"""Round-2 A/B measurements on one B200 (JSON lines):
  * decode kernel variants (sb200_gptq4_set_decode bit mask: 1 bulk-copy slab, 2 programmatic launch, 4 L2 prefetch, high
    nibble = CTAs per SM; with / without SB200_GPTQ4_STATIC_WEIGHTS) -- single LLaMA-7B linears and the fused-launch
    decoder layer at M = 1 / 4 / 16;
  * per-group tcgen05 kernel with the accumulator drained by tcgen05.ld .x16 vs pairs of .x8 (sb200_gptq4_set_tc_drain)
    at M = 256 / 512 / 2048, next to the tensor-memory-operand kernel at M = 2048.
usage: python scripts/exp/ab_r02.py [decode] [tc]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch

import bench_gptq as bg
from sparsebit_b200 import _lib, ops

dev = bg.dev
lib = _lib.load()
what = set(sys.argv[1:]) or {"decode", "tc"}
LAYER_WEIGHT_BYTES = 6_476_005_376 / 32 / 2 + 2 * 4 * 6_476_005_376 / 32 / 128


def layer_time(m, static=True):
    sets = {name: bg.make(k, n, max(2, min(8, (256 << 20) // (k * n // 2 * cnt) + 1)) * cnt) for name, k, n, cnt in bg.SHAPES}
    xq = torch.randn(m, 4096, device=dev).half().float()
    xd = torch.randn(m, 11008, device=dev).half().float()
    ys = {"qkvo": [torch.zeros(m, 4096, device=dev) for _ in range(4)], "gate_up": [torch.zeros(m, 11008, device=dev) for _ in range(2)],
          "down": [torch.zeros(m, 4096, device=dev)]}
    rounds = min(len(sets["qkvo"]) // 4, len(sets["gate_up"]) // 2, len(sets["down"]))

    def layer(i):
        q4, gu, dn = sets["qkvo"][4 * i:4 * i + 4], sets["gate_up"][2 * i:2 * i + 2], sets["down"][i]
        ops.gptq4_matmul_batch([(xq, w[0], y, w[1], w[2]) for w, y in zip(q4[:3], ys["qkvo"][:3])], 128, static_weights=static)
        ops.gptq4_matmul(xq, q4[3][0], ys["qkvo"][3], q4[3][1], q4[3][2], 128, static_weights=static)
        ops.gptq4_matmul_batch([(xq, w[0], y, w[1], w[2]) for w, y in zip(gu, ys["gate_up"])], 128, static_weights=static)
        ops.gptq4_matmul(xd, dn[0], ys["down"][0], dn[1], dn[2], 128, static_weights=static)

    return bg.timeit(lambda i: layer(i % rounds), 96, rounds)


if "decode" in what:
    # (mode, static_weights): 0 = plain launch; 2 = programmatic launch; 6 = + L2 prefetch of the later K blocks;
    # high nibble = resident CTAs per SM the K split aims at
    for mode, static in ((0, False), (2, False), (2, True), (6, True), (6 | (8 << 4), True), (2 | (8 << 4), True), (6 | (3 << 4), True),
                         (4, False), (3, True)):
        _lib.check(lib.sb200_gptq4_set_decode(mode))
        for m in (1, 4, 16):
            try:
                t = layer_time(m, static)
                print(json.dumps({"ab": "decode_layer", "mode": mode, "static_weights": static, "M": m, "us_per_layer": round(t * 1e6, 2),
                                  "weight_GBps": round(LAYER_WEIGHT_BYTES / t / 1e9, 1), "tok_per_s_32_layers": round(m / (t * 32), 1)}), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"ab": "decode_layer", "mode": mode, "M": m, "error": repr(e)[:200]}), flush=True)
        for name, k, n, _ in bg.SHAPES:
            wbytes = k * n // 2
            copies = max(2, min(24, (256 << 20) // wbytes + 1))
            ws = bg.make(k, n, copies)
            x = torch.randn(1, k, device=dev).half().float()
            y = torch.zeros(1, n, device=dev)
            try:
                t = bg.timeit(lambda i: ops.gptq4_matmul(x, ws[i % copies][0], y, ws[i % copies][1], ws[i % copies][2], 128, impl=1,
                                                         static_weights=static), 240, copies)
                print(json.dumps({"ab": "decode_linear", "mode": mode, "static_weights": static, "shape": name, "M": 1, "us": round(t * 1e6, 2),
                                  "weight_GBps": round((wbytes + 2 * n * (k // 128) * 4) / t / 1e9, 1)}), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"ab": "decode_linear", "mode": mode, "shape": name, "error": repr(e)[:200]}), flush=True)
            del ws
    _lib.check(lib.sb200_gptq4_set_decode(6))

if "tc" in what:
    for m in (256, 512, 2048):
        for name, k, n, _ in bg.SHAPES:
            ws = bg.make(k, n, 2)
            x = torch.randn(m, k, device=dev).half().float()
            y = torch.zeros(m, n, device=dev)
            variants = [("tc_group_x16", 2, 0), ("tc_group_x8pairs", 2, 1)] + ([("ts_c512", 3, 1)] if m >= 512 else [])
            for label, impl, narrow in variants:
                _lib.check(lib.sb200_gptq4_set_tc_drain(narrow))
                try:
                    t = bg.timeit(lambda i: ops.gptq4_matmul(x, ws[i % 2][0], y, ws[i % 2][1], ws[i % 2][2], 128, impl=impl), 20, 2)
                    print(json.dumps({"ab": "prefill", "impl": label, "shape": name, "M": m, "us": round(t * 1e6, 1),
                                      "TFLOPs": round(2.0 * m * k * n / t / 1e12, 1)}), flush=True)
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"ab": "prefill", "impl": label, "shape": name, "M": m, "error": repr(e)[:200]}), flush=True)
            del ws
    _lib.check(lib.sb200_gptq4_set_tc_drain(1))
