"""Kernel-only timing of both first-layer variants of the fused rollout at bench size (CUDA events).

    python profiles/time_variants.py [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n, T = 1 << 20, 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ("cuda", "cuda+direct", "pairs", "pairs+direct", "sorted", "tcgen05", "tcgen05_ws")
tune = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
for variant in variants:
    direct = variant.endswith("+direct")
    sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T,
                            variant=variant.split("+")[0], direct_rings=direct)
    if tune:
        from nfsp_b200._lib import lib
        lib().nfsp_rollout_tune(sp.env._h, tune[0], tune[1])
    ms = []
    for k in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sp.rollout(T, insert=False)
        b.record()
        b.synchronize()
        sp.counts.zero_()
        if k >= 3:
            ms.append(a.elapsed_time(b))
    ms.sort()
    err = __import__("ctypes").c_uint32(0)
    from nfsp_b200._lib import lib as _l
    _l().nfsp_env_kernel_error(sp.env._h, __import__("ctypes").byref(err))
    print("%s: kernel_error %d  median %.4f ms  min %.4f  -> %.3e decisions/s" % (variant, err.value, ms[len(ms) // 2], ms[0], n * T / (ms[len(ms) // 2] * 1e-3)))
