"""Width sweep of the MLP surrogate: Dense(5 -> nh) + LeakyReLU + Dense(nh -> 4) at nh = 10, 64, 256 on 2^24
samples, fp32 FMA path (ponni's operation order) against the tcgen05 path; one JSON line per (width, path).  Also the
fused fp64-field form of the shipped 5 -> 10 -> 4 network at config-2 size."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miniweatherml_b200 as mw

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B = int(os.environ.get("MW_SWEEP_B", str(1 << 24)))
rng = np.random.default_rng(5)
x = torch.tensor(rng.uniform(-1, 1, (5, B)).astype(np.float32), device="cuda")
for nh in (10, 64, 256):
    w = rng.uniform(-0.5, 0.5, 5 * nh + nh + nh * 4 + 4).astype(np.float32)
    ref = None
    for tc in (False, True):
        if not tc and nh == 256 and B > (1 << 22):
            xs = x[:, : 1 << 22].contiguous()            # the fp32 path spills its 256 activations: time a quarter
            ms = timed(lambda: mw.mlp_dense2_forward(w, xs, nh, 4, 0.1, use_tensor_cores=False), reps=2) * (B / (1 << 22))
            y = None
        else:
            ms = timed(lambda: mw.mlp_dense2_forward(w, x, nh, 4, 0.1, use_tensor_cores=tc))
            y = mw.mlp_dense2_forward(w, x, nh, 4, 0.1, use_tensor_cores=tc)
        err = None
        if tc and ref is not None: err = float((y - ref).abs().max())
        if not tc: ref = y
        flops = 2.0 * B * (5 * nh + nh * 4)
        print(json.dumps({"probe": "mlp_width_sweep", "nin": 5, "nh": nh, "nout": 4, "B": B, "path": "tcgen05 3xTF32" if tc else "fp32 FMA",
                          "ms": ms, "samples_per_s": B / ms * 1e3, "model_GFLOPs": flops / ms / 1e6, "alg_GBps": B * 36 / ms / 1e6,
                          "max_abs_diff_vs_fp32": err}), flush=True)
