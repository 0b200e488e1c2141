"""A/B timing of several builds of libmhla_b200.so in ONE process (GPU box only): every library is loaded through the
normal ctypes binding (`_capi.lib()` re-pointed at it), the same inputs are used for all of them, results are checked
bitwise against the first library, and the order is rotated over the repetitions so that clock / thermal drift does not
favour one build.

    python tools/ab_libs.py [--reps R] [--shapes headline,nonorm,wan,n8k] name=path.so name2=path2.so ...
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

args = [a for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
shape_sel = sys.argv[sys.argv.index("--shapes") + 1].split(",") if "--shapes" in sys.argv else ["headline", "nonorm", "wan"]
# name=path.so[@ENV=VAL[,ENV=VAL]]: the environment is set while the library makes its FIRST blockmix call (that is when it
# reads its tuning knobs); the same build under two knob settings needs two copies of the file (one dlopen handle per path)
libs, envs = [], {}
for a in args:
    name, rest = a.split("=", 1)
    path, _, env = rest.partition("@")
    libs.append((name, os.path.abspath(path)))
    envs[name] = dict(kv.split("=", 1) for kv in env.split(",")) if env else {}
SHAPES = {
    "headline": (2, 16, 128, 256, 64, True, False),
    "nonorm": (2, 16, 128, 256, 64, False, False),
    "wan": (1, 12, 150, 210, 128, False, True),
    "wan_norm": (2, 12, 150, 210, 128, True, True),
    "n8k": (2, 16, 32, 256, 64, True, False),
    "n128k": (2, 16, 512, 256, 64, True, False),
}
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def use(path):
    _capi._lib = None
    _capi.LIB_PATH = path
    return _capi.lib()


def timed(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


handles = {}
for name, path in libs:
    handles[name] = use(path)   # keep every CDLL alive; `use` only re-points the binding
    os.environ.update(envs[name])
    t = torch.randn(1, 1, 4, 128, 64, device=dev).bfloat16().relu() + 1e-3      # general kernel (N = 512): reads the knobs
    mhla_b200.mhla(t, t, t, torch.rand(4, 4, device=dev) / 4, normalize=True)
    torch.cuda.synchronize()
    for k_ in envs[name]:
        del os.environ[k_]
results = {}
for sname in shape_sel:
    B, H, M, w, D, normalize, rope = SHAPES[sname]
    mk = lambda: torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()  # noqa: E731
    q, k, v = mk().relu() + 1e-6, mk().relu() + 1e-6, mk()
    qr, kr = (mk(), mk()) if rope else (None, None)
    W = torch.rand(M, M, device=dev) / M
    ref = None
    times = {name: [] for name, _ in libs}
    for r in range(reps):
        order = libs[r % len(libs):] + libs[:r % len(libs)]
        for name, path in order:
            _capi._lib = handles[name]
            out = torch.empty_like(q)
            fn = lambda: mhla_b200.mhla(q, k, v, W, q_rope=qr, k_rope=kr, normalize=normalize, out=out)  # noqa: E731
            t = timed(fn)
            times[name].append(round(t, 2))
            if ref is None:
                ref = out.clone()
            elif not torch.equal(ref, out):
                print(f"!! {sname}: {name} differs from {libs[0][0]}: max abs {float((ref.float() - out.float()).abs().max()):.3e}")
    nbytes = 4 * q.numel() * 2
    for name, _ in libs:
        best = min(times[name])
        results[(sname, name)] = best
        print(json.dumps({"shape": sname, "lib": name, "us": times[name], "best_us": best,
                          "frac_of_6547": round(nbytes / (best * 1e-6) / 6547.2e9, 4)}), flush=True)
