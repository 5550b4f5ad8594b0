"""Build A/B variants of libqtrack.so (one per -D setting) into lqr_quadcopter_test_b200/variants/
and, on the GPU box, time the headline rollout with each:

    python tools/ab_variants.py build  name1:-DQT_X=0  name2:-DQT_Y=1,-DQT_Z=0 ...
    python tools/ab_variants.py build  only=qt_dare.cu  name1:-DQT_X=0 ...   # recompile one source, link the default objects of the rest
    python tools/ab_variants.py run [E]      # every variants/libqtrack_*.so + the default library
    python tools/ab_variants.py run_dare [batch]
"""
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
VAR = ROOT / "lqr_quadcopter_test_b200" / "variants"
sys.path.insert(0, str(ROOT))


def build(specs):
    from lqr_quadcopter_test_b200 import build as qb

    only = None
    if specs and specs[0].startswith("only="):
        only = specs[0][5:].split(",")
        specs = specs[1:]
        qb.build()                       # default objects of the other sources
    VAR.mkdir(exist_ok=True)
    for old in VAR.glob("libqtrack_*.so"):
        old.unlink()
    nvcc = qb._nvcc()
    for spec in specs:
        name, flags = spec.split(":", 1)
        flags = [f for f in flags.split(",") if f]
        objs = []
        tmp = VAR / f"obj_{name}"
        tmp.mkdir(exist_ok=True)
        procs = []
        mine = []
        for src in qb.SOURCES:
            if only is not None and src not in only:
                objs.append(str(qb.PKG / "build" / (src + ".o")))
                continue
            o = tmp / (src + ".o")
            objs.append(str(o))
            mine.append(str(o))
            procs.append(subprocess.Popen([nvcc, *qb.NVCC_FLAGS, *flags, "-c", str(qb.CSRC / src), "-o", str(o)]))
        assert all(p.wait() == 0 for p in procs), name
        subprocess.run([nvcc, "-shared", "-o", str(VAR / f"libqtrack_{name}.so"), *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], check=True)
        for o in mine:
            os.unlink(o)
        tmp.rmdir()
        print("built", name, flags, flush=True)


def run(E):
    libs = [None] + sorted(VAR.glob("libqtrack_*.so"))
    for lib in libs:
        env = dict(os.environ)
        if lib is not None:
            env["QTRACK_LIB"] = str(lib)
        r = subprocess.run([sys.executable, __file__, MODE, str(E)], env=env, capture_output=True, text=True)
        print(f"{'default' if lib is None else lib.stem[10:]:24s} {r.stdout.strip() or r.stderr[-400:]}", flush=True)


def one_policy(E):
    import torch

    from lqr_quadcopter_test_b200 import DeepTrackingPolicy, EnvConfig, SuccessCriteria, _lib, rollout_device

    ctx = _lib.get_context()
    seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
    res = []
    for hidden in ([64, 64], [32], [128, 128]):
        torch.manual_seed(0)
        pol = DeepTrackingPolicy(config={"hidden_sizes": hidden})
        env = EnvConfig()
        env.target.motion_type = "linear"
        for _ in range(2):
            rollout_device(pol, env, SuccessCriteria(), seeds)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record()
        for i in range(3):
            r = rollout_device(pol, env, SuccessCriteria(), seeds)
            evs[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(3))[1]
        steps = float(r["metrics"][0].sum().item())
        res.append(f"{hidden}: {steps / ms * 1e3:.3e} steps/s ({ms:.0f} ms)")
    print("   ".join(res))


def one(E):
    import torch

    from lqr_quadcopter_test_b200 import EnvConfig, RiccatiLQRController, SuccessCriteria, _lib, rollout_device

    ctx = _lib.get_context()
    seeds = torch.arange(E, dtype=torch.int64, device=ctx.device) + 42
    res = []
    for motion, cfg in (("linear", {}), ("circular", {"use_lqi": True, "q_int": [0.01, 0.01, 0.1]})):
        env = EnvConfig()
        env.target.motion_type = motion
        ctrl = RiccatiLQRController(config=cfg)
        for _ in range(3):
            rollout_device(ctrl, env, SuccessCriteria(), seeds)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        evs[0].record()
        for i in range(8):
            r = rollout_device(ctrl, env, SuccessCriteria(), seeds)
            evs[i + 1].record()
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(8)]
        ms = sorted(per)[len(per) // 2]
        spread = f"{min(per):.1f}-{max(per):.1f}"
        steps = float(r["metrics"][0].sum().item())
        import numpy as np
        res.append(f"{motion}{'_lqi' if cfg else ''}: {steps / ms * 1e3:.4e} steps/s (median {ms:.1f} ms of {spread}, nnz(K)={int(np.count_nonzero(ctrl.effective_gain()))})")
    print("   ".join(res))


def one_dare(batch):
    extra = [os.environ["QT_DARE_SIZES"]] if os.environ.get("QT_DARE_SIZES") else []
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_dare.py"), str(batch), "brief", *extra], capture_output=True, text=True)
    print(r.stdout.strip().replace("\n", "   ") or r.stderr[-400:])


MODE = "one"

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20)
    elif sys.argv[1] == "run_policy":
        MODE = "one_policy"
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18)
    elif sys.argv[1] == "run_dare":
        MODE = "one_dare"
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
    elif sys.argv[1] == "one_dare":
        one_dare(int(sys.argv[2]))
    elif sys.argv[1] == "one_policy":
        one_policy(int(sys.argv[2]))
    else:
        one(int(sys.argv[2]))
