"""A/B tuning harness for the fused step kernel.

    python profiles/tools/ab_variants.py build  name1:-DST_X=1,-DST_Y=2  name2:...     (here: nvcc cross-compiles into tests/_build/ab/)
    python profiles/tools/ab_variants.py run [--envs 65536] [--steps 1500]              (on the GPU box: every built variant, interleaved)

`run` executes bench.py (kernel-only legs) once per variant and round, interleaving the variants so that box-to-box and thermal drift
hit all of them alike, and prints kernel_ms / ms_per_step per variant (median over rounds).
"""
import glob
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
AB = os.path.join(ROOT, "puffer_phc_b200", "lib", "ab")
sys.path.insert(0, ROOT)


def main():
    if sys.argv[1] == "build":
        from puffer_phc_b200 import build as b
        for spec in sys.argv[2:]:
            name, _, flags = spec.partition(":")
            out = os.path.join(AB, f"libphc_{name}.so")
            b.build(out=out, extra_flags=[f for f in flags.split(",") if f])
            print("built", out)
        return
    envs, steps, rounds = "65536", "1500", 3
    args = sys.argv[2:]
    for i, a in enumerate(args):
        if a == "--envs": envs = args[i + 1]
        if a == "--steps": steps = args[i + 1]
        if a == "--rounds": rounds = int(args[i + 1])
    libs = [l for l in sorted(glob.glob(os.path.join(AB, "libphc_*.so"))) if not os.path.basename(l).startswith("libphc_p_")]   # p_ = profile builds
    res = {os.path.basename(l): [] for l in libs}
    for r in range(rounds):
        for l in libs:
            env = dict(os.environ, PHC_B200_LIB=l)
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--envs", envs, "--steps", steps, "--warmup", "20", "--no-e2e",
                                "--no-cpu-baseline", "--no-other-configs"], env=env, capture_output=True, text=True)
            try:
                j = json.loads(p.stdout.strip().splitlines()[-1])
                res[os.path.basename(l)].append((j["roofline"]["kernel_ms"], j["ms_per_step"], j["clocks"]["sm_mhz"]))
            except Exception:
                res[os.path.basename(l)].append((float("nan"), float("nan"), p.stderr[-300:]))
    for k, v in res.items():
        km = statistics.median(x[0] for x in v)
        print(f"{k:40s} kernel_ms median {km:.4f}  all {[round(x[0], 4) for x in v]}  step {[round(x[1], 4) for x in v]}  clk {[x[2] for x in v]}")


if __name__ == "__main__":
    main()
