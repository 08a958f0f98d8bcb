"""Same-box A/B of a kernel change: the product library vs an experimental build of the same sources.

Here (no GPU):    python tools/ab.py build x2 -DMMFB_F32X2=1        -> mmf_b200/csrc/libmmfb200_x2.so (travels with gpurun)
On the GPU box:   python tools/ab.py run x2 [--tests] [--steps 12]   (one gpurun call)
                  python tools/ab.py run lnlean --env MMFB_LN_BWD=lean [--tests]     (run-time switch, product library;
                  --env may be repeated and combined with a variant library of that name if one was built)
    1. (--tests) the whole `-m gpu` suite with MMFB_LIB pointing at the variant: parity first
    2. bench.py with the product library, then with the variant, back to back on the same box / clocks
    3. one line per arm + the ratio; the two bench JSON lines are kept in gpurun_out/ab_<name>_{base,variant}.json
                  python tools/ab.py sweep x2 lnlean:MMFB_LN_BWD=lean rng:MMFB_DROPOUT_ASYNC=1 all:MMFB_LN_BWD=lean,MMFB_DROPOUT_ASYNC=1
                  (the product library once, then every NAME[:ENV=VAL,...] in turn; a variant library libmmfb200_NAME.so
                  is used when it exists; one table at the end, JSON lines in gpurun_out/ab_sweep_*.json)
MMFB_LIB only selects which build of the SAME C ABI is loaded; there is no fallback involved.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib_path(name):
    return os.path.join(ROOT, "mmf_b200", "csrc", "libmmfb200_%s.so" % name)


def bench(env, steps, warmup, out):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", str(warmup),
           "--no-cpu-baseline", "--no-parity"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        sys.stderr.write(r.stderr[-2000:])
        raise SystemExit("bench failed (%s)" % out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", out), "w") as fh:
        fh.write(lines[-1] + "\n")
    return json.loads(lines[-1])


def sweep(specs, steps):
    env_b = {k: v for k, v in os.environ.items() if k != "MMFB_LIB"}
    rows = [("base", bench(env_b, steps, 4, "ab_sweep_base.json"))]
    for spec in specs:
        name, _, envs = spec.partition(":")
        env = dict(env_b)
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=", 1)
            env[k] = v
        if os.path.exists(lib_path(name)):
            env["MMFB_LIB"] = lib_path(name)
        try:
            rows.append((spec, bench(env, steps, 4, "ab_sweep_%s.json" % name)))
        except SystemExit as e:
            print("%-40s FAILED (%s)" % (spec, e))
    base = rows[0][1]["value"]
    for tag, d in rows:
        print("%-40s %8.1f samples/s  %7.3f ms/step  x%.4f  sm_mhz %s" % (
            tag, d["value"], d["ms_per_step"], d["value"] / base, (d.get("clocks") or {}).get("sm_mhz")))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "sweep":
        specs = [a for a in sys.argv[2:] if not a.startswith("--") and not a.isdigit()]
        steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 12
        return sweep(specs, steps)
    if len(sys.argv) < 3 or sys.argv[1] not in ("build", "run"):
        raise SystemExit(__doc__)
    mode, name = sys.argv[1], sys.argv[2]
    if mode == "build":
        sys.path.insert(0, ROOT)
        from mmf_b200.csrc.build import build
        print(build(variant=name, defines=[a for a in sys.argv[3:] if a.startswith("-D")]))
        return
    variant = lib_path(name)
    extra = {}
    for i, a in enumerate(sys.argv):
        if a == "--env":
            k, v = sys.argv[i + 1].split("=", 1)
            extra[k] = v
    if not os.path.exists(variant) and not extra:
        raise SystemExit("%s is missing: run `python tools/ab.py build %s -D...` before gpurun" % (variant, name))
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 12
    env_v = dict(os.environ, **extra)
    if os.path.exists(variant):
        env_v["MMFB_LIB"] = variant
    env_b = {k: v for k, v in os.environ.items() if k != "MMFB_LIB" and k not in extra}
    if "--tests" in sys.argv:
        r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q", "-x"], env=env_v, cwd=ROOT)
        if r.returncode != 0:
            raise SystemExit("variant %s fails the GPU suite: not measured" % name)
    base = bench(env_b, steps, 4, "ab_%s_base.json" % name)
    var = bench(env_v, steps, 4, "ab_%s_variant.json" % name)
    for tag, d in (("base", base), (name, var)):
        print("%-8s %8.1f samples/s  %7.3f ms/step  e2e %8.1f  dominant %s  sm_mhz %s" % (
            tag, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("achieved"),
            (d.get("clocks") or {}).get("sm_mhz")))
    print("ratio %s/base = %.4f" % (name, var["value"] / base["value"]))


if __name__ == "__main__":
    main()
