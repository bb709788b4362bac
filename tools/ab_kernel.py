"""GPU diagnostic: A/B of two builds of libls_b200.so on the same box, alternating runs of bench.py: kernel-only time
of a 16-step launch at B=512 (TED) and the device-resident value.  usage: python tools/ab_kernel.py libA.so libB.so [rounds]"""
import json
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
libs = sys.argv[1:3]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
res = {lib: [] for lib in libs}
for r in range(rounds):
    for lib in libs:
        env = dict(os.environ, LS_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--warmup", "5",
                              "--no-cpu-baseline"] + (["--min-seconds", os.environ["AB_MIN_SECONDS"]] if "AB_MIN_SECONDS" in os.environ else []), env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        res[lib].append((d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_min"], d["value"], d["clocks"]["sm_mhz"],
                         d["e2e"]["value"], d["clocks"].get("power_w")))
        print(os.path.basename(lib), "kernel_ms %.3f min %.3f value %.1f clk %s e2e %.1f power %s W" % res[lib][-1], flush=True)
for lib in libs:
    print(os.path.basename(lib), "mean kernel_ms %.3f  mean value %.1f  mean e2e %.1f" % (
        sum(x[0] for x in res[lib]) / rounds, sum(x[2] for x in res[lib]) / rounds, sum(x[4] for x in res[lib]) / rounds))
