"""CPU only, build container only (needs /root/reference): time the reference's UNMODIFIED optimizer files (through the test-only
shims of oracle/refharness) next to the oracle PORT that bench.py times on the GPU box, same host, same threads, same configs --
the evidence that the port is a faithful timing proxy of the reference.  Writes profiles/reference_vs_port_r02.json.
    python tools/reference_vs_port.py"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("mppi_ode_c1", 2000, 10), ("cem_ode_c2", 4096, 5), ("rpgd_ode_c3", 32, 5), ("mppi_mlp_c4", 65536, 2), ("mppi_ode_1m", 250_000, 2), ("mppi_ode_1m", 62_500, 3)]
SNIPPET = """
import json, sys
sys.argv = ['bench.py']
import bench
w, n, k, kind = {w!r}, {n}, {k}, {kind!r}
f = bench.cpu_reference_rate if kind == 'reference' else bench.cpu_mppi_rate
rate, sec, threads = f(w, n, k, 1)
bench._emit(dict(workload=w, rollouts=n, ticks=k, kind=kind, rollout_steps_per_s=rate, s_per_tick=sec, threads=threads))
"""


def main():
    out = []
    for w, n, k in CASES:
        row = {}
        for kind in ("reference", "port"):
            r = subprocess.run([sys.executable, "-c", SNIPPET.format(w=w, n=n, k=k, kind=kind)], cwd=REPO, capture_output=True, text=True, timeout=3600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if not line:
                raise RuntimeError(r.stderr[-2000:])
            row[kind] = json.loads(line[-1])
        row["port_over_reference"] = row["port"]["rollout_steps_per_s"] / row["reference"]["rollout_steps_per_s"]
        print(w, n, {k2: round(v["rollout_steps_per_s"]) for k2, v in row.items() if isinstance(v, dict)}, "ratio %.3f" % row["port_over_reference"], flush=True)
        out.append(row)
    with open(os.path.join(REPO, "profiles", "reference_vs_port_r02.json"), "w") as f:
        json.dump({"host_threads": os.cpu_count(), "note": "build container (no GPU); torch-CPU fp32; reference = unmodified /root/reference optimizer files "
                   "through oracle/refharness, port = oracle/{mppi,cem,rpgd}.py as timed by bench.py on the GPU box", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
