#!/usr/bin/env python
"""Static evidence from the SHIPPED library (no GPU needed): per kernel the registers / stack from `cuobjdump --dump-resource-usage` and a
histogram of the SASS mnemonics that matter for the claims in DESIGN.md -- tcgen05 / TMEM (UTCHMMA, LDTM, UTCBAR, SYNCS), the FP32 pipe
(FFMA / FMUL / FADD), MUFU, local-memory traffic (LDL / STL: spills) and block-wide barriers (BAR).

    python tools/sass_summary.py [control_toolkit_b200/libctk_b200.so] > profiles/sass_summary_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["FFMA", "FMUL", "FADD", "MUFU", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "SYNCS", "UTMALDG", "LDL", "STL", "BAR", "LDG", "STG", "LDS", "STS",
         "SHFL", "REDUX", "ATOMS", "ATOMG", "RED"]
# the instantiations bench.py / the GPU tests launch; everything else is summarised in one line
FOCUS = [
    r"mppi_ode_kernel<0, ?false, ?10, ?2, ?896, ?false>", r"mppi_ode_kernel<0, ?true, ?10, ?2, ?768, ?false>", r"mppi_ode_kernel<0, ?false, ?10, ?1, ?1024, ?false>",
    r"mppi_ode_kernel<0, ?false, ?0, ?1, ?1024, ?true>", r"mppi_ode_batch_kernel", r"mppi_rollout_kernel<ctk::MlpTcPred, ?0, ?false>",
    r"mppi_rollout_kernel<ctk::MlpTcFastPredT<true>, ?0, ?false>", r"mppi_rollout_kernel<ctk::MlpTcFastPredT<false>, ?0, ?false>",
    r"cem_tick_kernel<0, ?false, ?true>", r"cem_ode_kernel<0, ?false>", r"cem_refit_kernel", r"topk_level_kernel", r"rpgd_grad_coef_kernel",
    r"rpgd_select_kernel", r"env_mppi_rollout_kernel<ctk::DubinsEnv, ?false>", r"mppi_rollout_kernel<ctk::GruSimtPred, ?0, ?false>",
]


def run(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "control_toolkit_b200", "libctk_b200.so")
    res = {}
    cur = None
    for line in run("cuobjdump", "--dump-resource-usage", so).splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            res[cur] = (int(m.group(1)), int(m.group(2)))
    names = sorted(res)
    demangled = dict(zip(names, run("c++filt", *names).splitlines())) if names else {}
    hist = {}
    cur = None
    for line in run("cuobjdump", "-sass", so).splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
            hist[cur]["*"] += 1
    print(f"# {os.path.relpath(so, REPO)}: {len(res)} kernels (sm_100a).  Columns: registers / stack bytes / SASS instructions, then the mnemonics of interest.")
    tot = collections.Counter()
    for k in names:
        tot.update(hist.get(k, {}))
    print("# whole library: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
    print(f"# kernels with any LDL/STL (local-memory traffic): {sum(1 for k in names if hist.get(k, {}).get('LDL', 0) + hist.get(k, {}).get('STL', 0) > 0)} of {len(names)}")
    print()
    shown = set()
    for pat in FOCUS:
        for k in names:
            d = demangled.get(k, k)
            if re.search(pat, d) and k not in shown:
                shown.add(k)
                h = hist.get(k, collections.Counter())
                short = re.sub(r"\(.*$", "", d.replace("void ", "").replace("ctk::", ""))
                print(f"{short}\n    REG {res[k][0]}  STACK {res[k][1]}  instr {h['*']}   " + " ".join(f"{w}={h[w]}" for w in WATCH if h[w]))
    print()
    print(f"# the other {len(names) - len(shown)} instantiations: max REG {max(res[k][0] for k in names if k not in shown)}, "
          f"with LDL/STL: " + ", ".join(sorted({re.sub(r'<.*$', '', demangled[k].replace('void ', '').replace('ctk::', '')) for k in names if k not in shown and hist.get(k, {}).get('LDL', 0) + hist.get(k, {}).get('STL', 0) > 0})))


if __name__ == "__main__":
    main()
