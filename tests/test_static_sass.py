"""CPU checks of the SHIPPED library's machine code (cuobjdump; no GPU): properties the measured numbers in DESIGN.md depend on and
that a source change can silently lose -- every K1 instantiation fits its launch bound in the register file (a 73rd register at 896
threads would make the launch fail on the GPU box, where nothing compiles), the production instantiations do not spill inside the
rollout loops, and the MLP engines are tcgen05 / TMEM code (UTCHMMA, LDTM, UTCBAR), not a recompiled mma.sync path."""
import os
import re
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "control_toolkit_b200", "libctk_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")


def _resources():
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            res[cur] = tuple(int(x) for x in m.groups())
    return res


def _sass(mangled):
    return subprocess.run([CUOBJDUMP, "-sass", "-fun", mangled, LIB], capture_output=True, text=True).stdout


def _mnemonics(sass):
    return [m.group(1) for m in (re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln) for ln in sass.splitlines()) if m]


def test_library_is_sm_100a_and_every_k1_instantiation_fits_its_launch_bound():
    elf = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf  # one architecture, no multi-arch fat binary
    res = _resources()
    k1 = {k: v for k, v in res.items() if "15mppi_ode_kernelI" in k}
    assert len(k1) >= 12
    for name, (reg, _stack, _sh) in k1.items():
        # mppi_ode_kernel<KIND, LOG, PERIOD, ILP, MAXT, INJ>: MAXT is the launch bound of the instantiation (ctk_kernels_mppi_ode.cuh)
        maxt = int(re.search(r"ILi\d+ELb[01]ELi\d+ELi\d+ELi(\d+)ELb[01]EE", name).group(1))
        assert reg * maxt <= 65536, (name, reg, maxt)
    # the instantiation bench.py times: two rollouts per thread, 896 threads, 72 registers -- one block per SM, no register spill
    prod = "_ZN3ctk15mppi_ode_kernelILi0ELb0ELi10ELi2ELi896ELb0EEEvNS_11MppiOdeArgsE"
    assert prod in res and res[prod][0] <= 72


def test_production_k1_has_no_local_memory_traffic_inside_the_rollout_loops():
    """The only LDL / STL of the production instantiation belong to the slow path of the one cosf() of the prologue; they sit after
    the last backward branch of the rollout loops (the register-spill regression of round 2 put them inside: +3 % tick time)."""
    sass = _sass("_ZN3ctk15mppi_ode_kernelILi0ELb0ELi10ELi2ELi896ELb0EEEvNS_11MppiOdeArgsE")
    lines = [ln for ln in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln)]
    assert len(lines) > 5000
    addr = lambda ln: int(re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln).group(1), 16)  # noqa: E731
    local = [addr(ln) for ln in lines if re.search(r"\b(LDL|STL)\b", ln)]
    assert len(local) <= 8, len(local)
    # loops = backward branches; the hot ones span several hundred instructions (an unrolled ten-step segment pair)
    loops = []
    for ln in lines:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", ln)
        if m and int(m.group(1), 16) < addr(ln):
            loops.append((int(m.group(1), 16), addr(ln)))
    hot = [(a, b) for a, b in loops if b - a > 0x400]
    assert hot, loops
    for x in local:
        assert not any(a <= x <= b for a, b in hot), (hex(x), [(hex(a), hex(b)) for a, b in hot])


@pytest.mark.parametrize("pred,min_mma", [("NS_9MlpTcPredE", 12), ("NS_14MlpTcFastPredTILb1EEE", 8), ("NS_14MlpTcFastPredTILb0EEE", 8)])
def test_mlp_engines_are_tcgen05_code(pred, min_mma):
    name = f"_ZN3ctk19mppi_rollout_kernelI{pred}Li0ELb0EEEvNS_8MppiArgsE"
    ops = _mnemonics(_sass(name))
    assert len(ops) > 1000, name
    assert ops.count("UTCHMMA") >= min_mma and "LDTM" in ops and "UTCBAR" in ops and "SYNCS" in ops, {k: ops.count(k) for k in ("UTCHMMA", "LDTM", "UTCBAR", "SYNCS")}
    assert not any(o.startswith(("HMMA", "IMMA", "HGMMA")) for o in ops)  # no mma.sync / wgmma path
