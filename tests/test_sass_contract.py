"""CPU: what the BUILT library contains, read with cuobjdump (no GPU): the hot-path kernels keep their state in
registers (no local-memory stack), the default chain kernels move their tiles on the bulk-copy engine, and the fused
Dense(P)+chain kernel really issues tcgen05 MMAs against tensor memory.  A regression here (a spill after a code
change, a silently dropped TMA path) would not fail any numerics test -- it would only show up as a slower bench."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "normalizingflownetwork_b200", "libnfn_b200.so")

CFG2 = "ChainSpec<2, true, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0>"
CFG4 = "ChainSpec<1, true, 1, 1, 1, 1, 1>"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None
                                or not os.path.exists(LIB), reason="needs cuobjdump, c++filt and the built library")


def _run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=300).stdout


@pytest.fixture(scope="module")
def resources():
    """{demangled kernel name: (mangled name, registers, stack bytes)} from `cuobjdump -res-usage`."""
    lines = _run("cuobjdump", "-res-usage", LIB).splitlines()
    mangled, usage = [], []
    for i, line in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            u = re.search(r"REG:(\d+) STACK:(\d+)", lines[i + 1])
            mangled.append(m.group(1))
            usage.append((int(u.group(1)), int(u.group(2))))
    names = _run("c++filt", *mangled).splitlines()
    assert len(names) == len(mangled) > 100
    return {d: (m,) + u for d, m, u in zip(names, mangled, usage)}


def _sass_ops(mangled):
    out = _run("cuobjdump", "-sass", "-fun", mangled, LIB)
    return re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", out, flags=re.M)


def _find(resources, *needles):
    hits = [d for d in resources if all(n in d for n in needles)]
    assert hits, "no kernel matching %r in the library" % (needles,)
    return hits[0]


def test_library_is_sm100a_only():
    archs = set(re.findall(r"arch = (\S+)", _run("cuobjdump", "-lelf", "-lptx", LIB) + _run("cuobjdump", "-res-usage", LIB)))
    assert archs == {"sm_100a"}, archs


def test_hot_path_kernels_do_not_spill(resources):
    families = ("nfn::chain_kernel_w<", "nfn::chain_kernel<", "nfn::tc5::dense_tc5_kernel<", "nfn::dense_chain_kernel<",
                "nfn::dense_mdn_kernel<", "nfn::dense_kmn_kernel<", "nfn::kmn_kernel<", "nfn::colsum_kernel",
                "nfn::peer_allreduce_kernel", "nfn::logmeanexp_kernel", "variational_fwd", "variational_bwd",
                "dense_act_bwd_mma<16,")
    seen = {f: 0 for f in families}
    for name, (_, regs, stack) in resources.items():
        for f in families:
            if f in name:
                seen[f] += 1
                assert stack == 0, "%s uses %d bytes of local stack (spill)" % (name, stack)
                assert regs <= 255
    assert all(seen.values()), seen
    # the streaming MDN head at the BASELINE config-5 shape (d = 2)
    _, regs, stack = resources[_find(resources, "nfn::mdn_kernel<2, true, 4, true, nfn::MathFast>")]
    assert stack == 0 and regs <= 128   # 4 CTAs of 128 threads per SM by its launch bounds


def test_default_chain_kernels_use_the_bulk_copy_engine(resources):
    # P = 48 (cfg2): 2-D tensor maps with hardware swizzle in, tensor-map stores out; nothing through LDGSTS
    ops = _sass_ops(resources[_find(resources, "nfn::chain_kernel_w<", CFG2, ">, true, nfn::MathFast")][0])
    assert sum(o.startswith("UTMALDG") for o in ops) >= 1 and sum(o.startswith("UTMASTG") for o in ops) >= 1
    assert not any(o.startswith("LDGSTS") for o in ops)
    assert any(o.startswith("SYNCS.PHASECHK") for o in ops)      # mbarrier complete_tx wait, no CTA barrier per tile
    # P = 17 (cfg4) has no swizzle mode: linear bulk copies both ways
    ops = _sass_ops(resources[_find(resources, "nfn::chain_kernel_w<", CFG4, ">, true, nfn::MathFast")][0])
    assert "UBLKCP.S.G" in ops and "UBLKCP.G.S" in ops and not any(o.startswith("LDGSTS") for o in ops)
    # the round-1 generation stays available (opt-in) and is the cp.async one
    ops = _sass_ops(resources[_find(resources, "nfn::chain_kernel<", CFG2, ">, true, nfn::MathFast")][0])
    assert any(o.startswith("LDGSTS") for o in ops) and not any(o.startswith("UTMA") for o in ops)
    # no chain kernel touches the tensor cores: the op is elementwise per row and HBM-bound
    assert not any("MMA" in o for o in ops)


def test_fused_dense_chain_kernel_is_tcgen05(resources):
    name = _find(resources, "nfn::tc5::dense_tc5_kernel<", CFG2, "16, true, nfn::MathFast")
    ops = _sass_ops(resources[name][0])
    assert sum(o.startswith("UTCHMMA") for o in ops) >= 20          # tcgen05.mma kind::f16 (bf16 levels)
    assert sum(o.startswith("LDTM") for o in ops) >= 10             # tcgen05.ld of the TMEM accumulators
    assert any(o.startswith("UTCBAR") for o in ops)                 # tcgen05.commit -> mbarrier
    assert sum(o.startswith("UTCATOMSWS") for o in ops) >= 2        # tcgen05.alloc / dealloc
    assert sum(o.startswith("USETMAXREG") for o in ops) == 2        # setmaxnreg: compute vs issuing warpgroup
    assert "FADD2" in ops                                           # packed-pair bf16 level splits
    assert not any(o.startswith("HMMA") for o in ops)               # no mma.sync fallback inside this kernel
    # the mma.sync sibling (small batches / shapes tcgen05 does not take) is 3xTF32
    ops = _sass_ops(resources[_find(resources, "nfn::dense_chain_kernel<", CFG2, "16, true, nfn::MathFast")][0])
    assert "HMMA.1688.F32.TF32" in ops and not any(o.startswith("UTC") for o in ops)
