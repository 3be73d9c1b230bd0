"""Compile-level properties of the shipped kernels that DESIGN.md states, read from the built libfjgpu.so with cuobjdump (no GPU):
the closest-hit kernel issues packed FMAs and 256-bit node loads, keeps the warp provably converged (no divergence guards around
its votes and shuffles), fits the register budget of 7 CTAs per SM without spills in the node loop; the staged tree top of
k_extend2 is a bulk copy (TMA, non-tensor form)."""
import os
import re
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "fujiyama-renderer_b200", "csrc", "libfjgpu.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

RING = "_ZN2fj13k_extend_ringILi7ELi12ELb1ELb0EEEvNS_10RenderArgsE"          # the default: 7 CTAs / SM, 12 stack entries, FFMA2, direct refill
RING_WITH_RING = "_ZN2fj13k_extend_ringILi7ELi12ELb1ELb1EEEvNS_10RenderArgsE"
EXTEND2_TOP = "_ZN2fj9k_extend2ILi7ELb1ELb1ELb1ELi12ELb1EEEvNS_10RenderArgsE"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="libfjgpu.so or cuobjdump missing")


def _sass(fun):
    out = subprocess.run([CUOBJDUMP, "-sass", "-fun", fun, LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    lines = [l for l in out.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not re.match(r"\s*/\* 0x", l)]
    assert len(lines) > 500, "function %s not found in %s" % (fun, LIB)
    return lines


def _ops(lines):
    """Mnemonic of every instruction (predicate guard stripped)."""
    out = []
    for l in lines:
        w = re.sub(r"/\*[0-9a-f]+\*/", "", l).split()
        out.append(w[1] if w[0].startswith("@") else w[0])
    return out


def _usage(fun):
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
    for i, l in enumerate(out):
        if fun + ":" in l:
            return {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", out[i + 1])}
    raise AssertionError("no resource usage for " + fun)


def test_default_closest_hit_kernel_node_step():
    ops = _ops(_sass(RING))
    assert sum(o.startswith("FFMA2") for o in ops) == 12              # the 24 plane distances of a 4-wide step
    assert sum(o.startswith("LDG.E.ENL2.256") for o in ops) >= 2       # the node: two 32-byte loads
    assert sum(o.startswith("I2F.U8") for o in ops) == 24              # 8-bit planes
    assert sum(o.startswith("STG.E") and ".256" in o for o in ops) >= 1     # the 32-byte hit record in one store


def test_warp_stays_provably_converged():
    """BRA.DIV / extra WARPSYNCs appear as soon as the compiler cannot prove convergence at a vote or shuffle (it happened when
    the loop's exit test was not its first statement): each costs two instructions per vote inside the node loop."""
    for fun in (RING, RING_WITH_RING):
        s = _sass(fun)
        assert not any("BRA.DIV" in l for l in s), fun
        assert sum("WARPSYNC" in l for l in s) <= 2, fun


def test_register_and_shared_memory_budget():
    u = _usage(RING)
    assert u["REG"] <= 72                                              # 65536 / (7 CTAs x 128 threads)
    assert u["SHARED"] <= 19 * 1024 + 512                              # 7 CTAs inside a 140 KB carveout
    # local memory holds the rarely used deep part of the traversal stack (and nothing that is touched in the node loop)
    assert u["STACK"] <= (96 + 1 - 12) * 4 + 16


def test_staged_tree_top_is_a_bulk_copy():
    s = _sass(EXTEND2_TOP)
    assert any("UBLKCP" in l for l in s)                               # cp.async.bulk.shared::cluster.global (TMA, non-tensor)
    assert any("SYNCS" in l for l in s)                                # mbarrier
