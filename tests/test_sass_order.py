"""Pins the load ordering the fused and half-step-B kernels depend on (fused.cuh) in the SHIPPED
binary: the self loads (``LDG.E.128``) are issued first, the lane shuffle consumes one of them, a
predicate-carrying barrier (``BAR.RED.OR``; float kernels: a named barrier that only the warp's own 32 threads join)
follows, and only behind it come the neighbour loads.  Issued together with the self loads, every neighbour line would be fetched from HBM a second time
(neither L1 nor L2 merges a miss into a fill in flight): DRAM reads x2, the fused pass slower than two passes.  Round 1
guarded this only with a relative timing test on the GPU; a future ptxas that hoists the loads now fails here, on CPU.

Reads the SASS with cuobjdump (CUDA toolkit); skipped where the tool is absent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cytvdn_b200", "libcytvdn_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def sass_of(mangled_fragment):
    """[(address, instruction)] of the first function whose mangled name contains the fragment."""
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout.split("\n")
    start = [i for i, l in enumerate(out) if "Function :" in l and mangled_fragment in l]
    assert start, f"kernel {mangled_fragment} not found in {LIB}"
    ins = []
    for l in out[start[0] + 1:]:
        if "Function :" in l:
            break
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


VEC_LD = re.compile(r"(@!?U?P\d+\s+)?LDG\.E(\.EF)?\.128(\.CONSTANT)?\s")       # vector loads of the sweep (not the
VEC_ST = re.compile(r"(@!?U?P\d+\s+)?STG\.E(\.EF)?\.128\s")                           # .STRONG.GPU loads of the reduction)


def check_two_phase(ins, n_self, n_nbr, strict=True, warp_barrier=True):
    red = [k for k, (_, i) in enumerate(ins) if i.startswith("BAR.RED")]
    assert len(red) == 1, f"expected one predicate barrier in the sweep loop, found {len(red)}"
    bar = red[0]
    loads = [k for k, (_, i) in enumerate(ins) if VEC_LD.match(i)]
    stores = [k for k, (_, i) in enumerate(ins) if VEC_ST.match(i)]
    assert stores, "no vector stores found"
    before = [k for k in loads if k < bar]
    between = [k for k in loads if bar < k < min(stores)]
    assert len(before) + len(between) == len(loads), "vector loads after the first store"
    plain = [k for k in loads if ".CONSTANT" not in ins[k][1] and ".EF" not in ins[k][1]]
    if strict:
        assert len(before) == n_self, f"{len(before)} self loads in front of the barrier, expected {n_self}"
        assert len(between) == n_nbr, (f"{len(between)} neighbour loads between the barrier and the first store, expected "
                                       f"{n_nbr}: a hoisted neighbour load is fetched from HBM twice")
        if warp_barrier:
            # fused float kernels (round 2): the per-warp named barrier (32 threads), all loads on the coherent path
            assert "0x20" in ins[bar][1], ins[bar][1]
            assert not any(".CONSTANT" in ins[k][1] for k in loads)
        else:
            # CTA-wide barrier: the read-only (.CONSTANT) path is used for self loads only (ptxas may move those)
            assert all(k < bar for k in loads if ".CONSTANT" in ins[k][1])
    else:
        # register-starved instantiations (float64): ptxas sinks a few SELF loads below the barrier, which is harmless;
        # what must hold is that every neighbour load (coherent path) is behind it
        assert len(loads) == n_self + n_nbr and len([k for k in plain if k > bar]) == n_nbr
        assert all(k > bar for k in plain)
    shfl = [k for k, (_, i) in enumerate(ins) if i.startswith("SHFL.UP") or i.startswith("SHFL.DOWN")]
    assert any(min(before) < k < bar for k in shfl), "the shuffle that feeds the barrier's predicate must sit before it"


def test_fused_kernel_issues_neighbour_loads_behind_the_barrier():
    # tv_fused_kernel<float, 4, FISTA, AX2, !PEER, !SSE, !MIRROR>: the headline kernel (config 3 / config 5)
    ins = sass_of("tv_fused_kernelIfLi4ELb1ELb1ELb0ELb0ELb0E")
    check_two_phase(ins, 10, 12)                          # f, u, b x4, d x4 | (u-, u+, b+, d+) x 3 far axes


def test_fused_variants_keep_the_order():
    for frag, n_self, n_nbr in (("tv_fused_kernelIfLi4ELb0ELb1ELb0ELb0ELb0E", 6, 9),       # 4-D unaccelerated
                                ("tv_fused_kernelIfLi4ELb1ELb0ELb0ELb0ELb0E", 8, 8),       # 3-D FISTA
                                ("tv_fused_kernelIfLi4ELb0ELb0ELb0ELb0ELb0E", 5, 6),       # 3-D unaccelerated
                                ("tv_fused_kernelIdLi2ELb1ELb1ELb0ELb0ELb0E", 10, 12),     # float64 (not strict)
                                ("tv_fused_iso_kernelIfLb1ELb1ELb1E", 10, 17)):            # half-isotropic: + 5 partner loads
        check_two_phase(sass_of(frag), n_self, n_nbr, strict="IdLi2" not in frag)


def test_half_step_b_keeps_the_order():
    check_two_phase(sass_of("tv_datacube_kernelIfLi4ELb1ELb0E"), 6, 3, warp_barrier=False)     # f, u, b x4 | b+ on the three far axes
