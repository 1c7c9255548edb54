"""The built library really contains the Blackwell instructions the design claims (B200_PROFILING.md: the PTX names
never appear in SASS; these are the mnemonics to look for).  Runs on the CPU build box: cuobjdump needs no GPU."""
import os
import shutil
import subprocess

import pytest

from inclusivegan_b200.build import LIB_PATH, build


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    build()
    out = subprocess.run([exe, "-sass", LIB_PATH], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    return out.stdout


def functions(sass_text, needle):
    cur, hit = None, set()
    for line in sass_text.splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            cur = s.split(":", 1)[1].strip()
        elif needle in s and cur:
            hit.add(cur)
    return hit


def test_library_targets_sm_100a_only(sass):
    archs = {l.split("=")[1].strip() for l in sass.splitlines() if l.strip().startswith("arch =")}
    assert archs == {"sm_100a"}, archs


def test_distance_kernel_uses_tcgen05_tmem_and_tma(sass):
    mma = functions(sass, "UTCHMMA")                 # tcgen05.mma
    assert mma and all("dist_topc_kernel" in f for f in mma)
    assert functions(sass, "UTCHMMA.2CTA")           # cta_group::2 flavour
    assert functions(sass, "UTMALDG") == mma         # TMA loads feed exactly the MMA kernels
    assert functions(sass, "LDTM") == mma            # tcgen05.ld: accumulators read back from TMEM
    assert functions(sass, "UTCBAR.2CTA.MULTICAST")  # tcgen05.commit ... multicast frees both CTAs' stages
    # the instantiations the host dispatches: top-C with C in {16, 32, 64} and the collect mode, each as 1-CTA and 2-CTA
    assert len(mma) >= 8, sorted(mma)


def test_exact_paths_are_float64(sass):
    dfma = functions(sass, "DFMA")
    for name in ("rerank_kernel", "rerank_collect_kernel", "scan_dist_kernel", "project_kernel", "ball_member_kernel"):
        assert any(name in f for f in dfma), name


def test_every_kernel_is_the_repos_own(sass):
    """No library kernel (cub / thrust / cuBLAS device code) is linked into the product: every SASS function belongs to
    namespace b200 (csrc/*.cuh).  Round 1 sorted k > 32 results with cub::DeviceSegmentedRadixSort."""
    names = {l.split(":", 1)[1].strip() for l in sass.splitlines() if l.strip().startswith("Function :")}
    assert names
    foreign = sorted(n for n in names if not n.startswith("_ZN4b200"))
    assert not foreign, foreign[:5]
    assert any("scan_topk_kernel" in n for n in names) and any("rerank_warp_kernel" in n for n in names)
