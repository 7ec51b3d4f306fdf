"""Code-generation guard for the raster kernel (CPU only: cuobjdump reads the built object).

The recurrence path must stay a copy-free run of packed FP32 instructions.  ptxas sometimes
renames the 64-bit accumulators out of place and then pays ~20 MOV / IMAD.MOV per splat to
move them back (seen with -O3 and after innocuous source edits: 9-11 % slower on the B200,
DESIGN.md section 4.2), so this is checked where it is cheap to check."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "genetic-gaussian-splats_b200", "build", "ggs_raster.o")


def production_kernel_sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not os.path.exists(OBJ):
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "ggs_build", os.path.join(ROOT, "genetic-gaussian-splats_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    text = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    prod = [f for f in funcs if "raster_kernelILb0ELb0E" in f.split("\n", 1)[0]]
    assert len(prod) == 1, "raster_kernel<false, false> not found in the object"
    ops = []
    for line in prod[0].splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", line)
        if m:
            ops.append(m.group(1))
    return ops


def test_packed_fp32_and_mufu_are_in_the_kernel():
    ops = production_kernel_sass()
    names = [o.split()[1] if o.startswith("@") else o.split()[0] for o in ops]
    assert sum(n.startswith("FFMA2") for n in names) >= 20
    assert sum(n.startswith("FMUL2") for n in names) >= 8
    assert sum(n.startswith("FADD2") for n in names) >= 4
    assert sum(n.startswith("MUFU.EX2") for n in names) >= 8
    # a few spilled bytes are tolerated in the kernel prologue / epilogue, never in the
    # composite loop (from its first list load to the last exact-path MUFU.EX2)
    mufu = [i for i, n in enumerate(names) if n.startswith("MUFU.EX2")]
    hot = names[max(0, mufu[0] - 45): mufu[-1] + 20]
    assert not any(n.startswith(("STL", "LDL")) for n in hot), "local-memory traffic in the composite loop"
    assert not any(n.startswith(("S2R", "I2FP")) for n in hot), "thread constants rematerialised in the loop"


def test_recurrence_path_has_no_register_copies():
    ops = production_kernel_sass()
    first = next(i for i, o in enumerate(ops) if "MUFU.EX2" in o)
    # the recurrence block: from its first MUFU.EX2 to the BRA that closes it
    end = next(i for i in range(first, len(ops)) if re.match(r"(@!?U?P\d+\s+)?BRA\b", ops[i]))
    block = ops[first:end]
    packed = [o for o in block if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", o)]
    copies = [o for o in block if re.match(r"(IMAD\.MOV|MOV)\b", o)]
    assert len(packed) >= 24, block
    assert len(copies) <= 2, f"{len(copies)} register copies in the recurrence path:\n" + "\n".join(block)
    assert len(block) <= len(packed) + 12


VARIANTS = [
    ("8 rows x 2 warps", ["-DGGS_WARPS=2"]),
    ("8 rows x 1 warp", ["-DGGS_WARPS=1"]),
    ("16 rows x 2 warps", ["-DGGS_ROWS=16", "-DGGS_WARPS=2"]),
    ("list of 384, scan rounds of 128", ["-DGGS_LIST_CAP=384", "-DGGS_SCAN_CHUNK=128"]),
    ("pixel state as C++ values (fallback for the named PTX registers)", ["-DGGS_NAMED_REGS=0"]),
]


@pytest.mark.parametrize("name,flags", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_every_supported_geometry_of_the_raster_compiles(name, flags, tmp_path):
    """The pixel state lives in hand-declared PTX registers shared by many asm statements: a
    variant (rows per thread, warps per CTA, list sizes) that breaks their single declaration
    or their scoping must fail HERE, at ptxas, not on the GPU box."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    pkg = os.path.join(ROOT, "genetic-gaussian-splats_b200")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "csrc"), *flags,
           "-c", os.path.join(pkg, "csrc", "ggs_raster.cu"), "-o", str(tmp_path / "raster.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
