"""The C-ABI library loads on a CPU-only box, exports every symbol include/*.h declares, and
refuses to compute without a GPU (no CPU fallback)."""

import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from advspec_b200 import engine as eng
from advspec_b200 import model_spec, weights

HEADER = Path(__file__).resolve().parents[1] / "include" / "advspec_engine.h"


def test_every_declared_symbol_is_exported_and_bound():
    declared = set(re.findall(r"\b(advspec_[a-z_0-9]+)\s*\(", HEADER.read_text())) - {"advspec_status"}
    assert declared == set(eng.EXPORTED_SYMBOLS), declared ^ set(eng.EXPORTED_SYMBOLS)
    lib = eng.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    assert C.sizeof(eng.ModelDesc) == 96
    assert C.sizeof(eng.Timing) == 40


@pytest.mark.parametrize("name", ["tiny-llama", "tiny-qwen2", "tiny-gemma", "llama-3-8b", "gemma-7b", "phi-3-mini"])
def test_blob_layout_agrees_between_c_and_python(name):
    lib = eng.load_library()
    spec = model_spec.resolve(name)
    d = eng.make_desc(spec, 512, 64, 4)
    lay = weights.blob_layout(spec)
    assert lib.advspec_weight_blob_bytes(C.byref(d)) == lay.total
    for (layer, tname), (off, _, _) in lay.offsets.items():
        assert lib.advspec_weight_offset(C.byref(d), layer, tname.encode()) == off
    if not spec.qkv_bias:
        assert lib.advspec_weight_offset(C.byref(d), 0, b"bqkv") == C.c_size_t(-1).value


def test_invalid_descs_are_rejected():
    lib = eng.load_library()
    spec = model_spec.resolve("tiny-llama")
    d = eng.make_desc(spec, 512, 64, 4)
    d.head_dim = 80
    assert lib.advspec_weight_blob_bytes(C.byref(d)) == 0
    d = eng.make_desc(spec, 512, 64, 9)
    assert lib.advspec_weight_blob_bytes(C.byref(d)) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu():
    with pytest.raises(eng.EngineError, match="no CPU fallback"):
        eng.Engine(model_spec.resolve("tiny-llama"), 0, 256, 16, 2)


def test_bf16_rounding_helper_matches_torch():
    import numpy as np

    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    ours = weights.bf16_bits_to_f32(weights.f32_to_bf16_bits(x))
    ref = torch.from_numpy(x).bfloat16().float().numpy()
    assert (ours == ref).all()


def test_c_example_compiles_and_links_against_the_header(tmp_path):
    """examples/critique_panel.c is a pure-C consumer of include/advspec_engine.h: it must compile with a
    C compiler (the header is C, not C++) and link against the built library (no GPU needed to link)."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    cc = shutil.which("gcc") or shutil.which("cc")
    lib = root / "adversarial-spec_b200" / "libadvspec_b200.so"
    if cc is None or not lib.exists():
        import pytest
        pytest.skip("needs a C compiler and the built library")
    exe = tmp_path / "critique_panel"
    p = subprocess.run([cc, "-std=c99", "-O2", "-Wall", "-Werror", f"-I{root / 'include'}",
                        str(root / "examples" / "critique_panel.c"), f"-L{lib.parent}", "-ladvspec_b200",
                        "-Wl,--allow-shlib-undefined", "-o", str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert exe.exists()
