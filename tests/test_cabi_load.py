"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/hoigen_b200.h declares."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "hoigen_b200.h").read_text()
    return sorted(set(re.findall(r"HOIGEN_API\s+[\w\s\*]+?\b(hoigen_\w+)\s*\(", text)))


def test_library_builds_and_exports_all_declared_symbols():
    from hoigen_b200 import _build, _cabi
    _build.build()
    lib = _cabi.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared, "ctypes bindings and header disagree"
    assert lib.hoigen_abi_version() == _cabi.ABI_VERSION == 2


def test_sass_contains_blackwell_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from hoigen_b200 import _build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_build.build())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path present"


def test_tensor_core_kernels_contain_utcmma():
    """Per kernel: the GEMMs, attention, the adapter block, the fused cache kernel (both forms), the RoIAlign kernel, the
    ResNet-50 stem convolution and DETR's attention issue tcgen05.mma (SASS UTCHMMA) themselves."""
    import shutil
    import subprocess
    from hoigen_b200 import _build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_build.build())], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif "UTCHMMA" in line and cur:
            counts[cur] = counts.get(cur, 0) + 1
    for kernel in ("gemm2_bf16_kernel", "gemm_bf16_kernel", "attention_kernel", "adapter_tc_kernel", "cache_fused_pair_kernel",
                   "cache_fused_kernel", "roi_tc_kernel", "stem_conv_kernel", "attention_heads32_tc_kernel"):
        assert any(kernel in k and n > 0 for k, n in counts.items()), f"{kernel}: no UTCHMMA in its SASS"


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from hoigen_b200 import _cabi
    from hoigen_b200.encoder import VisionTransformer
    vt = VisionTransformer().eval()
    with pytest.raises(_cabi.HoigenError):
        vt(torch.zeros(1, 3, 224, 224), (torch.zeros(1, 4, 64), torch.zeros(1, 4, dtype=torch.bool)))


def test_product_code_never_imports_oracle():
    for path in (ROOT / "hoigen_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
