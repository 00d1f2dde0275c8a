"""The built library really contains Blackwell tensor-core / TMA instructions where DESIGN.md says so: SASS mnemonics per
kernel from `cuobjdump -sass` (no GPU needed; the library is built by the session fixture / build.py).  Guards against a
refactor that silently turns a tcgen05 / TMA kernel into a plain one."""

import importlib.util
import shutil
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_tcgen05_and_tma_mnemonics_are_where_the_design_says():
    import recsys_b200
    recsys_b200._lib.load()                               # builds the .so if it is missing or stale
    spec = importlib.util.spec_from_file_location("sass_evidence", ROOT / "tools" / "sass_evidence.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    k = mod.collect()

    def has(kernel_prefix, *mnemonics):
        hits = [v for name, v in k.items() if name.startswith(kernel_prefix)]
        assert hits, f"no kernel named {kernel_prefix}* carries tensor / TMA instructions"
        for v in hits:
            for m in mnemonics:
                assert any(key.startswith(m) for key in v), (kernel_prefix, m, v)

    has("tc_filter_kernel", "UTC", "LDTM", "UTMALDG")                   # retrieval filter: tcgen05.mma + tcgen05.ld + TMA
    has("scan_topk_kernel", "UTMALDG")                                  # exact scan: TMA-streamed catalog
    has("tower_pp_kernel", "UTC", "LDTM", "STTM", "UTMALDG")            # towers: A operand through TMEM (tcgen05.st)
    has("deep_heads_kernel", "UTC", "LDTM", "STTM", "UTMALDG")
    has("dense_tc_kernel", "UTC", "LDTM", "UTMALDG")
    assert {"UTMASTG", "UBLKCP"} <= set(k["tower_pp_kernel<2>"])        # mode 2: outputs leave through TMA stores
