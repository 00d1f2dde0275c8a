"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's serving/scoring hot path
(`src/models.py`, `app/recommendation_service.py`, and the faiss-cpu /
TensorFlow semantics those files delegate to).  It exists to *check* the CUDA
path and to be timed as the CPU baseline.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product package never does.

PARITY: composition pinned, primitive arithmetic UNPINNED.  TensorFlow, Keras,
tensorflow-recommenders and faiss are not importable in this image and the
reference ships no golden vectors, fixtures or numeric tests for this path
(SURVEY.md section 8c).  The reference's own Python files do run here over numpy
stand-ins of those four modules (tests/golden/make_reference_golden.py ->
tests/golden/reference_run.*), and tests/test_reference_run.py holds this
restatement to those outputs (1e-12 for the ranking forward, item for item for
the service).  That pins the composition of the path to the reference's code;
the arithmetic inside Dense / IndexFlatIP.search / normalize_L2 is still a
statement of their documented behaviour, not TF's or faiss's binaries.  Known-
answer tests (tests/test_oracle_known_answers.py), a torch-CPU cross-check and
fixtures generated from the oracle itself (tests/golden/make_golden.py) cover
the rest.
"""

from .build import build_chain_lib, load_chain_lib  # noqa: F401
from .flat_ip import (  # noqa: F401
    IndexFlatIP,
    matvec,
    merge_topk,
    normalize_L2,
    search_blas,
    search_chain,
)
from .models import (  # noqa: F401
    DeepCrossNetwork,
    MultiTaskModel,
    MultiTowerModel,
    dense,
    string_lookup,
)
from .service import RecommendationService  # noqa: F401
