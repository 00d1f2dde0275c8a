"""FastAPI wiring of the serving path.

CPU part (no GPU needed): the reference's own `app/main.py` imported UNCHANGED over this repo's RecommendationService
(when /root/reference is present), this repo's `create_app` with a stand-in service (routes, status codes, response
models), and the micro-batcher's coalescing logic.  GPU part (`-m gpu`): the same HTTP surface over the real service --
every answer equal to the direct service call -- and concurrent /recommend requests answered from coalesced batches."""

import asyncio
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import pytest

fastapi = pytest.importorskip("fastapi")
from fastapi.testclient import TestClient  # noqa: E402

REF_MAIN = Path("/root/reference/app/main.py")


class StubService:
    """Duck-typed stand-in with the reference service's surface (app/recommendation_service.py:18-110)."""

    version = "1.2.0"

    def __init__(self):
        self.calls = []

    def is_ready(self):
        return True

    def _recs(self, user_id, k):
        return [{"item_id": f"{user_id}-{i}", "score": 1.0 - 0.01 * i, "rank": i + 1} for i in range(k)]

    def recommend(self, user_id, k=10):
        self.calls.append(("recommend", user_id, k))
        if user_id == "boom":
            raise RuntimeError("kaput")
        return self._recs(user_id, k)

    def recommend_batch(self, user_ids, k=10):
        self.calls.append(("recommend_batch", tuple(user_ids), k))
        return [{"user_id": u, "recommendations": self._recs(u, k), "status": "success"} if u != "boom" else
                {"user_id": u, "recommendations": [], "status": "error", "error": "kaput"} for u in user_ids]

    def score(self, user_id, item_ids):
        if user_id == "nobody":
            raise ValueError("User 'nobody' not found in vocabulary.")
        return {i: 0.5 for i in item_ids}

    def get_model_info(self):
        return {"version": self.version, "model_path": "<stub>", "num_users": 3, "faiss_index_items": 7}


@pytest.mark.skipif(not REF_MAIN.exists(), reason="the reference tree is not present on this box")
def test_reference_main_runs_unchanged_over_this_service(rs, monkeypatch):
    """app/main.py:21 does `from app.recommendation_service import RecommendationService`; with that module name bound to
    this package the reference application imports and serves: start-up with a missing model directory degrades to 503
    exactly as with the reference service (FileNotFoundError from load(), app/main.py:110-119), and its pydantic response
    models accept this service's return values."""
    pkg = types.ModuleType("app")
    pkg.__path__ = []
    shim = types.ModuleType("app.recommendation_service")
    shim.RecommendationService = rs.RecommendationService
    monkeypatch.setitem(sys.modules, "app", pkg)
    monkeypatch.setitem(sys.modules, "app.recommendation_service", shim)
    spec = importlib.util.spec_from_file_location("reference_app_main", REF_MAIN)
    main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(main)
    assert main.RecommendationService is rs.RecommendationService
    with TestClient(main.app) as client:             # runs the reference's startup handler: load() -> FileNotFoundError
        r = client.get("/health")
        assert r.status_code == 200 and r.json() == {"status": "degraded", "model_loaded": False, "model_version": None}
        assert client.get("/").status_code == 200
        assert client.post("/recommend", json={"user_id": "1", "k": 3}).status_code == 503
        assert client.get("/model/info").status_code == 503
        # a ready service of THIS repo's class (no artifacts on a CPU box: the two numeric methods are canned)
        svc = rs.RecommendationService.from_parts(object(), types.SimpleNamespace(ntotal=7), ["1", "2"], {"0": "a"})
        assert svc.is_ready()
        stub = StubService()
        svc.recommend, svc.score = stub.recommend, stub.score
        main.recommendation_service = svc
        r = client.get("/health").json()
        assert r == {"status": "healthy", "model_loaded": True, "model_version": "1.2.0"}
        r = client.post("/recommend", json={"user_id": "1", "k": 3})
        assert r.status_code == 200
        body = r.json()
        assert body["count"] == 3 and body["model_version"] == "1.2.0" and [x["rank"] for x in body["recommendations"]] == [1, 2, 3]
        assert client.post("/recommend", json={"user_id": "1", "k": 101}).status_code == 422      # the reference's own bound
        assert client.post("/score", json={"user_id": "nobody", "item_ids": ["a"]}).status_code == 404
        assert client.post("/score", json={"user_id": "1", "item_ids": ["a", "b"]}).json() == {"user_id": "1", "scores": {"a": 0.5, "b": 0.5}}
        info = client.get("/model/info").json()
        assert info["num_users"] == 2 and info["faiss_index_items"] == 7 and info["version"] == "1.2.0"


def test_create_app_mirrors_the_reference_routes_and_status_codes(rs):
    stub = StubService()
    app = rs.create_app(service=stub, micro_batch=False)
    with TestClient(app) as client:
        assert client.get("/health").json() == {"status": "healthy", "model_loaded": True, "model_version": "1.2.0"}
        assert client.get("/").json()["status"] == "running"
        body = client.post("/recommend", json={"user_id": "7", "k": 4}).json()
        assert body["user_id"] == "7" and body["count"] == 4 and body["recommendations"][3] == {"item_id": "7-3", "score": 0.97, "rank": 4}
        assert client.post("/recommend", json={"user_id": "7"}).json()["count"] == 10            # k defaults to 10
        assert client.post("/recommend", json={"user_id": "7", "k": 0}).status_code == 422
        assert client.post("/recommend", json={"user_id": "boom", "k": 2}).status_code == 500
        assert client.post("/score", json={"user_id": "nobody", "item_ids": ["x"]}).status_code == 404
        assert client.post("/score", json={"user_id": "7", "item_ids": []}).status_code == 422
        assert client.post("/score", json={"user_id": "7", "item_ids": ["x"]}).json() == {"user_id": "7", "scores": {"x": 0.5}}
        assert client.get("/model/info").json()["faiss_index_items"] == 7
        r = client.post("/recommend/batch", params={"user_ids": ["0", "boom", "0"], "k": 3}).json()     # app/test_api.py:107-130
        assert r["count"] == 3 and [x["user_id"] for x in r["recommendations"]] == ["0", "boom", "0"]
        assert [x["status"] for x in r["recommendations"]] == ["success", "error", "success"]
    # no service and no artifacts: every model endpoint answers 503, /health says degraded (app/main.py:155-156)
    with TestClient(rs.create_app(model_dir="/nonexistent/model/dir")) as client:
        assert client.get("/health").json()["status"] == "degraded"
        assert client.post("/recommend", json={"user_id": "1"}).status_code == 503
        assert client.post("/score", json={"user_id": "1", "item_ids": ["a"]}).status_code == 503
        assert client.post("/recommend/batch", params={"user_ids": ["1"]}).status_code == 503


def test_micro_batcher_coalesces_concurrent_requests(rs):
    stub = StubService()

    async def scenario():
        b = rs.MicroBatcher(stub, max_batch=8, max_wait_ms=20.0)
        ks = [3, 10, 1, 7, 5]
        out = await asyncio.gather(*[b.recommend(f"u{i}", k) for i, k in enumerate(ks)])
        # one batched call with k = max(k), every caller gets the first k entries of ITS user's list
        assert [len(o) for o in out] == ks
        assert all(o == stub._recs(f"u{i}", 10)[:k] for i, (o, k) in enumerate(zip(out, ks)))
        assert stub.calls == [("recommend_batch", ("u0", "u1", "u2", "u3", "u4"), 10)]
        # a full batch flushes at once, the rest after the wait; a failing user fails only its own request
        res = await asyncio.gather(*[b.recommend("boom" if i == 4 else f"v{i}", 2) for i in range(11)], return_exceptions=True)
        assert isinstance(res[4], RuntimeError) and all(len(r) == 2 for i, r in enumerate(res) if i != 4)
        assert b.stats["requests"] == 16 and b.stats["batches"] == 3 and b.stats["largest_batch"] == 8
        return b

    asyncio.run(scenario())


def test_micro_batched_app_answers_like_the_direct_call(rs):
    stub = StubService()
    app = rs.create_app(service=stub, micro_batch=True, max_wait_ms=0.5)
    with TestClient(app) as client:
        for user, k in (("a", 3), ("b", 10), ("a", 1)):
            body = client.post("/recommend", json={"user_id": user, "k": k}).json()
            assert body["recommendations"] == stub._recs(user, k) and body["count"] == k
        assert client.post("/recommend", json={"user_id": "boom", "k": 2}).status_code == 500
    assert all(c[0] == "recommend_batch" for c in stub.calls)


@pytest.mark.gpu
def test_http_surface_over_the_real_service(rs):
    """The whole path behind the HTTP boundary on the GPU: answers equal the direct service calls (which
    tests/test_gpu_parity.py pins against the oracle), cold start, 404, the batch endpoint, and concurrent requests
    answered from coalesced batches with the same results."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    cfg = rs.ModelConfig()
    uv, iv = rs.synth.make_vocab(300), rs.synth.make_vocab(5000)
    enc = rs.MultiTowerModel(cfg, uv, iv, {})
    rs.load_weights(enc, rs.synth.make_weights(cfg, 300, 5000, seed=12))
    index, item_map = rs.build_item_index(enc, iv)
    svc = rs.RecommendationService.from_parts(enc, index, uv, item_map)
    app = rs.create_app(service=svc, micro_batch=True, max_batch=16, max_wait_ms=5.0)
    with TestClient(app) as client:
        assert client.get("/health").json()["model_loaded"] is True
        assert client.get("/model/info").json() == svc.get_model_info()
        for u, k in ((uv[0], 10), (uv[17], 1), (uv[299], 100)):
            body = client.post("/recommend", json={"user_id": u, "k": k}).json()
            assert body["recommendations"] == svc.recommend(u, k) and body["count"] == k
        cold = client.post("/recommend", json={"user_id": "nobody", "k": 3}).json()
        assert cold["recommendations"] == svc.recommend("nobody", 3)
        sc = client.post("/score", json={"user_id": uv[5], "item_ids": iv[:50]}).json()["scores"]
        assert sc == svc.score(uv[5], iv[:50])
        assert client.post("/score", json={"user_id": "nobody", "item_ids": iv[:2]}).status_code == 404
        ids = [uv[1], "nobody", uv[2], uv[1]]
        batch = client.post("/recommend/batch", params={"user_ids": ids, "k": 5}).json()
        assert batch["count"] == 4 and [b["user_id"] for b in batch["recommendations"]] == ids
        assert all(b["recommendations"] == svc.recommend(b["user_id"], 5) for b in batch["recommendations"])
        # 48 concurrent requests: same answers, fewer GPU passes than requests
        want = {u: svc.recommend(u, 10) for u in uv[:48]}
        before = dict(app.state.recsys["batcher"].stats)
        with ThreadPoolExecutor(max_workers=16) as ex:
            got = list(ex.map(lambda u: (u, client.post("/recommend", json={"user_id": u, "k": 10}).json()["recommendations"]), uv[:48]))
        assert all(want[u] == r for u, r in got)
        after = app.state.recsys["batcher"].stats
        assert after["requests"] - before["requests"] == 48 and after["batches"] - before["batches"] < 48
    torch.cuda.synchronize()
