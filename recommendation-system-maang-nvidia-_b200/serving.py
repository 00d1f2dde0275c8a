"""FastAPI wiring of the serving path: the reference's HTTP surface over the B200 service.

`app/main.py` of the reference imports `RecommendationService` from `app.recommendation_service` and needs nothing else
from it, so it runs unchanged once that import resolves to this package (INTEGRATION.md shows the one-line shim;
tests/test_serving.py does exactly that with the reference tree when it is present).  `create_app` below is the same
application built on this side of the boundary, for deployments that do not carry the reference tree:

    GET  /health            HealthResponse {status, model_loaded, model_version}        app/main.py:132-140
    GET  /                  welcome message                                             app/main.py:143-150
    POST /recommend         RecommendationRequest {user_id, k in 1..100} -> RecommendationResponse
                            {user_id, recommendations [{item_id, score, rank}], count, model_version}   app/main.py:153-171
    POST /score             ScoreRequest {user_id, item_ids 1..100} -> {user_id, scores}; unknown user -> 404   app/main.py:174-187
    GET  /model/info        get_model_info()                                            app/main.py:190-196
    POST /recommend/batch   user_ids (repeated query parameter), k -> {"recommendations": [{user_id, recommendations,
                            status}], "count"}          (the endpoint app/test_api.py:107-130 exercises; the body is
                            RecommendationService.recommend_batch, app/model_service.py:200-228)

Status codes as in the reference: 503 while no service is loaded, 404 for ValueError (unknown user in /score), 500 for
anything else.

Micro-batching (new here).  One /recommend request is one user-tower pass and one catalog pass for ONE query: 0.5 ms
of device time of which the catalog scan is bound by HBM whether it carries 1 query or 100 (below ~200 queries the
filter streams the bf16 shadow once per batch).  `MicroBatcher` therefore coalesces the /recommend calls that arrive
within `max_wait_ms` of each other (up to `max_batch`) into ONE `recommend_batch` call -- one tower pass, one catalog
pass -- and hands each caller the first k entries of its user's list (the lists are sorted with a total order, so the
top-k is a prefix of the top-k_max).  The GPU call runs in a worker thread; the event loop keeps accepting requests.
"""

import asyncio
import logging
from typing import Dict, List, Optional

logger = logging.getLogger(__name__)


class MicroBatcher:
    """Coalesces concurrent recommend(user_id, k) calls into recommend_batch calls.  `stats` counts what happened."""

    def __init__(self, service, max_batch: int = 64, max_wait_ms: float = 1.0):
        self.service = service
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._pending: List = []            # (user_id, k, future)
        self._task: Optional[asyncio.Task] = None
        self._lock: Optional[asyncio.Lock] = None
        self.stats = {"requests": 0, "batches": 0, "largest_batch": 0}

    async def recommend(self, user_id: str, k: int) -> List[Dict]:
        loop = asyncio.get_running_loop()
        if self._lock is None:
            self._lock = asyncio.Lock()
        fut = loop.create_future()
        async with self._lock:
            self._pending.append((user_id, int(k), fut))
            self.stats["requests"] += 1
            if len(self._pending) >= self.max_batch:
                self._flush_now(loop)
            elif self._task is None or self._task.done():
                self._task = loop.create_task(self._flush_later(loop))
        return await fut

    async def _flush_later(self, loop):
        await asyncio.sleep(self.max_wait)
        async with self._lock:
            self._flush_now(loop)

    def _flush_now(self, loop):
        batch, self._pending = self._pending, []
        if not batch:
            return
        self.stats["batches"] += 1
        self.stats["largest_batch"] = max(self.stats["largest_batch"], len(batch))
        loop.create_task(self._run(loop, batch))

    async def _run(self, loop, batch):
        users = [u for u, _, _ in batch]
        kmax = max(k for _, k, _ in batch)
        try:
            # the GPU call blocks its thread (it ends with a D2H copy): keep it off the event loop
            results = await loop.run_in_executor(None, self.service.recommend_batch, users, kmax)
            for (user, k, fut), res in zip(batch, results):
                if fut.done():
                    continue
                if res.get("status") == "success":
                    fut.set_result([dict(r) for r in res["recommendations"][:k]])
                else:
                    fut.set_exception(RuntimeError(res.get("error", "recommendation failed")))
        except Exception as e:  # noqa: BLE001
            for _, _, fut in batch:
                if not fut.done():
                    fut.set_exception(e)


def create_app(service=None, model_dir: str = "outputs/models/experiment_001", micro_batch: bool = True,
               max_batch: int = 64, max_wait_ms: float = 1.0):
    """The reference's FastAPI application (app/main.py) over `RecommendationService`.  `service`: an already built service
    (tests, notebooks); otherwise the artifacts in `model_dir` are loaded at start-up, as app/main.py:105-119 does."""
    from fastapi import FastAPI, HTTPException, Query
    from fastapi.middleware.cors import CORSMiddleware
    from pydantic import BaseModel, Field

    from .service import RecommendationService

    class RecommendationRequest(BaseModel):
        user_id: str = Field(..., description="User ID to get recommendations for")
        k: int = Field(10, ge=1, le=100, description="Number of recommendations to return")

    class RecommendationItem(BaseModel):
        item_id: str
        score: float
        rank: int

    class RecommendationResponse(BaseModel):
        user_id: str
        recommendations: List[RecommendationItem]
        count: int
        model_version: str

    class ScoreRequest(BaseModel):
        user_id: str
        item_ids: List[str] = Field(..., min_length=1, max_length=100)

    class ScoreResponse(BaseModel):
        user_id: str
        scores: Dict[str, float]

    class HealthResponse(BaseModel):
        status: str
        model_loaded: bool
        model_version: Optional[str] = None

    app = FastAPI(title="Recommendation System API",
                  description="Two-Tower + DCN recommendation API on the B200-native serving path.", version="1.1.0",
                  docs_url="/docs", redoc_url="/redoc")
    app.add_middleware(CORSMiddleware, allow_origins=["*"], allow_credentials=True, allow_methods=["*"], allow_headers=["*"])
    state = {"service": service, "batcher": None}
    app.state.recsys = state

    def ready():
        s = state["service"]
        return s is not None and s.is_ready()

    def batcher():
        if state["batcher"] is None and micro_batch and ready():
            state["batcher"] = MicroBatcher(state["service"], max_batch=max_batch, max_wait_ms=max_wait_ms)
        return state["batcher"]

    @app.on_event("startup")
    async def startup_event():
        if state["service"] is not None:
            return
        try:
            svc = RecommendationService(model_dir=model_dir)
            svc.load()
            state["service"] = svc
            logger.info("Recommendation service is loaded and ready.")
        except Exception as e:  # noqa: BLE001 -- as the reference: the endpoints then answer 503
            logger.error(f"Failed to load model during startup: {e}")

    @app.on_event("shutdown")
    async def shutdown_event():
        state["service"] = None
        state["batcher"] = None

    @app.get("/health", response_model=HealthResponse, tags=["Health"])
    async def health_check():
        ok = ready()
        return HealthResponse(status="healthy" if ok else "degraded", model_loaded=ok,
                              model_version=state["service"].get_model_info()["version"] if ok else None)

    @app.get("/", tags=["Info"], include_in_schema=False)
    async def root():
        return {"message": "Welcome to the Recommendation System API", "status": "running", "documentation": "/docs"}

    @app.post("/recommend", response_model=RecommendationResponse, tags=["Recommendations"])
    async def get_recommendations(request: RecommendationRequest):
        if not ready():
            raise HTTPException(status_code=503, detail="Model not loaded. Service is unavailable.")
        try:
            b = batcher()
            if b is not None:
                recs = await b.recommend(request.user_id, request.k)
            else:
                recs = state["service"].recommend(user_id=request.user_id, k=request.k)
            return RecommendationResponse(user_id=request.user_id, recommendations=recs, count=len(recs),
                                          model_version=state["service"].get_model_info()["version"])
        except ValueError as e:
            raise HTTPException(status_code=404, detail=str(e))
        except Exception as e:  # noqa: BLE001
            logger.error(f"Error during recommendation for user '{request.user_id}': {e}")
            raise HTTPException(status_code=500, detail="Internal server error")

    @app.post("/recommend/batch", tags=["Recommendations"])
    async def get_batch_recommendations(user_ids: List[str] = Query(...), k: int = Query(10, ge=1, le=100)):
        if not ready():
            raise HTTPException(status_code=503, detail="Model not loaded. Service is unavailable.")
        try:
            loop = asyncio.get_running_loop()
            results = await loop.run_in_executor(None, state["service"].recommend_batch, list(user_ids), int(k))
            return {"recommendations": results, "count": len(results)}
        except Exception as e:  # noqa: BLE001
            logger.error(f"Error during batch recommendation: {e}")
            raise HTTPException(status_code=500, detail="Internal server error")

    @app.post("/score", response_model=ScoreResponse, tags=["Scoring"])
    async def score_items(request: ScoreRequest):
        if not ready():
            raise HTTPException(status_code=503, detail="Model not loaded. Service is unavailable.")
        try:
            scores = state["service"].score(user_id=request.user_id, item_ids=request.item_ids)
            return ScoreResponse(user_id=request.user_id, scores=scores)
        except ValueError as e:
            raise HTTPException(status_code=404, detail=str(e))
        except Exception as e:  # noqa: BLE001
            logger.error(f"Error during scoring for user '{request.user_id}': {e}")
            raise HTTPException(status_code=500, detail="Internal server error")

    @app.get("/model/info", tags=["Model"])
    async def get_model_info():
        if not ready():
            raise HTTPException(status_code=503, detail="Model not loaded. Service is unavailable.")
        return state["service"].get_model_info()

    return app
