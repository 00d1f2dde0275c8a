"""Generate tests/golden/reference_run.{npz,json} by running the REFERENCE'S OWN files.  TEST INFRASTRUCTURE ONLY.

    python tests/golden/make_reference_golden.py          (needs /root/reference; runs in this container only)

Executed unmodified, from where they lie under /root/reference:
  src/config.py                  ModelConfig
  src/models.py                  MultiTaskModel / MultiTowerModel / DeepCrossNetwork . call
  src/trainer.py:221-246         ProductionTrainer._get_item_embeddings, _save_artifacts, _build_faiss
  app/recommendation_service.py  RecommendationService . load / recommend / score / _get_popular_items / get_model_info
  src/evaluation.py:25-81        AdvancedMetrics
  src/trainer.py:194-219         ProductionTrainer._evaluate (offline recall@k by raw dot + argpartition)
  app/model_service.py           the alternate service: raw-dot recommend (:104-150), score (:152-198), recommend_batch (:200-228)
  app/main.py                    the FastAPI application: /health, /, /recommend, /score, /model/info (a request transcript)
over the numpy stand-ins of tensorflow / keras / tensorflow_recommenders / faiss in tf_faiss_standin.py (none of the
four is installed here).  So the fixtures pin the reference's COMPOSITION -- which layers, in which order, on which
tensors, the control flow of the service, the artifact files it exchanges -- while every primitive (Dense, Embedding,
StringLookup, IndexFlatIP.search, normalize_L2) computes in float64 by the stand-in's statement of its documented
behaviour.  Weights, vocabularies and ids are regenerated from seeds by recsys_b200.synth, so the files stay small.
`generate()` returns the same content without writing, which the CPU suite uses to check the committed files.
"""

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
for p in (str(ROOT), str(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

SPEC = dict(n_users=300, n_items=200, batch=96, weight_seed=62, id_seed=52, table_scale=40.0,
            users=["1", "10", "137", "299", "42"], k=10, k_over=250, score_items=40)


# other ModelConfig values through the same reference code (ranking forward only): narrower embeddings, towers of unequal
# depth, two / zero cross layers, deeper / shallower deep nets
VARIANTS = {
    "B": dict(cfg=dict(embedding_dim=64, user_tower_dims=[96, 48], item_tower_dims=[32], cross_layers=2, dnn_dims=[64, 32, 16]),
              n_users=50, n_items=70, batch=40, weight_seed=71, id_seed=72),
    "C": dict(cfg=dict(embedding_dim=64, user_tower_dims=[], item_tower_dims=[128], cross_layers=0, dnn_dims=[32]),
              n_users=20, n_items=30, batch=24, weight_seed=81, id_seed=82),
}


HTTP_REQUESTS = [
    ("GET", "/health"), ("GET", "/"), ("GET", "/model/info"),
    ("POST", "/recommend", {"user_id": "1", "k": 5}), ("POST", "/recommend", {"user_id": "10"}),
    ("POST", "/recommend", {"user_id": "nobody", "k": 3}), ("POST", "/recommend", {"user_id": "1", "k": 0}),
    ("POST", "/recommend", {"user_id": "1", "k": 101}), ("POST", "/recommend", {"k": 3}),
    ("POST", "/score", {"user_id": "137", "item_ids": ["1", "2", "no-such-movie"]}),
    ("POST", "/score", {"user_id": "nobody", "item_ids": ["1"]}), ("POST", "/score", {"user_id": "1", "item_ids": []}),
]


def variant_inputs(name):
    import recsys_b200 as rs
    v = VARIANTS[name]
    cfg = rs.ModelConfig(**v["cfg"])
    w = rs.synth.make_weights(cfg, v["n_users"], v["n_items"], seed=v["weight_seed"])
    uv, iv = rs.synth.make_vocab(v["n_users"]), rs.synth.make_vocab(v["n_items"])
    rng = np.random.default_rng(v["id_seed"])
    users = [uv[i] for i in rng.integers(0, len(uv), v["batch"])]
    items = [iv[i] for i in rng.integers(0, len(iv), v["batch"])]
    users[1], items[2] = "nobody", "no-such-movie"
    return cfg, w, uv, iv, users, items


def synth_inputs():
    """Weights, vocabularies and the ranking batch (strings, with out-of-vocabulary ids on both sides)."""
    import recsys_b200 as rs
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, SPEC["n_users"], SPEC["n_items"], seed=SPEC["weight_seed"])
    # Keras-initialised towers map every row to nearly the same direction; wider tables spread the catalog out so that the
    # top-k order is decided by gaps far above fp32 rounding
    w["user_embedding"] = (w["user_embedding"] * np.float32(SPEC["table_scale"])).astype(np.float32)
    w["item_embedding"] = (w["item_embedding"] * np.float32(SPEC["table_scale"])).astype(np.float32)
    uv, iv = rs.synth.make_vocab(SPEC["n_users"]), rs.synth.make_vocab(SPEC["n_items"])
    rng = np.random.default_rng(SPEC["id_seed"])
    users = [uv[i] for i in rng.integers(0, len(uv), SPEC["batch"])]
    items = [iv[i] for i in rng.integers(0, len(iv), SPEC["batch"])]
    for pos, s in ((3, "nobody"), (17, "0"), (40, "")):
        users[pos] = s
    for pos, s in ((5, "no-such-movie"), (17, "9999")):
        items[pos] = s
    return cfg, w, uv, iv, users, items


def _set_dense(layer, kb):
    assert layer.kernel.shape == kb[0].shape and layer.bias.shape == kb[1].shape
    layer.kernel, layer.bias = np.array(kb[0]), np.array(kb[1])


def build_reference_model(ref, cfg_dict, w, uv, iv, users, items):
    """The reference's MultiTaskModel with the synthetic weights put into the layers it created itself."""
    tf = sys.modules["tensorflow"]
    M = ref["src.models"]
    cfg = ref["src.config"].ModelConfig(**cfg_dict)
    model = M.MultiTaskModel(cfg, list(uv), list(iv), {})
    model({"user_id": tf.constant(users[:2]), "movie_id": tf.constant(items[:2])}, training=False)   # builds the layers
    enc = model.encoder
    enc.user_embedding.embeddings = np.array(w["user_embedding"])
    enc.item_embedding.embeddings = np.array(w["item_embedding"])
    for side in ("user", "item"):
        layers = getattr(enc, f"{side}_tower").layers
        assert len(layers) == len(w[f"{side}_tower"])
        for layer, kb in zip(layers, w[f"{side}_tower"]):
            _set_dense(layer, kb)
    assert len(model.dcn.cross_weights) == len(w["cross_w"]) == cfg.cross_layers
    for i in range(cfg.cross_layers):
        model.dcn.cross_weights[i][...] = w["cross_w"][i]
        model.dcn.cross_biases[i][...] = w["cross_b"][i]
    for layer, kb in zip(model.dcn.deep_nets, w["deep"]):
        _set_dense(layer, kb)
    _set_dense(model.rating_head, w["rating_head"])
    _set_dense(model.ctr_head, w["ctr_head"])
    return cfg, model


API_SCRIPT_FUNCTIONS = ("test_health", "test_root", "test_model_info", "test_recommendations", "test_scoring",
                        "test_batch_recommendations")


def run_reference_api_script(client, reference_root="/root/reference"):
    """app/test_api.py is a script for a live server (requests against http://localhost:8000).  Its functions are called
    unmodified; only the transport is swapped: the module's `requests` name is bound to an object that sends the same
    calls through a FastAPI TestClient.  Per function: how it ended, the status codes it printed, whether it printed its
    own pass line."""
    import contextlib
    import importlib.util
    import io
    import re
    import requests

    spec = importlib.util.spec_from_file_location("reference_app_test_api", Path(reference_root) / "app" / "test_api.py")
    api = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(api)
    base = api.BASE_URL

    class ViaTestClient:
        exceptions = requests.exceptions

        @staticmethod
        def get(url, **kw):
            return client.get(url[len(base):] or "/", **kw)

        @staticmethod
        def post(url, **kw):
            return client.post(url[len(base):] or "/", **kw)

    api.requests = ViaTestClient
    outcomes = {}
    for name in API_SCRIPT_FUNCTIONS:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            try:
                getattr(api, name)()
                ended = "returned"
            except Exception as e:  # noqa: BLE001 -- the outcome is the data
                ended = f"{type(e).__name__}: {e}"
        text = buf.getvalue()
        outcomes[name] = {"ended": ended, "status": re.findall(r"Status: (\d+)", text), "pass_line": "\u2705" in text}
    return outcomes


def generate(reference_root="/root/reference"):
    import tf_faiss_standin as standin
    standin.install()
    standin.set_compute_dtype(np.float64)
    ref = standin.import_reference(reference_root)
    try:
        tf = sys.modules["tensorflow"]
        cfg0, w, uv, iv, users, items = synth_inputs()
        cfg, model = build_reference_model(ref, cfg0.to_dict(), w, uv, iv, users, items)

        arrays, doc = {}, {"spec": SPEC, "users_batch": users, "items_batch": items}
        # 1. ranking forward (src/models.py:125-131), float64 and float32 arithmetic
        for tag, dt in (("f64", np.float64), ("f32", np.float32)):
            standin.set_compute_dtype(dt)
            out = model({"user_id": tf.constant(users), "movie_id": tf.constant(items)}, training=False)
            assert list(out) == ["user_embedding", "item_embedding", "rating_prediction", "ctr_prediction"]
            for key, v in out.items():
                arrays[f"{tag}_{key}"] = v.numpy().astype(dt)
        standin.set_compute_dtype(np.float64)
        one_side = model.encoder({"user_id": tf.constant(users[:4])}, training=False)
        assert one_side["item_embedding"] is None                       # src/models.py:80-91: a missing side stays None

        for name in VARIANTS:
            vcfg, vw, vuv, viv, vusers, vitems = variant_inputs(name)
            _, vmodel = build_reference_model(ref, vcfg.to_dict(), vw, vuv, viv, vusers, vitems)
            vout = vmodel({"user_id": tf.constant(vusers), "movie_id": tf.constant(vitems)}, training=False)
            for key, val in vout.items():
                arrays[f"v{name}_f64_{key}"] = val.numpy().astype(np.float64)

        # 2. artifacts the way the trainer leaves them (src/trainer.py:228-246), then the service loads them (:35-59)
        T = ref["src.trainer"].ProductionTrainer
        with tempfile.TemporaryDirectory() as tmp:
            fake = type("TrainerState", (), {})()
            fake.output_dir, fake.config = Path(tmp), cfg
            fake._get_item_embeddings = lambda m, vocab: T._get_item_embeddings(fake, m, vocab)
            T._save_artifacts(fake, model, {"user_vocab": list(uv), "item_vocab": list(iv)})
            T._build_faiss(fake, model, list(iv))
            doc["artifact_files"] = sorted(p.name for p in Path(tmp).iterdir())
            doc["config_json"] = json.loads((Path(tmp) / "config.json").read_text())
            svc = ref["app.recommendation_service"].RecommendationService(tmp)
            assert not svc.is_ready()
            svc.load()
            assert svc.is_ready()
            info = svc.get_model_info()
            info["model_path"] = "<tmp>"
            doc["model_info"] = info
            arrays["catalog"] = np.array(svc.faiss_index._xb)           # the normalised rows _build_faiss added

            # 3. recommend / score / cold start (app/recommendation_service.py:64-102)
            doc["recommend"] = {u: svc.recommend(u, SPEC["k"]) for u in SPEC["users"]}
            doc["recommend_over"] = svc.recommend(SPEC["users"][0], SPEC["k_over"])      # k > ntotal: the -1 slots are dropped
            full = {u: svc.recommend(u, SPEC["n_items"]) for u in SPEC["users"]}          # every item's score, ranked
            pos = {it: i for i, it in enumerate(iv)}
            s64 = np.zeros((len(SPEC["users"]), SPEC["n_items"]))
            for r, u in enumerate(SPEC["users"]):
                uq = model.encoder({"user_id": tf.constant([u])}, training=False)["user_embedding"].numpy()
                uq = uq / np.linalg.norm(uq)
                s64[r] = arrays["catalog"].astype(np.float64) @ uq[0]
                assert [pos[x["item_id"]] for x in full[u]] == list(np.lexsort((np.arange(SPEC["n_items"]), -s64[r])))
            arrays["scores_f64"] = s64                                   # exact scores behind the float32 ones reported
            doc["cold_start"] = svc.recommend("nobody", 7)
            doc["cold_start_over"] = len(svc.recommend("nobody", 1000))
            score_ids = list(iv[:SPEC["score_items"]]) + ["no-such-movie"]
            doc["score_ids"] = score_ids
            doc["score"] = svc.score(SPEC["users"][2], score_ids)
            try:
                svc.score("nobody", score_ids[:2])
                doc["score_unknown_user"] = None
            except ValueError as e:
                doc["score_unknown_user"] = str(e)

            # 5. offline evaluation as the trainer does it (src/trainer.py:194-219): recall@k of the true item among the k
            #    largest RAW dots (no normalisation), over (user, true item) pairs; the true items are planted at known ranks
            import pandas as pd
            rng = np.random.default_rng(SPEC["id_seed"] + 1)
            val_users = [uv[i] for i in rng.integers(0, len(uv), 64)]
            item_embs = T._get_item_embeddings(fake, model, list(iv))
            user_embs = model.encoder({"user_id": tf.constant(val_users)}, training=False)["user_embedding"].numpy()
            raw = user_embs @ item_embs.T
            planted = [1, 3, 8, 15, 30, 70]
            true_items = [iv[int(np.lexsort((np.arange(len(iv)), -raw[r]))[planted[r % len(planted)] - 1])]
                          for r in range(len(val_users))]
            T._evaluate(fake, model, {"val_ds": object(), "item_vocab": list(iv),
                                      "val_df": pd.DataFrame({"user_id": val_users, "movie_id": true_items})})
            doc["offline_eval"] = {"users": val_users, "true_items": true_items, "planted_ranks": planted,
                                   "metrics": json.loads((Path(tmp) / "metrics.json").read_text())}

            # 6. the FastAPI application (app/main.py) over that service: a transcript of requests and answers
            from fastapi.testclient import TestClient
            import importlib
            main = importlib.import_module("app.main")
            client = TestClient(main.app)                                # no lifespan: start-up would look for outputs/models/...
            transcript = []

            def call(method, path, body=None):
                r = client.request(method, path, json=body)
                rec = {"method": method, "path": path, "json": body, "status": r.status_code}
                if r.status_code != 422:                                 # validation bodies are pydantic's wording, not the app's
                    rec["body"] = r.json()
                    if path == "/model/info" and r.status_code == 200:
                        rec["body"]["model_path"] = "<tmp>"
                transcript.append(rec)

            call("POST", "/recommend", {"user_id": "1", "k": 5})          # before a service is set: 503
            call("GET", "/health")
            main.recommendation_service = svc
            for req in HTTP_REQUESTS:
                call(*req)
            doc["http"] = transcript
            # 8. the alternate service the reference keeps beside it (app/model_service.py:72-90,104-150,200-228): item
            #    embeddings precomputed in 512-id batches, RAW dots (no normalisation), argpartition + argsort, and the
            #    batch call with its per-user status entries
            alt_mod = importlib.import_module("app.model_service")
            alt = alt_mod.RecommendationService(tmp)
            alt.model, alt.user_vocab, alt.item_vocab, alt.config = model.encoder, list(uv), list(iv), cfg.to_dict()
            alt._precompute_item_embeddings()
            assert alt.is_ready()
            batch_ids = [SPEC["users"][0], "nobody", SPEC["users"][1], SPEC["users"][0]]
            doc["alt_service"] = {"recommend": {u: alt.recommend(u, SPEC["k"]) for u in SPEC["users"]},
                                  "batch_ids": batch_ids, "batch": alt.recommend_batch(batch_ids, 5),
                                  "score": alt.score(SPEC["users"][2], score_ids),
                                  "user_norms": {u: float(np.linalg.norm(model.encoder(
                                      {"user_id": tf.constant([u])}, training=False)["user_embedding"].numpy())) for u in SPEC["users"]}}

            # 7. the reference's own API test script (app/test_api.py), its functions run as they are against that app
            doc["test_api"] = run_reference_api_script(client, reference_root)
            main.recommendation_service = None

        # 4. the ranked lists through AdvancedMetrics (src/evaluation.py:25-81)
        A = ref["src.evaluation"].AdvancedMetrics
        preds = [[x["item_id"] for x in doc["recommend"][u]] for u in SPEC["users"]]
        truth = [preds[0][0], preds[1][3], "no-such-movie", preds[3][9], preds[4][1]]
        doc["metrics_truth"] = truth
        doc["metrics"] = {
            "recall@5": float(A.recall_at_k(preds, truth, 5)), "recall@10": float(A.recall_at_k(preds, truth, 10)),
            "precision@5": float(A.precision_at_k(preds, truth, 5)), "ndcg@10": float(A.ndcg_at_k(preds, truth, 10)),
            "map@10": float(A.map_at_k(preds, truth, 10)), "mrr": float(A.mrr(preds, truth)),
            "coverage": float(A.coverage(preds, list(iv))), "diversity": float(A.diversity(preds))}
        return arrays, doc
    finally:
        standin.drop_reference()
        standin.uninstall()


def load_committed():
    with np.load(HERE / "reference_run.npz") as z:
        arrays = {k: z[k] for k in z.files}
    return arrays, json.loads((HERE / "reference_run.json").read_text())


if __name__ == "__main__":
    arrays, doc = generate()
    np.savez_compressed(HERE / "reference_run.npz", **arrays)
    (HERE / "reference_run.json").write_text(json.dumps(doc, indent=1, sort_keys=True) + "\n")
    top = doc["recommend"][SPEC["users"][0]]
    gaps = -np.diff(np.sort(arrays["scores_f64"], axis=1)[:, ::-1][:, :SPEC["k"] + 1], axis=1)
    print("wrote reference_run.npz / .json;", "first list:", [x["item_id"] for x in top][:5],
          "smallest gap among the top-(k+1) scores: %.3g" % gaps.min())
